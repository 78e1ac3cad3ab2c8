// atm_nb_lists.cuh -- Tier 2 pair-list kernels: spatial sort keys, bin scan, site placement, ghost links, the pack
// kernel (caller's slots -> cluster-order sites), cluster bounding boxes, the OUTER list build (radius cutoff + outer
// skin, with exclusion masks) and the prune to the INNER list that also emits the force kernel's work items.
// Included by atm_nb.cu only.
#pragma once

#include "atm_common.cuh"
#include "atm_nb_types.cuh"

namespace atm {

// ------------------------------------------------------------------------------------------------
// Rebuild step 1: sort keys (replica, bin, z) for every site; per-bin histogram.
// ------------------------------------------------------------------------------------------------
__global__ void nl_keys_kernel(NbDev d, const float4 *__restrict__ posq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.R * d.U) return;
    const int r = t / d.U, u = t - r * d.U;
    const bool ghost = u >= d.N;
    const int a = ghost ? d.ghost_atom[u - d.N] : u;
    const int slot = d.slot_of_atom[a];
    float4 p = __ldg(posq + (size_t)r * d.P + slot);
    if (ghost) {
        const float4 dd = __ldg(d.displ + slot);
        p.x = __fadd_rn(p.x, dd.x);
        p.y = __fadd_rn(p.y, dd.y);
        p.z = __fadd_rn(p.z, dd.z);
    }
    const float4 L = d.box[r], iL = d.invbox[r];
    float wx = p.x - L.x * floorf(p.x * iL.x), wy = p.y - L.y * floorf(p.y * iL.y), wz = p.z - L.z * floorf(p.z * iL.z);
    // bins are class-major: class 0 = environment, 1..G = displaced atoms of group g, G+1..2G = their ghosts; inside a
    // class one bin per xy column, so a displaced group of any size is cut into compact clusters like the environment
    const int g = d.group_of_atom[a];
    const int cls = g == 0 ? 0 : (ghost ? d.G + g : g);
    const int ix = min(max((int)(wx * iL.x * d.nx), 0), d.nx - 1);
    const int iy = min(max((int)(wy * iL.y * d.ny), 0), d.ny - 1);
    const int bin = cls * d.ncol + ix * d.ny + iy;
    const int zq = min(max((int)(wz * iL.z * 65536.0f), 0), 65535);
    d.keys[t] = ((unsigned long long)(r * d.nbins + bin) << 16) | (unsigned long long)zq;
    d.vals[t] = u;
    atomicAdd(&d.bin_count[r * d.nbins + bin], 1);
}

// ------------------------------------------------------------------------------------------------
// Own sort front end (replaces keys -> CUB radix sort -> place -> link -> pack -> bbox: 13 graph nodes become 4).
// The order produced is exactly the radix sort's: by (replica, bin), then z16, ties by site index -- so clusters, lists
// and results are bit-identical to the CUB path (which stays as the fallback for bins beyond bin_cap).
// ------------------------------------------------------------------------------------------------
// Step 1: every site into its bin's fixed-capacity segment, in arrival order (sorted later, inside the bin).
__global__ void nl_bin_kernel(NbDev d, const float4 *__restrict__ posq) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.R * d.U) return;
    const int r = t / d.U, u = t - r * d.U;
    const bool ghost = u >= d.N;
    const int a = ghost ? d.ghost_atom[u - d.N] : u;
    const int slot = d.slot_of_atom[a];
    float4 p = __ldg(posq + (size_t)r * d.P + slot);
    if (ghost) {
        const float4 dd = __ldg(d.displ + slot);
        p.x = __fadd_rn(p.x, dd.x);
        p.y = __fadd_rn(p.y, dd.y);
        p.z = __fadd_rn(p.z, dd.z);
    }
    const float4 L = d.box[r], iL = d.invbox[r];
    const float wx = p.x - L.x * floorf(p.x * iL.x), wy = p.y - L.y * floorf(p.y * iL.y), wz = p.z - L.z * floorf(p.z * iL.z);
    const int g = d.group_of_atom[a];
    const int cls = g == 0 ? 0 : (ghost ? d.G + g : g);
    const int ix = min(max((int)(wx * iL.x * d.nx), 0), d.nx - 1);
    const int iy = min(max((int)(wy * iL.y * d.ny), 0), d.ny - 1);
    const int bin = cls * d.ncol + ix * d.ny + iy;
    const int zq = min(max((int)(wz * iL.z * 65536.0f), 0), 65535);
    const int rb = r * d.nbins + bin;
    const int rank = atomicAdd(&d.bin_count[rb], 1);
    if (rank < d.bin_cap) d.binbuf[(size_t)rb * d.bin_cap + rank] = ((unsigned long long)zq << 32) | (unsigned int)u;
    else atomicOr(&d.flags[0], 4);   // a bin outgrew its segment: the structure is incomplete (poisoned like a list overflow)
}

// Step 3: one warp per (replica, bin): bitonic sort of the bin's entries in shared memory, then every slot of the bin's
// clusters is written -- real sites and the padding slots of the last cluster -- so no array needs clearing first.
constexpr int SORT_WARPS = 4;
__global__ void __launch_bounds__(32 * SORT_WARPS) nl_sort_place_kernel(NbDev d) {
    extern __shared__ unsigned long long s_keys[];   // [SORT_WARPS][bin_cap]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rb = blockIdx.x * SORT_WARPS + w;
    if (rb >= d.R * d.nbins) return;
    const int r = rb / d.nbins, bin = rb - r * d.nbins;
    const int start = d.bin_site_start[r * (d.nbins + 1) + bin];
    const int count = min(d.bin_site_start[r * (d.nbins + 1) + bin + 1] - start, d.bin_cap);
    if (count == 0) return;
    unsigned long long *k = s_keys + (size_t)w * d.bin_cap;
    int n = 2;
    while (n < count) n <<= 1;
    for (int i = lane; i < n; i += 32) k[i] = i < count ? d.binbuf[(size_t)rb * d.bin_cap + i] : ~0ull;
    __syncwarp();
    for (int kk = 2; kk <= n; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = k[i], b = k[ixj];
                    if ((a > b) == ((i & kk) == 0)) { k[i] = b; k[ixj] = a; }
                }
            }
            __syncwarp();
        }
    const int slot0 = CL * d.bin_cluster_start[r * (d.nbins + 1) + bin];
    const int nslots = CL * ((count + CL - 1) / CL);
    for (int rank = lane; rank < nslots; rank += 32) {
        const size_t rs = (size_t)r * d.Smax + slot0 + rank;
        if (rank < count) {
            const int u = (int)(unsigned int)(k[rank] & 0xffffffffull);
            const bool ghost = u >= d.N;
            const int a = ghost ? d.ghost_atom[u - d.N] : u;
            d.slot_site[rs] = u;
            d.site_slot[(size_t)r * d.U + u] = slot0 + rank;
            d.slot_src[rs] = (r * d.P + d.slot_of_atom[a]) | (ghost ? 0x80000000 : 0);
            d.slot_qp[rs] = d.qp_atom[a];
            d.par[rs] = d.par_atom[a];
            d.slot_out[rs] = ghost ? -1 : d.slot_of_atom[a];
        } else {   // padding of the bin's last cluster: read (masked) by the build, the prune and the force kernel
            d.slot_site[rs] = -1;
            d.slot_out[rs] = -1;
            d.par[rs] = make_float2(0.f, 0.f);
        }
        d.slot_ghost[rs] = -1;
    }
}

// Step 4: one thread per cluster: gathers the cluster's coordinates (as nb_pack_kernel does), links every displaced
// atom to its ghost site (nl_link_ghosts_kernel) and forms the bounding box (nl_bbox_kernel) in one pass.
__global__ void nl_pack_bbox_kernel(NbDev d, const float4 *__restrict__ posq) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= d.nclusters[r]) return;
    const float4 L = d.box[r], iL = d.invbox[r];
    const size_t base = (size_t)r * d.Smax + (size_t)c * CL;
    float3 lo = make_float3(0, 0, 0), hi = make_float3(0, 0, 0), x0 = make_float3(0, 0, 0);
    int valid = 0, cls = 0;
#pragma unroll
    for (int k = 0; k < CL; k++) {
        const int u = d.slot_site[base + k];
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u >= 0) {
            const int src = d.slot_src[base + k];
            const int idx = src & 0x7fffffff;
            p = __ldg(posq + idx);
            if (src < 0) {
                const float4 dd = __ldg(d.displ + (idx - r * d.P));
                p.x = __fadd_rn(p.x, dd.x);
                p.y = __fadd_rn(p.y, dd.y);
                p.z = __fadd_rn(p.z, dd.z);
            }
            p.w = d.slot_qp[base + k];
            const int a = u >= d.N ? d.ghost_atom[u - d.N] : u;
            if (u < d.N) {
                const int gm = d.ghost_of_atom[a];
                if (gm >= 0) d.slot_ghost[base + k] = d.site_slot[(size_t)r * d.U + d.N + gm];
            }
            if (!valid) {
                x0 = make_float3(p.x, p.y, p.z);
                const int g = d.group_of_atom[a];
                cls = g == 0 ? 0 : (u >= d.N ? d.G + g : g);
            }
            const float dx = wrap_delta(p.x - x0.x, L.x, iL.x), dy = wrap_delta(p.y - x0.y, L.y, iL.y),
                        dz = wrap_delta(p.z - x0.z, L.z, iL.z);
            lo.x = fminf(lo.x, dx); lo.y = fminf(lo.y, dy); lo.z = fminf(lo.z, dz);
            hi.x = fmaxf(hi.x, dx); hi.y = fmaxf(hi.y, dy); hi.z = fmaxf(hi.z, dz);
            valid |= 1 << k;
        }
        d.xs[base + k] = p;
    }
    const size_t rc = (size_t)r * d.Cmax + c;
    d.cc[rc] = make_float4(x0.x + 0.5f * (lo.x + hi.x), x0.y + 0.5f * (lo.y + hi.y), x0.z + 0.5f * (lo.z + hi.z), 0.f);
    d.ch[rc] = make_float4(0.5f * (hi.x - lo.x), 0.5f * (hi.y - lo.y), 0.5f * (hi.z - lo.z), 0.f);
    d.cmeta[rc] = cls | (valid << 16);
    const float rl = d.rlist_outer + (d.rlist - sqrtf(d.cutoff2));
    const float hmax_x = 0.5f * (hi.x - lo.x) + rl, hmax_y = 0.5f * (hi.y - lo.y) + rl, hmax_z = 0.5f * (hi.z - lo.z) + rl;
    if (hmax_x > 0.5f * L.x || hmax_y > 0.5f * L.y || hmax_z > 0.5f * L.z) atomicOr(&d.flags[0], 2);
}

// Rebuild step 2: per replica exclusive scans of the bin populations (sites and 8-padded clusters).
// zero_counts: the histogram is handed back zeroed for the next rebuild (own sort front end; its place step reads the
// scanned starts only).
__global__ void nl_scan_kernel(NbDev d, int zero_counts) {
    const int r = blockIdx.x;
    __shared__ int s_sites[1024], s_clusters[1024], s_wsum[2][32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int per = (d.nbins + nt - 1) / nt;
    const int b0 = min(tid * per, d.nbins), b1 = min(b0 + per, d.nbins);
    int ns = 0, nc = 0;
    for (int b = b0; b < b1; b++) {
        const int c = d.bin_count[r * d.nbins + b];
        ns += c;
        nc += (c + CL - 1) / CL;
    }
    // block-wide exclusive scan of the per-thread partial sums: warp shuffles, then the warp totals
    int is = ns, ic = nc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int vs = __shfl_up_sync(0xffffffffu, is, off), vc = __shfl_up_sync(0xffffffffu, ic, off);
        if (lane >= off) { is += vs; ic += vc; }
    }
    if (lane == 31) { s_wsum[0][wid] = is; s_wsum[1][wid] = ic; }
    __syncthreads();
    if (wid == 0) {
        const int nw = nt >> 5;
        int ws = lane < nw ? s_wsum[0][lane] : 0, wc = lane < nw ? s_wsum[1][lane] : 0;
        int xs_ = ws, xc_ = wc;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int vs = __shfl_up_sync(0xffffffffu, xs_, off), vc = __shfl_up_sync(0xffffffffu, xc_, off);
            if (lane >= off) { xs_ += vs; xc_ += vc; }
        }
        s_wsum[0][lane] = xs_ - ws;   // exclusive prefix of the warp totals
        s_wsum[1][lane] = xc_ - wc;
        if (lane == 31) {
            d.nclusters[r] = xc_;
            d.bin_site_start[r * (d.nbins + 1) + d.nbins] = xs_;
            d.bin_cluster_start[r * (d.nbins + 1) + d.nbins] = xc_;
        }
    }
    __syncthreads();
    s_sites[tid] = is - ns + s_wsum[0][wid];
    s_clusters[tid] = ic - nc + s_wsum[1][wid];
    ns = s_sites[tid];
    nc = s_clusters[tid];
    for (int b = b0; b < b1; b++) {
        const int c = d.bin_count[r * d.nbins + b];
        d.bin_site_start[r * (d.nbins + 1) + b] = ns;
        d.bin_cluster_start[r * (d.nbins + 1) + b] = nc;
        if (zero_counts) d.bin_count[r * d.nbins + b] = 0;
        ns += c;
        nc += (c + CL - 1) / CL;
    }
}

// Rebuild step 3: sorted position -> padded slot; static per-slot gather info.
__global__ void nl_place_kernel(NbDev d, const unsigned long long *__restrict__ keys_sorted,
                                const int *__restrict__ vals_sorted) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.R * d.U) return;
    const int r = t / d.U, idx = t - r * d.U;
    const int u = vals_sorted[t];
    const int bin = (int)(keys_sorted[t] >> 16) - r * d.nbins;
    const int rank = idx - d.bin_site_start[r * (d.nbins + 1) + bin];
    const int slot = CL * d.bin_cluster_start[r * (d.nbins + 1) + bin] + rank;
    const bool ghost = u >= d.N;
    const int a = ghost ? d.ghost_atom[u - d.N] : u;
    const size_t rs = (size_t)r * d.Smax + slot;
    d.slot_site[rs] = u;
    d.site_slot[(size_t)r * d.U + u] = slot;
    d.slot_src[rs] = (r * d.P + d.slot_of_atom[a]) | (ghost ? 0x80000000 : 0);
    d.slot_qp[rs] = d.qp_atom[a];
    d.par[rs] = d.par_atom[a];
    d.slot_out[rs] = ghost ? -1 : d.slot_of_atom[a];
}

// after every site has its slot: link each displaced atom's real slot to its ghost slot
__global__ void nl_link_ghosts_kernel(NbDev d) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.R * d.M) return;
    const int r = t / d.M, m = t - r * d.M;
    const int a = d.ghost_atom[m];
    const int s_real = d.site_slot[(size_t)r * d.U + a], s_gh = d.site_slot[(size_t)r * d.U + d.N + m];
    d.slot_ghost[(size_t)r * d.Smax + s_real] = s_gh;
}

// Every step: gather current coordinates into cluster order; ghosts get posq + displ (the same float add as
// CopyState, so a ghost sits exactly at the reference's posq2).
// zero_eacc = 0: the side-stream pack of a concurrent prune (it must not touch the accumulators of the step in flight).
__global__ void nb_pack_kernel(NbDev d, const float4 *__restrict__ posq, int zero_eacc) {
    pdl_trigger();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    // first kernel of a step: the energy / pair-count accumulators of the previous step have been consumed by its merge
    if (zero_eacc && blockIdx.x == 0 && threadIdx.x < EACC_SLOTS) d.eacc[(size_t)r * EACC_SLOTS + threadIdx.x] = 0ull;
    if (t >= CL * d.nclusters[r]) return;
    const size_t rs = (size_t)r * d.Smax + t;
    const int src = d.slot_src[rs];
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.slot_site[rs] >= 0) {
        const int idx = src & 0x7fffffff;
        p = __ldg(posq + idx);
        if (src < 0) {
            const float4 dd = __ldg(d.displ + (idx - r * d.P));
            p.x = __fadd_rn(p.x, dd.x);
            p.y = __fadd_rn(p.y, dd.y);
            p.z = __fadd_rn(p.z, dd.z);
        }
        p.w = d.slot_qp[rs];
    }
    d.xs[rs] = p;
}

// Rebuild step 4: cluster bounding boxes (centre, half extent) in the frame of the first member, class, valid mask.
__global__ void nl_bbox_kernel(NbDev d) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= d.nclusters[r]) return;
    const float4 L = d.box[r], iL = d.invbox[r];
    const size_t base = (size_t)r * d.Smax + (size_t)c * CL;
    float3 lo = make_float3(0, 0, 0), hi = make_float3(0, 0, 0), x0 = make_float3(0, 0, 0);
    int valid = 0, cls = 0;
    for (int k = 0; k < CL; k++) {
        const int u = d.slot_site[base + k];
        if (u < 0) continue;
        const float4 p = d.xs[base + k];
        if (!valid) {
            x0 = make_float3(p.x, p.y, p.z);
            const int a = u >= d.N ? d.ghost_atom[u - d.N] : u;
            const int g = d.group_of_atom[a];
            cls = g == 0 ? 0 : (u >= d.N ? d.G + g : g);
        }
        const float dx = wrap_delta(p.x - x0.x, L.x, iL.x), dy = wrap_delta(p.y - x0.y, L.y, iL.y),
                    dz = wrap_delta(p.z - x0.z, L.z, iL.z);
        lo.x = fminf(lo.x, dx); lo.y = fminf(lo.y, dy); lo.z = fminf(lo.z, dz);
        hi.x = fmaxf(hi.x, dx); hi.y = fmaxf(hi.y, dy); hi.z = fmaxf(hi.z, dz);
        valid |= 1 << k;
    }
    const size_t rc = (size_t)r * d.Cmax + c;
    d.cc[rc] = make_float4(x0.x + 0.5f * (lo.x + hi.x), x0.y + 0.5f * (lo.y + hi.y), x0.z + 0.5f * (lo.z + hi.z), 0.f);
    d.ch[rc] = make_float4(0.5f * (hi.x - lo.x), 0.5f * (hi.y - lo.y), 0.5f * (hi.z - lo.z), 0.f);
    d.cmeta[rc] = cls | (valid << 16);
    // the per-step image shift of a partner relative to the cluster centre needs half extent + list radius <= L/2; the
    // centres stay those of the rebuild until the next one (atoms drift by up to half the outer skin, partners too:
    // the outer skin is in rlist_outer, the inner skin is added for the pruned list built from drifted coordinates)
    const float rl = d.rlist_outer + (d.rlist - sqrtf(d.cutoff2));
    const float hmax_x = 0.5f * (hi.x - lo.x) + rl, hmax_y = 0.5f * (hi.y - lo.y) + rl, hmax_z = 0.5f * (hi.z - lo.z) + rl;
    if (hmax_x > 0.5f * L.x || hmax_y > 0.5f * L.y || hmax_z > 0.5f * L.z) atomicOr(&d.flags[0], 2);
}

// ------------------------------------------------------------------------------------------------
// Rebuild step 5: one warp per list.  List l of replica r:
//   l <  Cmax            : primary list of cluster l   (env cluster -> C, ligand cluster -> S1, ghost cluster -> S2)
//   l >= Cmax            : secondary list (same-group pairs, target C) of ligand cluster firstL + (l - Cmax)
// ------------------------------------------------------------------------------------------------
struct ListInfo {
    int cluster, target, cap;
    size_t offset;
    bool valid;
};

__device__ __forceinline__ ListInfo decode_list(const NbDev &d, int r, int l) {
    ListInfo li;
    const int *bcs = d.bin_cluster_start + (size_t)r * (d.nbins + 1);
    const int nenv = bcs[d.ncol], firstG = bcs[d.ncol * (d.G + 1)], ncl = d.nclusters[r];
    const size_t per_replica = (size_t)d.CenvMax * d.capC + (size_t)d.CXmax * d.capX + (size_t)d.CLmax * d.capC;
    const size_t rbase = (size_t)r * per_replica;
    li.valid = false;
    li.cluster = 0; li.target = TGT_C; li.cap = d.capC; li.offset = rbase;
    if (l < d.Cmax) {
        if (l >= ncl) return li;
        li.cluster = l;
        if (l < nenv) {
            li.target = TGT_C;
            li.cap = d.capC;
            li.offset = rbase + (size_t)l * d.capC;
        } else {
            li.target = l < firstG ? TGT_S1 : TGT_S2;
            li.cap = d.capX;
            li.offset = rbase + (size_t)d.CenvMax * d.capC + (size_t)(l - nenv) * d.capX;
        }
        li.valid = true;
    } else {
        const int k = l - d.Cmax;
        if (k >= firstG - nenv) return li;
        li.cluster = nenv + k;
        li.target = TGT_C;
        li.cap = d.capC;
        li.offset = rbase + (size_t)d.CenvMax * d.capC + (size_t)d.CXmax * d.capX + (size_t)k * d.capC;
        li.valid = true;
    }
    return li;
}

constexpr int BUILD_WARPS = 4;
constexpr int TBL_CAP = 160;  // exclusion partners of one cluster kept in shared memory

// OUTER list: every site inside (cutoff + outer skin) of the cluster's bounding box, with exclusion masks.
// Built rarely; the per-step work uses the pruned INNER list (nl_prune_kernel).
__global__ void __launch_bounds__(32 * BUILD_WARPS) nl_build_kernel(NbDev d) {
    __shared__ int s_pass[BUILD_WARPS][36];   // queue of clusters that passed stage 1: a batch of up to 32 behind up to 3 waiting ones
    __shared__ int2 s_tbl[BUILD_WARPS][TBL_CAP];
    __shared__ int s_tn[BUILD_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int l = blockIdx.x * BUILD_WARPS + w;
    const int r = blockIdx.y;
    const int nlists = d.Cmax + d.CLmax;
    if (l >= nlists) return;
    const ListInfo li = decode_list(d, r, l);
    int *nsteps_out = d.outer_nsteps + (size_t)r * nlists + l;
    if (!li.valid) {
        if (lane == 0) *nsteps_out = 0;
        return;
    }
    const int A = li.cluster;
    const size_t rcA = (size_t)r * d.Cmax + A;
    const size_t rsite = (size_t)r * d.Smax;
    const float4 cA = d.cc[rcA], hA = d.ch[rcA];
    const int metaA = d.cmeta[rcA];
    const int clsA = metaA & 0xffff, validA = (metaA >> 16) & 0xff;
    const float4 L = d.box[r], iL = d.invbox[r];
    const float rl2 = d.rlist_outer * d.rlist_outer;
    const int *bcs = d.bin_cluster_start + (size_t)r * (d.nbins + 1);
    const int nenv = bcs[d.ncol], ncl = d.nclusters[r];
    unsigned int *out = d.jlist_outer + li.offset;
    int count = 0;

    // exclusion table of this cluster: (partner slot, bit of the member that excludes it)
    if (lane == 0) s_tn[w] = 0;
    __syncwarp();
    if (lane < CL && ((validA >> lane) & 1)) {
        const int u = d.slot_site[rsite + (size_t)A * CL + lane];
        const int a = u >= d.N ? d.ghost_atom[u - d.N] : u;
        for (int e = d.excl_start[a]; e < d.excl_start[a + 1]; e++) {
            const int b = d.excl_list[e];
            int idx = atomicAdd(&s_tn[w], 1);
            if (idx < TBL_CAP) s_tbl[w][idx] = make_int2(d.site_slot[(size_t)r * d.U + b], 1 << lane);
            const int gm = d.ghost_of_atom[b];
            if (gm >= 0) {
                idx = atomicAdd(&s_tn[w], 1);
                if (idx < TBL_CAP) s_tbl[w][idx] = make_int2(d.site_slot[(size_t)r * d.U + d.N + gm], 1 << lane);
            }
        }
    }
    __syncwarp();
    const int T = s_tn[w];
    const bool tbl_ok = T <= TBL_CAP;
    // 64-bit Bloom filter over the clusters that hold an exclusion partner: almost every candidate site skips the table
    unsigned long long bloom = 0ull;
    if (tbl_ok) {
        for (int t = lane; t < T; t += 32) bloom |= 1ull << ((s_tbl[w][t].x >> 3) & 63);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) bloom |= __shfl_xor_sync(0xffffffffu, bloom, off);
    }

    // candidate cluster ranges: env columns near A (only when env clusters can be partners), then all ligand/ghost clusters
    const float cwx = cA.x - L.x * floorf(cA.x * iL.x), cwy = cA.y - L.y * floorf(cA.y * iL.y);
    const float colw_x = L.x / d.nx, colw_y = L.y / d.ny;
    int ix_lo = (int)floorf((cwx - hA.x - d.rlist_outer) / colw_x), ix_hi = (int)floorf((cwx + hA.x + d.rlist_outer) / colw_x);
    int iy_lo = (int)floorf((cwy - hA.y - d.rlist_outer) / colw_y), iy_hi = (int)floorf((cwy + hA.y + d.rlist_outer) / colw_y);
    if (ix_hi - ix_lo + 1 >= d.nx) { ix_lo = 0; ix_hi = d.nx - 1; }
    if (iy_hi - iy_lo + 1 >= d.ny) { iy_lo = 0; iy_hi = d.ny - 1; }
    const bool env_partners = pair_target(clsA, 0, d.G) == li.target;  // does this list take env sites at all?
    const int n_ix = env_partners ? ix_hi - ix_lo + 1 : 0;
    // iy range may wrap: split into up to two contiguous bin segments
    int seg_lo[2], seg_hi[2], nseg = 0;
    if (iy_lo >= 0 && iy_hi < d.ny) { seg_lo[0] = iy_lo; seg_hi[0] = iy_hi; nseg = 1; }
    else if (iy_lo < 0) { seg_lo[0] = 0; seg_hi[0] = iy_hi; seg_lo[1] = iy_lo + d.ny; seg_hi[1] = d.ny - 1; nseg = 2; }
    else { seg_lo[0] = iy_lo; seg_hi[0] = d.ny - 1; seg_lo[1] = 0; seg_hi[1] = iy_hi - d.ny; nseg = 2; }

    // stage 2 of the search, one candidate SITE per lane: the sites of up to four clusters of the queue (from entry q) are
    // tested against the bounding box of A, masked by the exclusions, and appended to the list in queue order
    auto sweep = [&](int q, int nq) {
        const int g = q + (lane >> 3), k = lane & 7;
        bool take = false;
        unsigned int entry = 0;
        if ((lane >> 3) < nq) {
            const int B2 = s_pass[w][g];
            const int j = B2 * CL + k;
            const int u = d.slot_site[rsite + j];
            if (u >= 0) {
                const float4 p = __ldg(d.xs + rsite + j);
                const float bx = fmaxf(fabsf(wrap_delta(p.x - cA.x, L.x, iL.x)) - hA.x, 0.f);
                const float by = fmaxf(fabsf(wrap_delta(p.y - cA.y, L.y, iL.y)) - hA.y, 0.f);
                const float bz = fmaxf(fabsf(wrap_delta(p.z - cA.z, L.z, iL.z)) - hA.z, 0.f);
                if (bx * bx + by * by + bz * bz <= rl2) {
                    unsigned int m = (~validA) & 0xff;
                    if (B2 == A) m |= (0xffu << k) & 0xff;  // within a cluster: pairs (i<j) once
                    if (tbl_ok) {
                        if ((bloom >> (B2 & 63)) & 1ull) {
                            for (int t = 0; t < T; t++) {
                                const int2 te = s_tbl[w][t];
                                if (te.x == j) m |= te.y;
                            }
                        }
                    } else {  // rare: a cluster with more exclusion partners than the table holds
                        const int aj = u >= d.N ? d.ghost_atom[u - d.N] : u;
                        for (int e = d.excl_start[aj]; e < d.excl_start[aj + 1]; e++) {
                            const int b = d.excl_list[e];
                            const int s_real = d.site_slot[(size_t)r * d.U + b];
                            if ((s_real >> 3) == A) m |= 1u << (s_real & 7);
                            const int gm = d.ghost_of_atom[b];
                            if (gm >= 0) {
                                const int s_gh = d.site_slot[(size_t)r * d.U + d.N + gm];
                                if ((s_gh >> 3) == A) m |= 1u << (s_gh & 7);
                            }
                        }
                    }
                    if (m != 0xff) {
                        take = true;
                        entry = ((unsigned int)j << 8) | m;
                    }
                }
            }
        }
        const unsigned int tmask = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int pos = count + __popc(tmask & ((1u << lane) - 1));
            if (pos < li.cap) out[pos] = entry;
        }
        count += __popc(tmask);
    };
    int npend = 0;   // clusters that passed stage 1 and wait in the queue for a FULL sweep of four
    const int n_ranges = n_ix * nseg + 1;
    for (int rg = 0; rg < n_ranges; rg++) {
        int c_begin, c_end;
        if (rg < n_ix * nseg) {
            int ix = ix_lo + rg / nseg;
            ix = ((ix % d.nx) + d.nx) % d.nx;
            const int sgi = rg % nseg;
            c_begin = bcs[ix * d.ny + seg_lo[sgi]];
            c_end = bcs[ix * d.ny + seg_hi[sgi] + 1];
        } else {
            c_begin = nenv;
            c_end = ncl;
        }
        // An environment cluster A owns, among the environment clusters of a range, those below it with the parity of A,
        // itself, and those above it with the other parity (the ownership rule below): enumerate exactly those -- half
        // the candidates, every lane on a cluster A can own, same ascending order.
        const bool by_parity = clsA == 0 && rg < n_ix * nseg;
        int n_cand = c_end - c_begin, lo0 = 0, n_lo = 0, n_self = 0, hi0 = 0;
        if (by_parity) {
            const int lo_end = min(c_end, A), hi_begin = max(c_begin, A + 1);
            lo0 = c_begin + ((c_begin ^ A) & 1);
            n_lo = lo_end > lo0 ? (lo_end - lo0 + 1) >> 1 : 0;
            n_self = (A >= c_begin && A < c_end) ? 1 : 0;
            hi0 = hi_begin + (1 - ((hi_begin ^ A) & 1));
            n_cand = n_lo + n_self + (c_end > hi0 ? (c_end - hi0 + 1) >> 1 : 0);
        }
        for (int base = 0; base < n_cand; base += 32) {
            // stage 1: one candidate cluster per lane, box-box distance; passing clusters compacted (in order) to smem
            const int t = base + lane;
            const int B = !by_parity ? c_begin + t : (t < n_lo ? lo0 + 2 * t : (t < n_lo + n_self ? A : hi0 + 2 * (t - n_lo - n_self)));
            bool pass = false;
            if (t < n_cand) {
                const size_t rcB = (size_t)r * d.Cmax + B;
                const int clsB = by_parity ? 0 : d.cmeta[rcB] & 0xffff;
                bool owner;
                if (clsA == clsB) owner = (A == B) || (((A + B) & 1) ? (A < B) : (A > B));
                else owner = clsA > clsB;
                if (owner && pair_target(clsA, clsB, d.G) == li.target) {
                    const float4 cB = d.cc[rcB], hB = d.ch[rcB];
                    const float dx = fmaxf(fabsf(wrap_delta(cB.x - cA.x, L.x, iL.x)) - hA.x - hB.x, 0.f);
                    const float dy = fmaxf(fabsf(wrap_delta(cB.y - cA.y, L.y, iL.y)) - hA.y - hB.y, 0.f);
                    const float dz = fmaxf(fabsf(wrap_delta(cB.z - cA.z, L.z, iL.z)) - hA.z - hB.z, 0.f);
                    pass = dx * dx + dy * dy + dz * dz <= rl2;
                }
            }
            const unsigned int cmask = __ballot_sync(0xffffffffu, pass);
            const int npass = __popc(cmask);
            if (npass == 0) continue;
            if (pass) s_pass[w][npend + __popc(cmask & ((1u << lane) - 1))] = B;
            npend += npass;
            __syncwarp();
            // full sweeps only; the remainder (< 4 clusters) stays at the front of the queue for the next batch, so that
            // every sweep but the very last one of a list has all 32 lanes on candidate sites (same order of entries)
            int q = 0;
            for (; q + 4 <= npend; q += 4) sweep(q, 4);
            const int rem = npend - q;
            const int keep = lane < rem ? s_pass[w][q + lane] : 0;
            __syncwarp();
            if (lane < rem) s_pass[w][lane] = keep;
            npend = rem;
            __syncwarp();
        }
    }
    if (npend > 0) sweep(0, npend);
    // longest list of each capacity class, every build: the host grows a capacity BEFORE a list can outgrow it
    if (lane == 0) atomicMax(&d.flags[li.cap == d.capC ? FLAG_MAXLEN_C : FLAG_MAXLEN_X], count);
    if (count > li.cap) {
        // the list does not fit: flag it.  The scalar stage of every later step sees the flag and poisons the energy
        // record and the merged forces with NaN until the next rebuild clears it -- a truncated list never yields
        // silently wrong numbers
        if (lane == 0) {
            atomicOr(&d.flags[0], 1);
            atomicMax(&d.flags[1], count);
            *nsteps_out = 0;
        }
        return;
    }
    // pad the tail of the last 32-entry step with masked sentinels
    const int nsteps = (count + 31) >> 5;
    for (int p = count + lane; p < nsteps * 32; p += 32) out[p] = 0xffu;
    if (lane == 0) {
        *nsteps_out = nsteps;
        atomicAdd((unsigned long long *)&d.flags[2], (unsigned long long)count);
        atomicAdd(&d.flags[5], (nsteps + ITEM_STEPS - 1) / ITEM_STEPS);
    }
}

// INNER list: the outer entries that are inside (cutoff + inner skin) of at least one ATOM of the cluster at the
// current coordinates.  Cheap (coalesced reads of the outer list, no exclusion work), run every few steps.
// Prunes list l of replica r (one warp); returns the number of kept entries (0 for an empty / unused list).
__device__ __forceinline__ int prune_one_list(const NbDev &d, int r, int l, int nlists, int lane, ListInfo &li, float4 *sa) {
    if (l >= nlists) return 0;
    const int nst_outer = d.outer_nsteps[(size_t)r * nlists + l];
    int *nsteps_out = d.list_nsteps + (size_t)r * nlists + l;
    if (nst_outer == 0) {
        if (lane == 0) *nsteps_out = 0;
        return 0;
    }
    li = decode_list(d, r, l);
    const int A = li.cluster;
    const size_t rcA = (size_t)r * d.Cmax + A;
    const size_t rsite = (size_t)r * d.Smax;
    const float4 cA = d.cc[rcA];
    const int validA = (d.cmeta[rcA] >> 16) & 0xff;
    const float4 L = d.box[r], iL = d.invbox[r];
    const float rl2 = d.rlist * d.rlist;
    // cluster atoms relative to the cluster centre as (-2a, |a|^2): |p - a|^2 = |p|^2 + (-2a).p + |a|^2 costs three
    // FFMAs and a min per atom; all coordinates are within ~1.5 nm of the centre, so the expansion loses nothing that
    // matters for a skin test.  They live in shared memory (broadcast reads): 32 fewer registers per thread buy the
    // occupancy this latency-bound kernel needs.
    if (lane < CL) {
        const float4 p = __ldg(d.xs + rsite + (size_t)A * CL + lane);
        const bool ok = (validA >> lane) & 1;
        const float ax = wrap_delta(p.x - cA.x, L.x, iL.x), ay = wrap_delta(p.y - cA.y, L.y, iL.y), az = wrap_delta(p.z - cA.z, L.z, iL.z);
        const float aw = ok ? fmaf(az, az, fmaf(ay, ay, ax * ax)) : 1e30f;
        sa[lane] = make_float4(-2.f * ax, -2.f * ay, -2.f * az, aw);
        // packed copy for the f32x2 distance test: pair p = atoms 2p, 2p+1
        float *q0 = reinterpret_cast<float *>(sa + CL + (lane >> 1)) + (lane & 1);
        float *q1 = reinterpret_cast<float *>(sa + CL + CL / 2 + (lane >> 1)) + (lane & 1);
        q0[0] = -2.f * ax; q0[2] = -2.f * ay;
        q1[0] = -2.f * az; q1[2] = aw;
    }
    __syncwarp();
    const unsigned int *in = d.jlist_outer + li.offset;
    unsigned int *out = d.jlist + li.offset;
    int count = 0;
#if ATM_PRUNE_BLOCK > 1
    // software pipeline in blocks of PB list steps: the entries (streaming from DRAM) are loaded two blocks ahead, the
    // partner coordinates (gathers from L2) one block ahead, so PB independent gathers are in flight per lane while
    // the previous block is tested.  The kept entries are written in list order, exactly as the one-step loop did.
    constexpr int PB = ATM_PRUNE_BLOCK;
    unsigned int ea[PB], eb[PB];
    float4 pa[PB];
#pragma unroll
    for (int q = 0; q < PB; q++) ea[q] = q < nst_outer ? __ldg(in + q * 32 + lane) : 0xffu;
#pragma unroll
    for (int q = 0; q < PB; q++) eb[q] = PB + q < nst_outer ? __ldg(in + (PB + q) * 32 + lane) : 0xffu;
#pragma unroll
    for (int q = 0; q < PB; q++) pa[q] = __ldg(d.xs + rsite + (ea[q] >> 8));
    for (int st0 = 0; st0 < nst_outer; st0 += PB) {
        unsigned int ecur[PB], en[PB];
        float4 pcur[PB];
#pragma unroll
        for (int q = 0; q < PB; q++) { ecur[q] = ea[q]; pcur[q] = pa[q]; ea[q] = eb[q]; }
#pragma unroll
        for (int q = 0; q < PB; q++) en[q] = st0 + 2 * PB + q < nst_outer ? __ldg(in + (st0 + 2 * PB + q) * 32 + lane) : 0xffu;
        if (st0 + PB < nst_outer) {
#pragma unroll
            for (int q = 0; q < PB; q++) pa[q] = __ldg(d.xs + rsite + (ea[q] >> 8));
        }
#pragma unroll
        for (int q = 0; q < PB; q++) eb[q] = en[q];
#pragma unroll
        for (int q = 0; q < PB; q++) {
            const unsigned int ec = ecur[q];
            bool keep = false;
            if ((ec & 0xffu) != 0xffu) {
                const float4 p = pcur[q];
                const float px = wrap_delta(p.x - cA.x, L.x, iL.x), py = wrap_delta(p.y - cA.y, L.y, iL.y),
                            pz = wrap_delta(p.z - cA.z, L.z, iL.z);
                // the 8 atoms as 4 packed pairs (fma.rn.f32x2): sp[k] = (-2ax0,-2ax1,-2ay0,-2ay1), sp[k+4] = (-2az0,-2az1,|a0|^2,|a1|^2)
                const float4 *sp = sa + CL;
                float d2min = 1e30f;
#pragma unroll
                for (int k = 0; k < CL / 2; k++) {
                    const float4 u = sp[k], v = sp[k + CL / 2];
                    const float2 t = __ffma2_rn(make_float2(px, px), make_float2(u.x, u.y),
                                                __ffma2_rn(make_float2(py, py), make_float2(u.z, u.w),
                                                           __ffma2_rn(make_float2(pz, pz), make_float2(v.x, v.y), make_float2(v.z, v.w))));
                    d2min = fminf(d2min, fminf(t.x, t.y));
                }
                keep = d2min + fmaf(pz, pz, fmaf(py, py, px * px)) <= rl2;
            }
            const unsigned int kmask = __ballot_sync(0xffffffffu, keep);
            if (keep) out[count + __popc(kmask & ((1u << lane) - 1))] = ec;
            count += __popc(kmask);
        }
    }
#else
    // software pipeline: entries three steps ahead (they stream from DRAM), coordinates one step ahead
    unsigned int e0 = __ldg(in + lane);
    unsigned int e1 = nst_outer > 1 ? __ldg(in + 32 + lane) : 0xffu;
    unsigned int e2 = nst_outer > 2 ? __ldg(in + 64 + lane) : 0xffu;
    float4 pnext = __ldg(d.xs + rsite + (e0 >> 8));
    for (int st = 0; st < nst_outer; st++) {
        const unsigned int ec = e0;
        const float4 p = pnext;
        e0 = e1;
        e1 = e2;
        e2 = (st + 3 < nst_outer) ? __ldg(in + (st + 3) * 32 + lane) : 0xffu;
        if (st + 1 < nst_outer) pnext = __ldg(d.xs + rsite + (e0 >> 8));
        bool keep = false;
        if ((ec & 0xffu) != 0xffu) {
            const float px = wrap_delta(p.x - cA.x, L.x, iL.x), py = wrap_delta(p.y - cA.y, L.y, iL.y),
                        pz = wrap_delta(p.z - cA.z, L.z, iL.z);
            float d2min = 1e30f;
#pragma unroll
            for (int k = 0; k < CL; k++) { const float4 a = sa[k]; d2min = fminf(d2min, fmaf(px, a.x, fmaf(py, a.y, fmaf(pz, a.z, a.w)))); }
            keep = d2min + fmaf(pz, pz, fmaf(py, py, px * px)) <= rl2;
        }
        const unsigned int kmask = __ballot_sync(0xffffffffu, keep);
        if (keep) out[count + __popc(kmask & ((1u << lane) - 1))] = ec;
        count += __popc(kmask);
    }
#endif
    const int nsteps = (count + 31) >> 5;
    for (int p = count + lane; p < nsteps * 32; p += 32) out[p] = 0xffu;
    if (lane == 0) *nsteps_out = nsteps;
    return count;
}

#ifndef ATM_PRUNE_WARPS
#define ATM_PRUNE_WARPS 4
#endif
constexpr int PRUNE_WARPS = ATM_PRUNE_WARPS;   // lists per block of the prune kernel

// Measured (B200, 22 / 3 replicas of the 23k-atom system, whole prune call): one-step loop with the cluster atoms in
// registers (107 registers, 16 warps / SM) 315 / 62 us; blocks of 2 steps with the cluster atoms in shared memory at
// <= 64 registers (32 warps / SM) 228 / 50 us; blocks of 4 at 80 registers 243 / 53 us; 48 / 40 / 32 registers: 272 /
// 250 / 277 us (spills).
#ifndef ATM_PRUNE_MIN_BLOCKS
#define ATM_PRUNE_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(32 * PRUNE_WARPS, ATM_PRUNE_MIN_BLOCKS) nl_prune_kernel(NbDev d) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int l = blockIdx.x * PRUNE_WARPS + w;
    const int r = blockIdx.y;
    ListInfo li;
    li.cluster = 0; li.target = TGT_C; li.offset = 0;
    __shared__ float4 s_atoms[PRUNE_WARPS][2 * CL];   // per-atom form + the packed-pair form of the same 8 atoms
    const int count = prune_one_list(d, r, l, d.Cmax + d.CLmax, lane, li, s_atoms[w]);
    const int A = li.cluster;
    const int nsteps = (count + 31) >> 5;
    (void)A;
    // work items of this list: (<= ITEM_STEPS)-step chunks (their order only affects scheduling: every accumulation
    // downstream is fixed point, hence order independent).  Buckets by chunk length: the force kernel hands out the
    // longest chunks first (longest-processing-time order keeps the tail of the launch short when only a few replicas
    // share the GPU).  The counters every list touches -- kept entries, live items, the bucket of full chunks -- are
    // summed over the block first: same-address atomics serialise in the L2, one per block instead of one per list.
#ifdef ATM_EVEN_SPLIT
    // A list longer than ITEM_STEPS is cut into EQUAL chunks (17 steps -> 9 + 8 instead of 16 + 1): no work item is
    // mostly prologue.  Chunk lengths differ from list to list, so every chunk length has its own bucket counter.
    const int nch = (nsteps + ITEM_STEPS - 1) / ITEM_STEPS;
    const int len_lo = nch > 0 ? nsteps / nch : 0, n_hi = nch > 0 ? nsteps - len_lo * nch : 0;   // n_hi chunks of len_lo + 1 first
    __shared__ int s_count[PRUNE_WARPS], s_items[PRUNE_WARPS];
    if (lane == 0) { s_count[w] = count; s_items[w] = nch; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tc = 0, ti = 0;
#pragma unroll
        for (int k = 0; k < PRUNE_WARPS; k++) { tc += s_count[k]; ti += s_items[k]; }
        if (tc > 0) atomicAdd((unsigned long long *)&d.iflags[6], (unsigned long long)tc);
        if (ti > 0) atomicAdd(&d.iflags[4], ti);
    }
    int base_hi = 0, base_lo = 0;
    if (lane == 0) {
        if (n_hi > 0) base_hi = atomicAdd(&d.iflags[ITEM_BUCKET0 + len_lo + 1], n_hi);
        if (nch - n_hi > 0 && len_lo > 0) base_lo = atomicAdd(&d.iflags[ITEM_BUCKET0 + len_lo], nch - n_hi);
    }
    base_hi = __shfl_sync(0xffffffffu, base_hi, 0);
    base_lo = __shfl_sync(0xffffffffu, base_lo, 0);
    for (int c = lane; c < nch; c += 32) {
        const bool hi = c < n_hi;
        const int len = hi ? len_lo + 1 : len_lo;
        const int first = hi ? c * (len_lo + 1) : n_hi * (len_lo + 1) + (c - n_hi) * len_lo;
        d.items[(size_t)len * d.max_items + (hi ? base_hi + c : base_lo + (c - n_hi))] =
            make_int4((int)(li.offset + (size_t)first * 32), A | (li.target << 28), r | (len << 8), first);
    }
#else
    const int nfull = nsteps / ITEM_STEPS, rem = nsteps - nfull * ITEM_STEPS;
    __shared__ int s_count[PRUNE_WARPS], s_nfull[PRUNE_WARPS], s_items[PRUNE_WARPS], s_base_full;
    if (lane == 0) { s_count[w] = count; s_nfull[w] = nfull; s_items[w] = nfull + (rem > 0 ? 1 : 0); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tc = 0, tf = 0, ti = 0;
#pragma unroll
        for (int k = 0; k < PRUNE_WARPS; k++) { tc += s_count[k]; tf += s_nfull[k]; ti += s_items[k]; }
        if (tc > 0) atomicAdd((unsigned long long *)&d.iflags[6], (unsigned long long)tc);
        s_base_full = tf > 0 ? atomicAdd(&d.iflags[ITEM_BUCKET0 + ITEM_STEPS], tf) : 0;
        if (ti > 0) atomicAdd(&d.iflags[4], ti);
    }
    __syncthreads();
    int base_full = s_base_full, base_rem = 0;
#pragma unroll
    for (int k = 0; k < PRUNE_WARPS; k++) base_full += k < w ? s_nfull[k] : 0;
    if (lane == 0 && rem > 0) base_rem = atomicAdd(&d.iflags[ITEM_BUCKET0 + rem], 1);
    for (int c = lane; c < nfull; c += 32)
        d.items[(size_t)ITEM_STEPS * d.max_items + base_full + c] =
            make_int4((int)(li.offset + (size_t)c * ITEM_STEPS * 32), A | (li.target << 28), r | (ITEM_STEPS << 8), c * ITEM_STEPS);
    if (lane == 0 && rem > 0)
        d.items[(size_t)rem * d.max_items + base_rem] =
            make_int4((int)(li.offset + (size_t)nfull * ITEM_STEPS * 32), A | (li.target << 28), r | (rem << 8), nfull * ITEM_STEPS);
#endif
}

}  // namespace atm
