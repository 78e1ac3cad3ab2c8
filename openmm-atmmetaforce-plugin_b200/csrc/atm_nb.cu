// atm_nb.cu -- Tier 2: the fused two-state direct-space NonbondedForce path for sm_100a.
//
// What it evaluates (DESIGN.md "Two-state direct space"): the reference runs two complete inner-context
// evaluations per step, U1,F1 at x and U2,F2 at x+d (openmmapi/src/ATMMetaForceImpl.cpp:113,116).  Only pairs
// that involve a displaced atom differ between the two states, so this back-end evaluates
//     C  : pairs whose two atoms move together (env-env, same displacement group)      -> both states
//     S1 : pairs between differently displaced atoms at the state-1 coordinates (x)     -> state 1 only
//     S2 : the same atom pairs at the state-2 coordinates (x+d), via "ghost" sites      -> state 2 only
// in ONE launch:  F1 = C + S1, F2 = C + S2, U2 - U1 = U(S2) - U(S1) (formed from the few thousand
// state-specific pairs only, so it does not suffer the cancellation of two 1e5 kJ/mol totals).
//
// Layout: "sites" = N atoms + M ghosts (displaced atoms at x+d).  Sites are binned by (class, xy column) -- class =
// environment, displaced atoms of group g, ghosts of group g -- z-sorted inside a bin and cut into clusters of 8
// (so clusters are class-pure and compact whatever the size of a displaced group).  Every cluster owns a list of
// individual partner sites inside cutoff+skin of its atoms.  The force kernel gives one warp a (cluster, list
// chunk): each LANE holds one partner site j, the 8 cluster atoms are broadcast from shared memory, so the inner loop has no
// shuffles and no shared-memory traffic; f_j goes out with three 64-bit fixed-point RED.ADDs per lane per
// 8 pairs, f_i is reduced by a 27-shuffle transpose-reduction once per work item.
#include <cub/device/device_radix_sort.cuh>
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "atm_common.cuh"

namespace atm {

constexpr int CL = 8;            // sites per cluster
#ifndef ATM_ITEM_STEPS
#define ATM_ITEM_STEPS 16
#endif
constexpr int ITEM_STEPS = ATM_ITEM_STEPS;     // 32-entry list steps per work item
constexpr int NB_THREADS = 128;    // force kernel block size (4 warps, one work item each)
#ifndef ATM_NB_MIN_BLOCKS
#define ATM_NB_MIN_BLOCKS 5
#endif
constexpr int NB_MIN_BLOCKS = ATM_NB_MIN_BLOCKS;   // 5 -> 20 warps / SM at <= 102 registers (cluster atoms live in shared memory)
constexpr int TGT_C = 0, TGT_S1 = 1, TGT_S2 = 2, TGT_SKIP = 3;
constexpr double ENERGY_SCALE = 4294967296.0;  // 2^32 fixed point for the energy accumulators
constexpr int ITEM_BUCKET0 = 16;  // flags[ITEM_BUCKET0 + n] = number of work items with n list steps (n = 1..ITEM_STEPS)
constexpr int NUM_FLAGS = 40;
#ifndef ATM_PRUNE_BLOCK
#define ATM_PRUNE_BLOCK 2          // list steps per software-pipeline block of the prune kernel (1 = one-step loop)
#endif
constexpr int EACC_SLOTS = 8;                  // Uc, U(S1), U(S2), pairs in cutoff per target (C, S1, S2), Urec(1), Urec(2)
constexpr double PME_SCALE = 1099511627776.0;  // 2^40 fixed point of the charge-grid accumulation

struct NbDev {  // everything the kernels need, passed by value
    int N, P, R, M, G, U;
    int nx, ny, ncol, nbins;
    int Smax, Cmax, CLmax, CXmax, CenvMax;
    int capC, capX;
    int max_items;  // capacity of one item bucket
    float cutoff2, rlist, rlist_outer, alpha, two_alpha_over_sqrtpi;
    // static, by atom
    const float *qp_atom;
    const float2 *par_atom;
    const int *excl_start, *excl_list;
    const int *group_of_atom, *ghost_atom, *ghost_of_atom, *slot_of_atom, *atom_of_slot;
    const float4 *displ;  // slot order (handle)
    const float4 *box, *invbox;  // [R]
    // per rebuild
    unsigned long long *keys;
    int *vals;
    int *bin_count, *bin_site_start, *bin_cluster_start, *nclusters;
    int *slot_site, *site_slot, *slot_src, *slot_out, *slot_ghost;
    float *slot_qp;
    float4 *xs;
    float2 *par;
    float4 *cc, *ch;
    int *cmeta;
    unsigned int *jlist, *jlist_outer;
    int *list_nsteps, *outer_nsteps;
    int4 *items;  // work items: {list offset in entries, cluster | target << 28, replica | steps << 8, first entry step}
    int *flags;
    // accumulators
    unsigned long long *buf;
    unsigned long long *eacc;
    double *energies;
    const double *params;
    // smooth PME reciprocal space (optional, atm_pme_setup)
    int pme_on, pme_order, gx, gy, gz;
    unsigned long long *pme_acc;   // [R][2][ng] fixed point: Q1 (environment + displaced atoms), Q2 - Q1 (ghosts - displaced)
    double *pme_grid;              // [R][2][ng] real charge grids of the two states (overwritten by the potentials)
    double2 *pme_spec;             // [R][2][gx][gy][gz/2+1]
    const double *pme_mod;         // |b(m)|^2 moduli: gx + gy + gz doubles
    double pme_self_sum;           // sum of (q sqrt(ke))^2 over all atoms
    double pme_qtot2;              // (sum of q sqrt(ke))^2: neutralising-background term -pi Q^2 / (2 V alpha^2)
    double disp_coeff;             // long-range dispersion correction = disp_coeff / V (0 = off)
};

struct NbState {
    bool ready = false, list_valid = false, groups_valid = false;
    atm_nonbonded_desc desc{};
    std::vector<float> h_qp;
    std::vector<float2> h_par;
    std::vector<int> h_excl_start, h_excl_list;
    std::vector<int2> h_excl_pairs, h_exc_pairs;
    std::vector<float4> h_exc_par;
    std::vector<int> h_group_of_atom, h_ghost_atom, h_ghost_of_atom;
    std::vector<double> h_box;  // [R][3]
    bool box_set = false, box_dirty = true;
    double disp_coeff_full = 0.0;  // 8 pi N^2 (<eps sig^12>/(9 rc^9) - <eps sig^6>/(3 rc^3)), applied when switched on
    bool disp_on = false;
    NbDev d{};
    // owned device memory (freed in nb_destroy)
    std::vector<void *> owned;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    unsigned long long *keys_alt = nullptr;
    int *vals_alt = nullptr;
    int2 *d_excl_pairs = nullptr, *d_exc_pairs = nullptr;
    float4 *d_exc_par = nullptr;
    int n_excl = 0, n_exc = 0;
    int sort_bits = 64;
    size_t jlist_entries = 0;
    int n_items = 0, max_items = 0, items_seen = 0;
    // PME (optional)
    std::vector<void *> pme_owned;
    cufftHandle pme_plan_fwd = 0, pme_plan_bwd = 0;
    bool pme_plans = false;
    uint64_t generation = 0;        // bumped by every rebuild
    uint64_t alloc_generation = 0;  // bumped by every (re)allocation: buffers and grid bounds change, graphs are stale
    bool verified = false, needs_realloc = false, flags_pending = false;
    int *h_flags = nullptr;         // pinned
    cudaEvent_t flags_event = nullptr;
    cudaGraphExec_t rebuild_graph = nullptr, prune_graph = nullptr;
    const void *rebuild_graph_posq = nullptr, *prune_graph_posq = nullptr;
    uint64_t rebuild_graph_generation = 0, prune_graph_generation = 0;
    int rebuild_graph_launches = 0, prune_graph_launches = 0;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    cudaGraphExec_t graph_exec = nullptr;
    atm_step_io graph_io{};
    uint64_t graph_generation = 0;
    int graph_nodes = 0;
    int64_t stats[8] = {0};
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int class_kind(int cls, int G) { return cls == 0 ? 0 : (cls <= G ? 1 : 2); }
__device__ __forceinline__ int class_group(int cls, int G) { return cls == 0 ? 0 : (cls <= G ? cls : cls - G); }

__device__ __forceinline__ int pair_target(int ca, int cb, int G) {
    const int ka = class_kind(ca, G), kb = class_kind(cb, G);
    if (ka == 0 && kb == 0) return TGT_C;
    if ((ka == 1 && kb == 2) || (ka == 2 && kb == 1)) return TGT_SKIP;
    const bool same = class_group(ca, G) == class_group(cb, G);
    if (ka == 2 || kb == 2) return (ka == 2 && kb == 2 && same) ? TGT_SKIP : TGT_S2;
    if (ka == 1 && kb == 1) return same ? TGT_C : TGT_S1;
    return TGT_S1;
}

__device__ __forceinline__ float wrap_delta(float d, float L, float invL) {
    // round-to-nearest through the 1.5*2^23 trick (|d/L| < 2^22): avoids the quarter-rate FRND instruction
    return d - L * __fadd_rn(__fadd_rn(d * invL, 12582912.0f), -12582912.0f);
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor drains; it must not touch the predecessor's output before pdl_wait() returns (which
// waits for the predecessor grid to complete and its writes to be visible).  pdl_trigger() in the predecessor lets the
// dependent grid be scheduled as soon as every predecessor block has started.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void red_add_fixed(unsigned long long *addr, float f) {
    long long v = __float2ll_rn(f * 4294967296.0f);
    atomicAdd(addr, (unsigned long long)v);
}

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

// Rebuild step 2: per replica exclusive scans of the bin populations (sites and 8-padded clusters).
__global__ void nl_scan_kernel(NbDev d) {
    const int r = blockIdx.x;
    __shared__ int s_sites[1024], s_clusters[1024];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (d.nbins + nt - 1) / nt;
    const int b0 = min(tid * per, d.nbins), b1 = min(b0 + per, d.nbins);
    int ns = 0, nc = 0;
    for (int b = b0; b < b1; b++) {
        int c = d.bin_count[r * d.nbins + b];
        ns += c;
        nc += (c + CL - 1) / CL;
    }
    s_sites[tid] = ns;
    s_clusters[tid] = nc;
    __syncthreads();
    if (tid == 0) {
        int as = 0, ac = 0;
        for (int i = 0; i < nt; i++) {
            int ts = s_sites[i], tc = s_clusters[i];
            s_sites[i] = as;
            s_clusters[i] = ac;
            as += ts;
            ac += tc;
        }
        d.nclusters[r] = ac;
        d.bin_site_start[r * (d.nbins + 1) + d.nbins] = as;
        d.bin_cluster_start[r * (d.nbins + 1) + d.nbins] = ac;
    }
    __syncthreads();
    ns = s_sites[tid];
    nc = s_clusters[tid];
    for (int b = b0; b < b1; b++) {
        int c = d.bin_count[r * d.nbins + b];
        d.bin_site_start[r * (d.nbins + 1) + b] = ns;
        d.bin_cluster_start[r * (d.nbins + 1) + b] = nc;
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
__global__ void nb_pack_kernel(NbDev d, const float4 *__restrict__ posq) {
    pdl_trigger();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    // first kernel of a step: the energy / pair-count accumulators of the previous step have been consumed by its merge
    if (blockIdx.x == 0 && threadIdx.x < EACC_SLOTS) d.eacc[(size_t)r * EACC_SLOTS + threadIdx.x] = 0ull;
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
    // the per-step image shift of a partner relative to the cluster centre needs half extent + list radius <= L/2
    const float hmax_x = 0.5f * (hi.x - lo.x) + d.rlist_outer, hmax_y = 0.5f * (hi.y - lo.y) + d.rlist_outer,
                hmax_z = 0.5f * (hi.z - lo.z) + d.rlist_outer;
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
    __shared__ int s_pass[BUILD_WARPS][32];
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
        for (int base = c_begin; base < c_end; base += 32) {
            // stage 1: one candidate cluster per lane, box-box distance; passing clusters compacted (in order) to smem
            const int B = base + lane;
            bool pass = false;
            if (B < c_end) {
                const size_t rcB = (size_t)r * d.Cmax + B;
                const int clsB = d.cmeta[rcB] & 0xffff;
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
            if (pass) s_pass[w][__popc(cmask & ((1u << lane) - 1))] = B;
            __syncwarp();
            // stage 2: one candidate SITE per lane (four clusters per sweep)
            for (int q = 0; q < npass; q += 4) {
                const int g = q + (lane >> 3), k = lane & 7;
                bool take = false;
                unsigned int entry = 0;
                if (g < npass) {
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
            }
            __syncwarp();
        }
    }
    if (count > li.cap) {
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
        sa[lane] = make_float4(-2.f * ax, -2.f * ay, -2.f * az, ok ? fmaf(az, az, fmaf(ay, ay, ax * ax)) : 1e30f);
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
                float d2min = 1e30f;
#pragma unroll
                for (int k = 0; k < CL; k++) { const float4 a = sa[k]; d2min = fminf(d2min, fmaf(px, a.x, fmaf(py, a.y, fmaf(pz, a.z, a.w)))); }
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
    __shared__ float4 s_atoms[PRUNE_WARPS][CL];
    const int count = prune_one_list(d, r, l, d.Cmax + d.CLmax, lane, li, s_atoms[w]);
    const int A = li.cluster;
    const int nsteps = (count + 31) >> 5;
    // work items of this list: (<= ITEM_STEPS)-step chunks (their order only affects scheduling: every accumulation
    // downstream is fixed point, hence order independent).  Buckets by chunk length: the force kernel hands out the
    // longest chunks first (longest-processing-time order keeps the tail of the launch short when only a few replicas
    // share the GPU).  The counters every list touches -- kept entries, live items, the bucket of full chunks -- are
    // summed over the block first: same-address atomics serialise in the L2, one per block instead of one per list.
    const int nfull = nsteps / ITEM_STEPS, rem = nsteps - nfull * ITEM_STEPS;
    __shared__ int s_count[PRUNE_WARPS], s_nfull[PRUNE_WARPS], s_items[PRUNE_WARPS], s_base_full;
    if (lane == 0) { s_count[w] = count; s_nfull[w] = nfull; s_items[w] = nfull + (rem > 0 ? 1 : 0); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tc = 0, tf = 0, ti = 0;
#pragma unroll
        for (int k = 0; k < PRUNE_WARPS; k++) { tc += s_count[k]; tf += s_nfull[k]; ti += s_items[k]; }
        if (tc > 0) atomicAdd((unsigned long long *)&d.flags[6], (unsigned long long)tc);
        s_base_full = tf > 0 ? atomicAdd(&d.flags[ITEM_BUCKET0 + ITEM_STEPS], tf) : 0;
        if (ti > 0) atomicAdd(&d.flags[4], ti);
    }
    __syncthreads();
    int base_full = s_base_full, base_rem = 0;
#pragma unroll
    for (int k = 0; k < PRUNE_WARPS; k++) base_full += k < w ? s_nfull[k] : 0;
    if (lane == 0 && rem > 0) base_rem = atomicAdd(&d.flags[ITEM_BUCKET0 + rem], 1);
    for (int c = lane; c < nfull; c += 32)
        d.items[(size_t)ITEM_STEPS * d.max_items + base_full + c] =
            make_int4((int)(li.offset + (size_t)c * ITEM_STEPS * 32), A | (li.target << 28), r | (ITEM_STEPS << 8), c * ITEM_STEPS);
    if (lane == 0 && rem > 0)
        d.items[(size_t)rem * d.max_items + base_rem] =
            make_int4((int)(li.offset + (size_t)nfull * ITEM_STEPS * 32), A | (li.target << 28), r | (rem << 8), nfull * ITEM_STEPS);
}

// ------------------------------------------------------------------------------------------------
// The force kernel.
// ------------------------------------------------------------------------------------------------
// erfc(x) = t (a0 + a1 t + ... + a6 t^6) exp(-x^2), t = 1/(1 + p x); max relative error 6.8e-8 on [0, 4.2]
// (fit: DESIGN.md "erfc"); evaluated in fp32 the rounding error (~4e-7) dominates.
#define ERFC_P 0.357431514f
#define ERFC_A0 1.9957954259e-01f
#define ERFC_A1 2.2717719300e-01f
#define ERFC_A2 5.6467497521e-02f
#define ERFC_A3 5.3674605718e-01f
#define ERFC_A4 -4.8585730230e-01f
#define ERFC_A5 6.4534500206e-01f
#define ERFC_A6 -1.7945805777e-01f

// MUFU wrappers without the denormal / range fix-up code the CUDA math library adds around them
__device__ __forceinline__ float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// round to nearest integer for |x| < 2^22 without the (quarter-rate) FRND instruction
__device__ __forceinline__ float fast_rint(float x) { return __fadd_rn(__fadd_rn(x, 12582912.0f), -12582912.0f); }

struct PairConst {
    float cutoff2;
    float p_alpha;      // ERFC_P * alpha
    float neg_a2_log2e; // -alpha^2 log2(e)
    float two_a_sqrtpi; // 2 alpha / sqrt(pi)
};

// One pair.  qq = q_i q_j k_e (charges are stored pre-multiplied by sqrt(k_e)); sig = (s_i+s_j)/2; eps4 = 4 sqrt(e_i e_j).
// Returns F/r (fscale) and, when ENERGY, the pair energy.
template <bool ENERGY>
__device__ __forceinline__ float pair_interaction(float r2, float qq, float sig, float eps4, const PairConst &pc, float &energy) {
    float rinv = mufu_rsqrt(r2);
#ifdef ATM_RSQRT_NEWTON  // A/B switch: measured force/energy parity is identical without the Newton step
    rinv = rinv * fmaf(-0.5f * r2, rinv * rinv, 1.5f);
#endif
    const float rinv2 = rinv * rinv;
    const float r = r2 * rinv;
    const float s2 = sig * sig * rinv2;
    const float s6 = s2 * s2 * s2;
    const float es6 = eps4 * s6;
    const float flj = es6 * fmaf(12.0f, s6, -6.0f);
    const float t = mufu_rcp(fmaf(pc.p_alpha, r, 1.0f));
    const float ex = mufu_ex2(pc.neg_a2_log2e * r2);
    float poly = fmaf(ERFC_A6, t, ERFC_A5);
    poly = fmaf(poly, t, ERFC_A4);
    poly = fmaf(poly, t, ERFC_A3);
    poly = fmaf(poly, t, ERFC_A2);
    poly = fmaf(poly, t, ERFC_A1);
    poly = fmaf(poly, t, ERFC_A0);
    // Coulomb: E = qq exp(-a^2 r^2) P(t) t / r,  F r = E + qq (2a/sqrt(pi)) exp(-a^2 r^2)  -- five instructions
    const float g = poly * t * rinv;
    const float qe = qq * ex;
    const float ec = qe * g;
    const float fc = fmaf(qe, pc.two_a_sqrtpi, ec);
    if (ENERGY) energy = fmaf(es6, s6, -es6) + ec;
    return (flj + fc) * rinv2;
}

struct ItemCtx {
    int r, A, target, nst;
    const unsigned int *list;
    size_t rsite, comp_stride;
};

template <bool ENERGY, bool STATS>
__device__ __forceinline__ void nb2_item_epilogue(const NbDev &d, const ItemCtx &it, int lane, unsigned long long *buf,
                                                  float (&fix)[CL], float (&fiy)[CL], float (&fiz)[CL], double e_acc, int npairs) {
    // transpose-reduce the 24 i-force accumulators: after three halving exchanges lane (l&7) owns atom l&7
    {
        const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
        float w[4][3];
#pragma unroll
        for (int mm = 0; mm < 4; mm++) {
            const float kx = b0 ? fix[2 * mm + 1] : fix[2 * mm], sx = b0 ? fix[2 * mm] : fix[2 * mm + 1];
            const float ky = b0 ? fiy[2 * mm + 1] : fiy[2 * mm], sy = b0 ? fiy[2 * mm] : fiy[2 * mm + 1];
            const float kz = b0 ? fiz[2 * mm + 1] : fiz[2 * mm], sz = b0 ? fiz[2 * mm] : fiz[2 * mm + 1];
            w[mm][0] = kx + __shfl_xor_sync(0xffffffffu, sx, 1);
            w[mm][1] = ky + __shfl_xor_sync(0xffffffffu, sy, 1);
            w[mm][2] = kz + __shfl_xor_sync(0xffffffffu, sz, 1);
        }
        float x2[2][3];
#pragma unroll
        for (int mm = 0; mm < 2; mm++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float kk = b1 ? w[2 * mm + 1][c] : w[2 * mm][c], ss = b1 ? w[2 * mm][c] : w[2 * mm + 1][c];
                x2[mm][c] = kk + __shfl_xor_sync(0xffffffffu, ss, 2);
            }
        float y[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float kk = b2 ? x2[1][c] : x2[0][c], ss = b2 ? x2[0][c] : x2[1][c];
            y[c] = kk + __shfl_xor_sync(0xffffffffu, ss, 4);
            y[c] += __shfl_xor_sync(0xffffffffu, y[c], 8);
            y[c] += __shfl_xor_sync(0xffffffffu, y[c], 16);
        }
        // lane l (< 8) now holds atom index (b0 + 2 b1 + 4 b2) = l
        if (lane < CL) {
            const int i = it.A * CL + lane;
            red_add_fixed(buf + i, y[0]);
            red_add_fixed(buf + it.comp_stride + i, y[1]);
            red_add_fixed(buf + 2 * it.comp_stride + i, y[2]);
        }
    }
    // energies: warp sum in double, one fixed-point atomic per warp
    if (ENERGY || STATS) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            if (ENERGY) e_acc += __shfl_xor_sync(0xffffffffu, e_acc, off);
            if (STATS) npairs += __shfl_xor_sync(0xffffffffu, npairs, off);
        }
        if (lane == 0) {
            unsigned long long *ea = d.eacc + (size_t)it.r * EACC_SLOTS;
            if (ENERGY) atomicAdd(ea + it.target, (unsigned long long)__double2ll_rn(e_acc * ENERGY_SCALE));
            if (STATS) atomicAdd(ea + 3 + it.target, (unsigned long long)npairs);
        }
    }
}

// Partner data is staged through a per-lane shared-memory ring with cp.async (prefetch distance 3 steps); the cluster
// atoms are broadcast from shared memory instead of living in 48 registers, which buys a fifth resident block per SM.
constexpr int NB_WARPS = NB_THREADS / 32;
#ifndef ATM_PF_DIST
#define ATM_PF_DIST 3
#endif
#ifndef ATM_RING
#define ATM_RING 4
#endif
constexpr int RING = ATM_RING;        // ring slots per lane (power of two, > PF_DIST)
constexpr int PF_DIST = ATM_PF_DIST;  // coordinate steps in flight

// The list entries of a work item (<= ITEM_STEPS x 128 B, contiguous, streaming from DRAM) are brought into shared
// memory by ONE bulk asynchronous copy (the TMA unit: cp.async.bulk, SASS UBLKCP) that signals an mbarrier, instead of
// a rolling register prefetch of one LDG per step.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_arm(unsigned long long *bar, void *smem, const void *gmem, unsigned int bytes) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct __align__(128) Nb2Smem {
    unsigned int el[NB_WARPS][ITEM_STEPS][32];  // the item's list entries (bulk copy destination)
    float4 xj[NB_WARPS][RING][32];
    float2 pj[NB_WARPS][RING][32];
    float4 xi[NB_WARPS][CL];
    float2 pi[NB_WARPS][CL];
    unsigned long long bar[NB_WARPS];           // one mbarrier per warp
};

template <bool ENERGY, bool STATS>
__device__ __forceinline__ void nb2_item(const NbDev &d, const ItemCtx &it, int lane, int w, Nb2Smem &sm, unsigned int parity) {
    const float4 L = d.box[it.r], iL = d.invbox[it.r];
    const float4 cA = d.cc[(size_t)it.r * d.Cmax + it.A];
    PairConst pc;
    pc.cutoff2 = d.cutoff2;
    pc.p_alpha = ERFC_P * d.alpha;
    pc.neg_a2_log2e = -d.alpha * d.alpha * 1.4426950408889634f;
    pc.two_a_sqrtpi = d.two_alpha_over_sqrtpi;

    // the whole entry list of the item: one bulk copy, in flight while the cluster atoms are staged
    if (lane == 0) bulk_load_arm(&sm.bar[w], &sm.el[w][0][0], it.list, (unsigned)it.nst * 128u);
    // cluster atoms -> shared memory (lanes 0..7), shifted next to the cluster centre
    if (lane < CL) {
        float4 x = __ldg(d.xs + it.rsite + (size_t)it.A * CL + lane);
        x.x -= L.x * fast_rint((x.x - cA.x) * iL.x);
        x.y -= L.y * fast_rint((x.y - cA.y) * iL.y);
        x.z -= L.z * fast_rint((x.z - cA.z) * iL.z);
        sm.xi[w][lane] = x;
        sm.pi[w][lane] = __ldg(d.par + it.rsite + (size_t)it.A * CL + lane);
    }
    mbar_wait(&sm.bar[w], parity);
#pragma unroll
    for (int q = 0; q < PF_DIST; q++) {
        if (q < it.nst) {
            const unsigned int eq = sm.el[w][q][lane];
            cp_async16(&sm.xj[w][q][lane], d.xs + it.rsite + (eq >> 8));
            cp_async8(&sm.pj[w][q][lane], d.par + it.rsite + (eq >> 8));
        }
        cp_async_commit();
    }
    __syncwarp();

    float fix[CL], fiy[CL], fiz[CL];
#pragma unroll
    for (int k = 0; k < CL; k++) fix[k] = fiy[k] = fiz[k] = 0.f;
    unsigned long long *buf = d.buf + (size_t)it.target * 3 * it.comp_stride + it.rsite;
    double e_acc = 0.0;
    int npairs = 0;

    for (int st = 0; st < it.nst; st++) {
        cp_async_wait<PF_DIST - 1>();  // the group of step st has landed (groups retire in order)
        const int slot = st & (RING - 1);
        const float4 xjc = sm.xj[w][slot][lane];
        const float2 pjc = sm.pj[w][slot][lane];
        const unsigned int e = sm.el[w][st][lane];
        // keep PF_DIST steps in flight
        {
            const int sp = st + PF_DIST;
            if (sp < it.nst) {
                const unsigned int e_next = sm.el[w][sp][lane];
                const int ps = sp & (RING - 1);
                cp_async16(&sm.xj[w][ps][lane], d.xs + it.rsite + (e_next >> 8));
                cp_async8(&sm.pj[w][ps][lane], d.par + it.rsite + (e_next >> 8));
            }
            cp_async_commit();
        }
        const int j = e >> 8;
        const unsigned int m = e & 0xffu;
        const float xjx = xjc.x - L.x * fast_rint((xjc.x - cA.x) * iL.x);
        const float xjy = xjc.y - L.y * fast_rint((xjc.y - cA.y) * iL.y);
        const float xjz = xjc.z - L.z * fast_rint((xjc.z - cA.z) * iL.z);
        float fjx = 0.f, fjy = 0.f, fjz = 0.f, e_step = 0.f;
        bool any = false;
#pragma unroll
        for (int k = 0; k < CL; k++) {
            const float4 xi = sm.xi[w][k];
            const float2 pi = sm.pi[w][k];
            const float dx = xi.x - xjx, dy = xi.y - xjy, dz = xi.z - xjz;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const bool in = (r2 < pc.cutoff2) && !(m & (1u << k));
            float en = 0.f;
            // a pair outside the cutoff (or excluded, or a padding slot) is evaluated at r^2 = 1e30: every term underflows
            // to exactly zero (flush-to-zero MUFU paths, no inf/NaN even for r = 0), so ONE select on r^2 replaces the
            // selects on the force scale and on the energy
            const float fs = pair_interaction<ENERGY>(in ? r2 : 1e30f, xi.w * xjc.w, pi.x + pjc.x, pi.y * pjc.y, pc, en);
            if (ENERGY) e_step += en;
            if (STATS) npairs += in ? 1 : 0;
            any |= in;
            fix[k] = fmaf(dx, fs, fix[k]); fiy[k] = fmaf(dy, fs, fiy[k]); fiz[k] = fmaf(dz, fs, fiz[k]);
            fjx = fmaf(-dx, fs, fjx); fjy = fmaf(-dy, fs, fjy); fjz = fmaf(-dz, fs, fjz);
        }
        if (ENERGY) e_acc += (double)e_step;
        if (any) {  // (deriving this from fj != 0 after the loop instead of 8 predicate ORs measured 2 % SLOWER per launch)
            red_add_fixed(buf + j, fjx);
            red_add_fixed(buf + it.comp_stride + j, fjy);
            red_add_fixed(buf + 2 * it.comp_stride + j, fjz);
        }
    }
    cp_async_wait<0>();
    nb2_item_epilogue<ENERGY, STATS>(d, it, lane, buf, fix, fiy, fiz, e_acc, npairs);
}

// ------------------------------------------------------------------------------------------------
// Excluded pairs (Ewald correction -qq erf(ar)/r, minimum image) and 1-4 exceptions (plain Coulomb + LJ, no image).
// One thread per (replica, pair).  Pairs whose atoms move together go to C, others are evaluated in both states.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_pair_force(const NbDev &d, int r, int target, int si, int sj, float fx, float fy, float fz) {
    unsigned long long *buf = d.buf + (size_t)target * 3 * d.R * d.Smax + (size_t)r * d.Smax;
    const size_t cs = (size_t)d.R * d.Smax;
    red_add_fixed(buf + si, fx); red_add_fixed(buf + cs + si, fy); red_add_fixed(buf + 2 * cs + si, fz);
    red_add_fixed(buf + sj, -fx); red_add_fixed(buf + cs + sj, -fy); red_add_fixed(buf + 2 * cs + sj, -fz);
}

__device__ __forceinline__ void special_pairs_body(const NbDev &d, int t, int r, const int2 *__restrict__ excl, int n_excl,
                                                   const int2 *__restrict__ exc, const float4 *__restrict__ exc_par, int n_exc) {
    double e_tgt[3] = {0.0, 0.0, 0.0};
    if (t < n_excl + n_exc) {
        const bool is_exc = t >= n_excl;
        const int2 pr = is_exc ? exc[t - n_excl] : excl[t];
        const int ga = d.group_of_atom[pr.x], gb = d.group_of_atom[pr.y];
        const float4 L = d.box[r], iL = d.invbox[r];
        const int nstate = (ga == gb) ? 1 : 2;
        for (int s = 0; s < nstate; s++) {
            const int target = (ga == gb) ? TGT_C : (s == 0 ? TGT_S1 : TGT_S2);
            int si = d.site_slot[(size_t)r * d.U + pr.x], sj = d.site_slot[(size_t)r * d.U + pr.y];
            if (s == 1) {
                if (ga != 0) si = d.site_slot[(size_t)r * d.U + d.N + d.ghost_of_atom[pr.x]];
                if (gb != 0) sj = d.site_slot[(size_t)r * d.U + d.N + d.ghost_of_atom[pr.y]];
            }
            const float4 a = d.xs[(size_t)r * d.Smax + si], b = d.xs[(size_t)r * d.Smax + sj];
            float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
            float fs, en;
            if (!is_exc) {
                dx = wrap_delta(dx, L.x, iL.x); dy = wrap_delta(dy, L.y, iL.y); dz = wrap_delta(dz, L.z, iL.z);
                const float r2 = dx * dx + dy * dy + dz * dz;
                const float qq = a.w * b.w;
                if (r2 > 0.f && qq != 0.f && d.alpha > 0.f) {
                    const float rinv = 1.0f / sqrtf(r2), rr = r2 * rinv, ar = d.alpha * rr;
                    const float erf_ar = erff(ar);
                    en = -qq * rinv * erf_ar;
                    fs = -qq * rinv * (erf_ar - ar * expf(-ar * ar) * 1.1283791670955126f) * rinv * rinv;
                } else { en = 0.f; fs = 0.f; }
            } else {
                const float4 pp = exc_par[t - n_excl];  // ke*chargeProd, sigma, 4 eps
                const float r2 = dx * dx + dy * dy + dz * dz;
                const float rinv = 1.0f / sqrtf(r2), rinv2 = rinv * rinv;
                const float s2 = pp.y * pp.y * rinv2, s6 = s2 * s2 * s2;
                en = pp.z * s6 * (s6 - 1.0f) + pp.x * rinv;
                fs = (pp.z * s6 * (12.0f * s6 - 6.0f) + pp.x * rinv) * rinv2;
            }
            add_pair_force(d, r, target, si, sj, dx * fs, dy * fs, dz * fs);
            e_tgt[target] += (double)en;
        }
    }
    // block reduction of the three energies (warp shuffle, then one atomic per warp)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double v = e_tgt[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0 && v != 0.0) {
            unsigned long long *ea = d.eacc + (size_t)r * EACC_SLOTS;
            atomicAdd(ea + k, (unsigned long long)__double2ll_rn(v * ENERGY_SCALE));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused scalar stage + merge.  Per replica: u = U(S2) - U(S1), soft-core, softplus, sp (double, once per block),
// then  F[slot] += C + (1-sp) S1 + sp S2  gathered from cluster order, accumulators zeroed behind the read.
// Semantics: CommonATMMetaForceKernels.cpp:182-201 + kernels/atmmetaforce.cc:8-16, blend in double.
// ------------------------------------------------------------------------------------------------
constexpr int MERGE2_THREADS = 256;

// Scalar stage, one thread per replica: u = U(S2) - U(S1), soft-core, softplus, sp -- all in double on the device
// (the reference does this on the host after two blocking energy downloads, CommonATMMetaForceKernels.cpp:164-199).
__device__ double scalar_stage_replica(const NbDev &d, int r, const double *__restrict__ energy_ext, int include_energy,
                                       bool write_record) {
    volatile unsigned long long *ea = d.eacc + (size_t)r * EACC_SLOTS;  // written by atomics of other blocks: read through L2
    const double uc = (double)(long long)ea[0] / ENERGY_SCALE, u1 = (double)(long long)ea[1] / ENERGY_SCALE,
                 u2 = (double)(long long)ea[2] / ENERGY_SCALE;
    double U1 = uc + u1, U2 = uc + u2, du = u2 - u1;
    double *e = d.energies + (size_t)r * ATM_NUM_ENERGY_SLOTS;
    double rec1 = 0.0, rec2 = 0.0, eself = 0.0;
    if (d.pme_on) {
        // reciprocal energies of the two states (accumulated in double, so their difference is as good as the sum)
        const double r1 = (double)(long long)ea[6] / ENERGY_SCALE, r2 = (double)(long long)ea[7] / ENERGY_SCALE;
        const double self = -d.pme_self_sum * (double)d.alpha * 0.5641895835477563;  // -alpha/sqrt(pi) sum q^2
        const float4 Lb = d.box[r];
        const double bg = -3.141592653589793 * d.pme_qtot2 /
                          (2.0 * (double)Lb.x * (double)Lb.y * (double)Lb.z * (double)d.alpha * (double)d.alpha);
        U1 += r1 + self + bg;
        U2 += r2 + self + bg;
        du += r2 - r1;
        rec1 = r1; rec2 = r2; eself = self;
    }
    if (d.disp_coeff != 0.0) {  // same constant in both states: u is unaffected
        const float4 Lb = d.box[r];
        const double ed = d.disp_coeff / ((double)Lb.x * (double)Lb.y * (double)Lb.z);
        U1 += ed;
        U2 += ed;
    }
    if (energy_ext) {
        U1 += energy_ext[2 * r];
        U2 += energy_ext[2 * r + 1];
        du += energy_ext[2 * r + 1] - energy_ext[2 * r];
    }
    const Scalars s = scalar_stage(d.params + (size_t)r * ATM_NUM_PARAMS, U1, U2, du);
    if (write_record) {
        e[ATM_E_UREC1] = rec1; e[ATM_E_UREC2] = rec2; e[ATM_E_USELF] = eself;
        e[ATM_E_U1] = U1; e[ATM_E_U2] = U2; e[ATM_E_U] = s.u; e[ATM_E_USC] = s.usc; e[ATM_E_EBIAS] = s.ebias;
        e[ATM_E_ENERGY] = include_energy ? s.energy : 0.0; e[ATM_E_SP] = s.sp;
        e[ATM_E_NPAIRS] = (double)(ea[3] + ea[4] + ea[5]);
        e[ATM_E_NPAIRS_C] = (double)ea[3]; e[ATM_E_NPAIRS_S1] = (double)ea[4]; e[ATM_E_NPAIRS_S2] = (double)ea[5];
    }
    return s.sp;  // the accumulators are zeroed by the next step's pack kernel
}

// ------------------------------------------------------------------------------------------------
// The per-step compute launch: work items of the pair lists (one warp each), then blocks of excluded / exception
// pairs.
// ------------------------------------------------------------------------------------------------
struct SpecialArgs {
    const int2 *excl;
    const int2 *exc;
    const float4 *exc_par;
    int n_excl, n_exc;
    int blocks_per_replica;  // ceil((n_excl + n_exc) / NB_THREADS)
};

template <bool STATS>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS)
nb2_kernel(NbDev d, int n_item_blocks, int energy_common, SpecialArgs sp) {
    pdl_trigger();  // the merge may be scheduled once every block of this grid has started (it waits for completion)
    pdl_wait();     // cluster-order coordinates come from the pack kernel
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x < n_item_blocks) {
        // The pruned list's item count lives on the device; the grid is sized from the count the last verified build
        // saw plus a margin, and this grid-stride loop picks up whatever a later prune added beyond it.
        __shared__ Nb2Smem sm;
        const int w = threadIdx.x >> 5;
        if (lane == 0) mbar_init(&sm.bar[w], 1);
        __syncwarp();
        unsigned int parity = 0;  // phase of this warp's mbarrier: flips with every completed bulk copy
        const int n_items = d.flags[4], stride = n_item_blocks * (NB_THREADS / 32);
        for (int warp = blockIdx.x * (NB_THREADS / 32) + (threadIdx.x >> 5); warp < n_items; warp += stride, parity ^= 1u) {
            __syncwarp();  // the previous item's readers of this warp's shared-memory slots are done
            int wi = warp, b = ITEM_STEPS;
            for (; b > 1; --b) {  // longest chunks first
                const int c = d.flags[ITEM_BUCKET0 + b];
                if (wi < c) break;
                wi -= c;
            }
            const int4 item = __ldg(d.items + (size_t)b * d.max_items + wi);
            ItemCtx it;
            it.r = item.z & 0xff;
            it.A = item.y & 0x0fffffff;
            it.target = (item.y >> 28) & 3;
            it.nst = item.z >> 8;
            it.list = d.jlist + (unsigned int)item.x;
            it.rsite = (size_t)it.r * d.Smax;
            it.comp_stride = (size_t)d.R * d.Smax;
            if (it.target == TGT_C && !energy_common) nb2_item<false, STATS>(d, it, lane, w, sm, parity);
            else nb2_item<true, STATS>(d, it, lane, w, sm, parity);
        }
    } else {
        const int sb = blockIdx.x - n_item_blocks;
        const int r = sb / sp.blocks_per_replica, chunk = sb - r * sp.blocks_per_replica;
        special_pairs_body(d, chunk * NB_THREADS + threadIdx.x, r, sp.excl, sp.n_excl, sp.exc, sp.exc_par, sp.n_exc);
    }
}

// Merge, one thread per cluster-order slot: the three accumulators are read (and zeroed) coalesced, only the final
// read-modify-write of the caller's force buffer is a scatter.  The thread of a displaced atom also folds in (and
// zeroes) the S2 accumulator of its ghost site; ghost and padding slots have no thread work.
// F[slot] += C + llrint(sp * S2 + (1 - sp) * S1), blend in double (kernels/atmmetaforce.cc:8-16 semantics).
// The scalar stage is fused in: thread 0 of every block forms u, the soft core, the softplus bias and sp = dW/du of the
// block's replica in double (the reference does this on the host after two blocking energy downloads,
// CommonATMMetaForceKernels.cpp:164-199); every block gets the same bits, block 0 of a replica writes the energy record.
__global__ void __launch_bounds__(MERGE2_THREADS)
nb_merge_kernel(NbDev d, long long *__restrict__ force, const long long *__restrict__ f1_ext,
                const long long *__restrict__ f2_ext, const double *__restrict__ energy_ext, int include_energy) {
    __shared__ double s_sp;
    const int r = blockIdx.y;
    const int s = blockIdx.x * MERGE2_THREADS + threadIdx.x;
    const size_t rsite = (size_t)r * d.Smax;
    // everything that does not depend on the force kernel first
    const bool live = s < CL * d.nclusters[r];
    const int i = live ? d.slot_out[rsite + s] : -1;  // caller's slot of this site's atom, -1 for ghosts and padding
    const int gs = i >= 0 ? d.slot_ghost[rsite + s] : -1;
    pdl_wait();
    if (threadIdx.x == 0) s_sp = scalar_stage_replica(d, r, energy_ext, include_energy, blockIdx.x == 0);
    __syncthreads();
    if (i < 0) return;
    const double sp = s_sp, sp1 = 1.0 - sp;
    const size_t cs = (size_t)d.R * d.Smax;
    long long *bufC = (long long *)d.buf + rsite, *buf1 = bufC + 3 * cs, *buf2 = bufC + 6 * cs;
    long long fc[3], fa[3], fb[3], fg[3], fo_old[3];
    size_t fo[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {  // every load first
        fc[c] = bufC[c * cs + s];
        fa[c] = buf1[c * cs + s];
        fb[c] = buf2[c * cs + s];
        fg[c] = gs >= 0 ? buf2[c * cs + gs] : 0;
        fo[c] = (size_t)r * 3 * d.P + (size_t)c * d.P + i;
        fo_old[c] = force[fo[c]];
    }
    bool nz1[3], nz2[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        nz1[c] = fa[c] != 0;
        nz2[c] = fb[c] != 0;
        if (f1_ext) fa[c] += f1_ext[fo[c]];
        if (f2_ext) fb[c] += f2_ext[fo[c]];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // hand the accumulators back zeroed; the state-specific ones are zero already for every site that is not
        // within the cutoff of a displaced atom or a ghost (96 % of them): skip those writes
        bufC[c * cs + s] = 0;
        if (nz1[c]) buf1[c * cs + s] = 0;
        if (nz2[c]) buf2[c * cs + s] = 0;
        if (gs >= 0) buf2[c * cs + gs] = 0;
        const double v = __dadd_rn(__dmul_rn(sp, (double)(fb[c] + fg[c])), __dmul_rn(sp1, (double)fa[c]));
        force[fo[c]] = fo_old[c] + fc[c] + __double2ll_rn(v);
    }
}

// ------------------------------------------------------------------------------------------------
// Two-state smooth PME reciprocal space.  Works on the cluster-order sites in double:
// the environment is spread ONCE; the displaced atoms (state 1) and their ghosts (state 2) are spread into two small
// extra accumulators; Q1 = env + lig, Q2 = env + ghost.  Two batched FFT pairs give the two potentials; environment
// sites gather from both (their reciprocal force differs between the states because the ligand's field moved).
// Charges carry sqrt(k_e), so the influence function needs no Coulomb constant.  Grid accumulation is 2^40 fixed
// point (deterministic), everything after it double precision: U2 - U1 keeps its digits.
// ------------------------------------------------------------------------------------------------
constexpr int PME_MAX_ORDER = 8;

// cardinal B-spline weights theta[k] and derivatives dtheta[k], k = 0..ORDER-1, for fractional offset w (Essmann 1995);
// ORDER is a compile-time constant so that everything stays in registers
template <int ORDER>
__device__ __forceinline__ void pme_bspline(double w, double (&theta)[ORDER], double (&dtheta)[ORDER]) {
#pragma unroll
    for (int k = 0; k < ORDER; k++) theta[k] = 0.0;
    theta[1] = w;
    theta[0] = 1.0 - w;
#pragma unroll
    for (int k = 3; k < ORDER; k++) {
        const double div = 1.0 / (k - 1.0);
        theta[k - 1] = div * w * theta[k - 2];
#pragma unroll
        for (int j = 1; j <= k - 2; j++) theta[k - j - 1] = div * ((w + j) * theta[k - j - 2] + (k - j - w) * theta[k - j - 1]);
        theta[0] = div * (1.0 - w) * theta[0];
    }
    dtheta[0] = -theta[0];
#pragma unroll
    for (int k = 1; k < ORDER; k++) dtheta[k] = theta[k - 1] - theta[k];
    const double div = 1.0 / (ORDER - 1.0);
    theta[ORDER - 1] = div * w * theta[ORDER - 2];
#pragma unroll
    for (int j = 1; j <= ORDER - 2; j++)
        theta[ORDER - j - 1] = div * ((w + j) * theta[ORDER - j - 2] + (ORDER - j - w) * theta[ORDER - j - 1]);
    theta[0] = div * (1.0 - w) * theta[0];
}

template <int ORDER>
struct PmeSite {
    int k0[3];
    double th[3][ORDER], dth[3][ORDER];
};

template <int ORDER>
__device__ __forceinline__ void pme_site_setup(const NbDev &d, const float4 &x, const float4 &L, PmeSite<ORDER> &ps) {
    const int n[3] = {d.gx, d.gy, d.gz};
    const double xr[3] = {(double)x.x / (double)L.x, (double)x.y / (double)L.y, (double)x.z / (double)L.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double u = (xr[c] - floor(xr[c])) * n[c];
        int fl = (int)floor(u);
        if (fl >= n[c]) fl = n[c] - 1;
        pme_bspline<ORDER>(u - fl, ps.th[c], ps.dth[c]);
        ps.k0[c] = fl - ORDER + 1;
    }
}

__device__ __forceinline__ int pme_wrap(int i, int n) {
    i += i < 0 ? n : 0;
    return i - (i >= n ? n : 0);
}

// One thread per site slot: ORDER^3 fixed-point atomics.  Two accumulators per replica:
//   acc[0] = Q1 = environment + displaced atoms,  acc[1] = Q2 - Q1 = ghosts - displaced atoms
// so the environment (almost every site) is spread exactly once.
// (A cooperative variant -- B-spline weights staged in shared memory, the block walking the (site, grid point) items
// with z fastest so that a warp-wide RED touches ~13 sectors instead of 32 -- was measured and is SLOWER, 295 vs 257 us
// at 22 replicas: the limit is the L2 atomic-operation rate (64 M 64-bit REDs per launch), not the sector count.)
template <int ORDER>
__global__ void __launch_bounds__(128) pme_spread_kernel(NbDev d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= CL * d.nclusters[r]) return;
    const size_t rs = (size_t)r * d.Smax + s;
    if (d.slot_site[rs] < 0) return;
    const int cls = d.cmeta[(size_t)r * d.Cmax + (s >> 3)] & 0xffff;
    const int kind = class_kind(cls, d.G);
    const float4 x = d.xs[rs];
    PmeSite<ORDER> ps;
    pme_site_setup<ORDER>(d, x, d.box[r], ps);
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    unsigned long long *acc1 = d.pme_acc + (size_t)r * 2 * ng, *accd = acc1 + ng;
    const double q = (double)x.w * PME_SCALE;
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
        const int ia = pme_wrap(ps.k0[0] + a, d.gx);
#pragma unroll
        for (int b = 0; b < ORDER; b++) {
            const int ib = pme_wrap(ps.k0[1] + b, d.gy);
            const double qab = q * ps.th[0][a] * ps.th[1][b];
            const size_t row = ((size_t)ia * d.gy + ib) * d.gz;
#pragma unroll
            for (int c = 0; c < ORDER; c++) {
                const int ic = pme_wrap(ps.k0[2] + c, d.gz);
                const long long v = __double2ll_rn(qab * ps.th[2][c]);
                if (kind != 2) atomicAdd(acc1 + row + ic, (unsigned long long)v);        // environment, displaced atoms -> Q1
                if (kind == 1) atomicAdd(accd + row + ic, (unsigned long long)(-v));     // displaced atoms leave in state 2
                if (kind == 2) atomicAdd(accd + row + ic, (unsigned long long)v);        // ghosts arrive in state 2
            }
        }
    }
}

// Q1, Q2 = Q1 + (Q2 - Q1) as doubles; the accumulators are handed back zeroed
__global__ void pme_finalize_kernel(NbDev d) {
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    const size_t i = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);  // two cells per thread: 128-bit accesses
    const int r = blockIdx.y;
    if (i >= ng) return;
    unsigned long long *acc = d.pme_acc + (size_t)r * 2 * ng;
    double *grid = d.pme_grid + (size_t)r * 2 * ng;
    if (i + 1 < ng && (ng & 1) == 0) {
        const ulonglong2 a1 = *reinterpret_cast<const ulonglong2 *>(acc + i), ad = *reinterpret_cast<const ulonglong2 *>(acc + ng + i);
        *reinterpret_cast<ulonglong2 *>(acc + i) = make_ulonglong2(0ull, 0ull);
        if (ad.x != 0ull || ad.y != 0ull) *reinterpret_cast<ulonglong2 *>(acc + ng + i) = make_ulonglong2(0ull, 0ull);
        const long long q1x = (long long)a1.x, q1y = (long long)a1.y;
        *reinterpret_cast<double2 *>(grid + i) = make_double2((double)q1x * (1.0 / PME_SCALE), (double)q1y * (1.0 / PME_SCALE));
        *reinterpret_cast<double2 *>(grid + ng + i) = make_double2((double)(q1x + (long long)ad.x) * (1.0 / PME_SCALE),
                                                                   (double)(q1y + (long long)ad.y) * (1.0 / PME_SCALE));
    } else {
        for (size_t k = i; k < ng && k < i + 2; k++) {
            const long long q1 = (long long)acc[k], dq = (long long)acc[ng + k];
            acc[k] = 0ull;
            if (dq != 0) acc[ng + k] = 0ull;
            grid[k] = (double)q1 * (1.0 / PME_SCALE);
            grid[ng + k] = (double)(q1 + dq) * (1.0 / PME_SCALE);
        }
    }
}

// multiply the spectra by exp(-pi^2 m^2/alpha^2) / (pi V m^2 B(m)); accumulate the two reciprocal energies
__global__ void __launch_bounds__(256) pme_convolve_kernel(NbDev d) {
    const int nzh = d.gz / 2 + 1;
    const size_t nspec = (size_t)d.gx * d.gy * nzh;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y, state = blockIdx.z;
    double en = 0.0;
    if (i < nspec) {
        const int c = (int)(i % nzh), b = (int)((i / nzh) % d.gy), a = (int)(i / ((size_t)nzh * d.gy));
        double2 *spec = d.pme_spec + ((size_t)r * 2 + state) * nspec;
        if (a == 0 && b == 0 && c == 0) {
            spec[i] = make_double2(0.0, 0.0);
        } else {
            const float4 L = d.box[r];
            const double ma = (double)(a <= d.gx / 2 ? a : a - d.gx) / (double)L.x, mb = (double)(b <= d.gy / 2 ? b : b - d.gy) / (double)L.y,
                         mc = (double)c / (double)L.z;
            const double m2 = ma * ma + mb * mb + mc * mc;
            const double V = (double)L.x * (double)L.y * (double)L.z;
            const double fac = 9.869604401089358 / ((double)d.alpha * (double)d.alpha);  // pi^2 / alpha^2
            const double eterm = exp(-fac * m2) / (3.141592653589793 * V * m2 * d.pme_mod[a] * d.pme_mod[d.gx + b] * d.pme_mod[d.gx + d.gy + c]);
            double2 v = spec[i];
            const double w = (c == 0 || (2 * c == d.gz)) ? 1.0 : 2.0;  // half spectrum: the conjugate half counts too
            en = 0.5 * w * eterm * (v.x * v.x + v.y * v.y);
            v.x *= eterm; v.y *= eterm;
            spec[i] = v;
        }
    }
    // block reduction, one fixed-point atomic per block
    __shared__ double red[256 / 32];
    for (int off = 16; off > 0; off >>= 1) en += __shfl_xor_sync(0xffffffffu, en, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = en;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 256 / 32; k++) t += red[k];
        atomicAdd(d.eacc + (size_t)r * EACC_SLOTS + 6 + state, (unsigned long long)__double2ll_rn(t * ENERGY_SCALE));
    }
}

// one thread per site: F = -q (n/L) sum dtheta theta theta phi, into the state-specific accumulators
template <int ORDER>
__global__ void __launch_bounds__(128, 4) pme_gather_kernel(NbDev d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= CL * d.nclusters[r]) return;
    const size_t rs = (size_t)r * d.Smax + s;
    if (d.slot_site[rs] < 0) return;
    const int cls = d.cmeta[(size_t)r * d.Cmax + (s >> 3)] & 0xffff;
    const int kind = class_kind(cls, d.G);
    const float4 x = d.xs[rs];
    const float4 L = d.box[r];
    PmeSite<ORDER> ps;
    pme_site_setup<ORDER>(d, x, L, ps);
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    const double *phi1 = d.pme_grid + (size_t)r * 2 * ng, *phi2 = phi1 + ng;
    double f1x = 0, f1y = 0, f1z = 0, f2x = 0, f2y = 0, f2z = 0;
    const bool want1 = kind != 2, want2 = kind != 1;
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
        const int ia = pme_wrap(ps.k0[0] + a, d.gx);
#pragma unroll
        for (int b = 0; b < ORDER; b++) {
            const int ib = pme_wrap(ps.k0[1] + b, d.gy);
            const size_t row = ((size_t)ia * d.gy + ib) * d.gz;
            const double tx = ps.dth[0][a] * ps.th[1][b], ty = ps.th[0][a] * ps.dth[1][b], tz = ps.th[0][a] * ps.th[1][b];
#pragma unroll
            for (int c = 0; c < ORDER; c++) {
                const int ic = pme_wrap(ps.k0[2] + c, d.gz);
                const double wx = tx * ps.th[2][c], wy = ty * ps.th[2][c], wz = tz * ps.dth[2][c];
                if (want1) { const double p = __ldg(phi1 + row + ic); f1x += wx * p; f1y += wy * p; f1z += wz * p; }
                if (want2) { const double p = __ldg(phi2 + row + ic); f2x += wx * p; f2y += wy * p; f2z += wz * p; }
            }
        }
    }
    const double q = (double)x.w;
    const double sx = -q * d.gx / (double)L.x, sy = -q * d.gy / (double)L.y, sz = -q * d.gz / (double)L.z;
    const size_t cs = (size_t)d.R * d.Smax, rsite = (size_t)r * d.Smax;
    unsigned long long *buf1 = d.buf + 3 * cs + rsite, *buf2 = d.buf + 6 * cs + rsite;
    if (want1) {
        atomicAdd(buf1 + s, (unsigned long long)__double2ll_rn(sx * f1x * FORCE_SCALE));
        atomicAdd(buf1 + cs + s, (unsigned long long)__double2ll_rn(sy * f1y * FORCE_SCALE));
        atomicAdd(buf1 + 2 * cs + s, (unsigned long long)__double2ll_rn(sz * f1z * FORCE_SCALE));
    }
    if (want2) {
        atomicAdd(buf2 + s, (unsigned long long)__double2ll_rn(sx * f2x * FORCE_SCALE));
        atomicAdd(buf2 + cs + s, (unsigned long long)__double2ll_rn(sy * f2y * FORCE_SCALE));
        atomicAdd(buf2 + 2 * cs + s, (unsigned long long)__double2ll_rn(sz * f2z * FORCE_SCALE));
    }
}

// Launch with (optionally) the programmatic-stream-serialization attribute: the kernel may begin while its predecessor
// in the stream drains (see pdl_wait / pdl_trigger).  ATM_B200_PDL=0 in the environment switches the overlap off (A/B).
static bool tight_grid_enabled() {  // ATM_B200_TIGHT_GRID=0: size the force kernel's grid from the list capacity (A/B)
    static const bool on = [] {
        const char *e = getenv("ATM_B200_TIGHT_GRID");
        return !(e && e[0] == '0');
    }();
    return on;
}

static bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("ATM_B200_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static void launch_pme_spread(const NbDev &d, cudaStream_t stream) {
    const dim3 grid((d.Smax + 127) / 128, d.R);
    switch (d.pme_order) {
        case 4: pme_spread_kernel<4><<<grid, 128, 0, stream>>>(d); break;
        case 5: pme_spread_kernel<5><<<grid, 128, 0, stream>>>(d); break;
        case 6: pme_spread_kernel<6><<<grid, 128, 0, stream>>>(d); break;
        case 7: pme_spread_kernel<7><<<grid, 128, 0, stream>>>(d); break;
        default: pme_spread_kernel<8><<<grid, 128, 0, stream>>>(d); break;
    }
}

static void launch_pme_gather(const NbDev &d, cudaStream_t stream) {
    const dim3 grid((d.Smax + 127) / 128, d.R);
    switch (d.pme_order) {
        case 4: pme_gather_kernel<4><<<grid, 128, 0, stream>>>(d); break;
        case 5: pme_gather_kernel<5><<<grid, 128, 0, stream>>>(d); break;
        case 6: pme_gather_kernel<6><<<grid, 128, 0, stream>>>(d); break;
        case 7: pme_gather_kernel<7><<<grid, 128, 0, stream>>>(d); break;
        default: pme_gather_kernel<8><<<grid, 128, 0, stream>>>(d); break;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(NbState *nb, T **ptr, size_t count) {
    void *p = nullptr;
    cudaError_t err = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (err != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(err));
        return ATM_ERR_NOMEM;
    }
    nb->owned.push_back(p);
    *ptr = (T *)p;
    return ATM_OK;
}

template <typename T>
static int dev_upload(NbState *nb, T **ptr, const std::vector<T> &v, cudaStream_t stream) {
    int rc = dev_alloc(nb, ptr, v.size());
    if (rc) return rc;
    if (!v.empty()) ATM_CUDA_CHECK(cudaMemcpyAsync(*ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    return ATM_OK;
}

void nb_destroy(atm_handle *h) {
    if (!h->nb) return;
    if (h->nb->graph_exec) cudaGraphExecDestroy(h->nb->graph_exec);
    if (h->nb->rebuild_graph) cudaGraphExecDestroy(h->nb->rebuild_graph);
    if (h->nb->prune_graph) cudaGraphExecDestroy(h->nb->prune_graph);
    for (void *p : h->nb->pme_owned) cudaFree(p);
    if (h->nb->pme_plans) { cufftDestroy(h->nb->pme_plan_fwd); cufftDestroy(h->nb->pme_plan_bwd); }
    if (h->nb->h_flags) cudaFreeHost(h->nb->h_flags);
    if (h->nb->flags_event) cudaEventDestroy(h->nb->flags_event);
    for (auto &ev : h->nb->prof_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (void *p : h->nb->owned) cudaFree(p);
    delete h->nb;
    h->nb = nullptr;
}

// displacement groups: atoms with the same non-zero float-rounded displacement vector move together
static int derive_groups(atm_handle *h) {
    NbState *nb = h->nb;
    const int N = h->N;
    nb->h_group_of_atom.assign(N, 0);
    nb->h_ghost_atom.clear();
    nb->h_ghost_of_atom.assign(N, -1);
    std::map<std::array<uint32_t, 3>, int> groups;
    for (int a = 0; a < N; a++) {
        float f[3] = {(float)h->displ_by_atom[3 * (size_t)a], (float)h->displ_by_atom[3 * (size_t)a + 1],
                      (float)h->displ_by_atom[3 * (size_t)a + 2]};
        if (f[0] == 0.f && f[1] == 0.f && f[2] == 0.f) continue;
        std::array<uint32_t, 3> key;
        memcpy(key.data(), f, 12);
        auto it = groups.find(key);
        int g;
        if (it == groups.end()) {
            g = (int)groups.size() + 1;
            groups[key] = g;
        } else {
            g = it->second;
        }
        nb->h_group_of_atom[a] = g;
        nb->h_ghost_of_atom[a] = (int)nb->h_ghost_atom.size();
        nb->h_ghost_atom.push_back(a);
    }
    ATM_REQUIRE(groups.size() <= 100, ATM_ERR_UNSUPPORTED,
                "more than 100 distinct displacement vectors (%zu) are not supported", groups.size());
    nb->d.M = (int)nb->h_ghost_atom.size();
    nb->d.G = (int)groups.size();
    return ATM_OK;
}

static int nb_allocate(atm_handle *h, cudaStream_t stream);

int nb_on_displacements_changed(atm_handle *h, cudaStream_t stream) {
    if (!h->nb || !h->nb->ready) return ATM_OK;
    // the site layout depends on the displacement groups and on the slot order: reallocate and require a rebuild
    h->nb->list_valid = false;
    return nb_allocate(h, stream);
}

static void free_owned(NbState *nb) {
    for (void *p : nb->owned) cudaFree(p);
    nb->owned.clear();
    nb->sort_tmp = nullptr;
}

static int nb_allocate(atm_handle *h, cudaStream_t stream) {
    NbState *nb = h->nb;
    ATM_CUDA_CHECK(cudaStreamSynchronize(stream));
    nb->generation++;
    nb->alloc_generation++;
    nb->verified = false;
    nb->flags_pending = false;
    if (!nb->h_flags) ATM_CUDA_CHECK(cudaMallocHost(&nb->h_flags, sizeof(int) * 8));
    if (!nb->flags_event) ATM_CUDA_CHECK(cudaEventCreateWithFlags(&nb->flags_event, cudaEventDisableTiming));
    free_owned(nb);
    int rc = derive_groups(h);
    if (rc) return rc;
    NbDev &d = nb->d;
    const int N = h->N, R = h->R;
    d.N = N; d.P = h->P; d.R = R;
    d.U = N + d.M;
    d.cutoff2 = (float)(nb->desc.cutoff * nb->desc.cutoff);
    d.rlist = (float)(nb->desc.cutoff + nb->desc.skin);
    d.rlist_outer = (float)(nb->desc.cutoff + std::max(nb->desc.skin, nb->desc.skin_outer));
    d.alpha = (float)nb->desc.ewald_alpha;
    d.two_alpha_over_sqrtpi = (float)(2.0 * nb->desc.ewald_alpha / sqrt(M_PI));
    d.displ = h->d_displ;
    d.params = h->d_params;

    // column grid from the first replica's box and the mean density: cubes holding ~8 atoms
    ATM_REQUIRE(nb->box_set, ATM_ERR_STATE, "atm_set_box must be called before the neighbour structure is allocated");
    const double Lx = nb->h_box[0], Ly = nb->h_box[1], Lz = nb->h_box[2];
    const double edge = cbrt((double)CL * Lx * Ly * Lz / std::max(1, N));
    d.nx = std::max(1, (int)floor(Lx / edge + 0.5));
    d.ny = std::max(1, (int)floor(Ly / edge + 0.5));
    d.ncol = d.nx * d.ny;
    d.nbins = d.ncol * (2 * d.G + 1);
    std::vector<int> group_count(d.G + 1, 0);
    for (int a = 0; a < N; a++) group_count[nb->h_group_of_atom[a]]++;
    d.CLmax = 0;
    // a class adds at most one partially filled cluster per non-empty column bin
    for (int g = 1; g <= d.G; g++) d.CLmax += group_count[g] / CL + std::min(group_count[g], d.ncol);
    d.CXmax = 2 * d.CLmax;
    d.CenvMax = (N - d.M + CL - 1) / CL + d.ncol;
    d.Cmax = d.CenvMax + d.CXmax;
    d.Smax = d.Cmax * CL;
    // list capacities from the expected partner count of an 8-atom cube (half list for env, full for ligand/ghost)
    {
        const double rho = (double)N / (Lx * Ly * Lz);
        const double a = edge, rl = d.rlist_outer;
        const double vol = a * a * a + 6 * a * a * rl + 3 * M_PI * a * rl * rl + 4.0 / 3.0 * M_PI * rl * rl * rl;
        const double full = vol * rho;
        if (d.capC == 0) d.capC = 32 * (int)ceil(0.5 * full * 1.6 / 32.0 + 2);
        if (d.capX == 0) d.capX = 32 * (int)ceil(full * 2.0 / 32.0 + 2);
    }
    ATM_REQUIRE((long long)d.Smax < (1ll << 24), ATM_ERR_UNSUPPORTED, "more than 2^24 sites per replica");

    // static by-atom arrays
    float *qp; float2 *par; int *es, *el, *goa, *ga, *gof, *soa, *aos;
    if ((rc = dev_upload(nb, &qp, nb->h_qp, stream))) return rc;
    if ((rc = dev_upload(nb, &par, nb->h_par, stream))) return rc;
    if ((rc = dev_upload(nb, &es, nb->h_excl_start, stream))) return rc;
    if ((rc = dev_upload(nb, &el, nb->h_excl_list, stream))) return rc;
    if ((rc = dev_upload(nb, &goa, nb->h_group_of_atom, stream))) return rc;
    if ((rc = dev_upload(nb, &ga, nb->h_ghost_atom, stream))) return rc;
    if ((rc = dev_upload(nb, &gof, nb->h_ghost_of_atom, stream))) return rc;
    std::vector<int> slot_of_atom(N);
    for (int s = 0; s < N; s++) slot_of_atom[h->atom_index[s]] = s;
    if ((rc = dev_upload(nb, &soa, slot_of_atom, stream))) return rc;
    if ((rc = dev_upload(nb, &aos, h->atom_index, stream))) return rc;
    d.qp_atom = qp; d.par_atom = par; d.excl_start = es; d.excl_list = el; d.group_of_atom = goa;
    d.ghost_atom = ga; d.ghost_of_atom = gof; d.slot_of_atom = soa; d.atom_of_slot = aos;
    if ((rc = dev_upload(nb, &nb->d_excl_pairs, nb->h_excl_pairs, stream))) return rc;
    if ((rc = dev_upload(nb, &nb->d_exc_pairs, nb->h_exc_pairs, stream))) return rc;
    if ((rc = dev_upload(nb, &nb->d_exc_par, nb->h_exc_par, stream))) return rc;
    nb->n_excl = (int)nb->h_excl_pairs.size();
    nb->n_exc = (int)nb->h_exc_pairs.size();

    float4 *box, *invbox;
    if ((rc = dev_alloc(nb, &box, R))) return rc;
    if ((rc = dev_alloc(nb, &invbox, R))) return rc;
    d.box = box; d.invbox = invbox;
    nb->box_dirty = true;

    const size_t RU = (size_t)R * d.U, RS = (size_t)R * d.Smax, RC = (size_t)R * d.Cmax;
    if ((rc = dev_alloc(nb, &d.keys, RU))) return rc;
    if ((rc = dev_alloc(nb, &nb->keys_alt, RU))) return rc;
    if ((rc = dev_alloc(nb, &d.vals, RU))) return rc;
    if ((rc = dev_alloc(nb, &nb->vals_alt, RU))) return rc;
    if ((rc = dev_alloc(nb, &d.bin_count, (size_t)R * d.nbins))) return rc;
    if ((rc = dev_alloc(nb, &d.bin_site_start, (size_t)R * (d.nbins + 1)))) return rc;
    if ((rc = dev_alloc(nb, &d.bin_cluster_start, (size_t)R * (d.nbins + 1)))) return rc;
    if ((rc = dev_alloc(nb, &d.nclusters, R))) return rc;
    if ((rc = dev_alloc(nb, &d.slot_site, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.site_slot, RU))) return rc;
    if ((rc = dev_alloc(nb, &d.slot_src, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.slot_out, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.slot_ghost, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.slot_qp, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.xs, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.par, RS))) return rc;
    if ((rc = dev_alloc(nb, &d.cc, RC))) return rc;
    if ((rc = dev_alloc(nb, &d.ch, RC))) return rc;
    if ((rc = dev_alloc(nb, &d.cmeta, RC))) return rc;
    const size_t per_replica = (size_t)d.CenvMax * d.capC + (size_t)d.CXmax * d.capX + (size_t)d.CLmax * d.capC;
    nb->jlist_entries = per_replica * R;
    if ((rc = dev_alloc(nb, &d.jlist, nb->jlist_entries))) return rc;
    if ((rc = dev_alloc(nb, &d.jlist_outer, nb->jlist_entries))) return rc;
    if ((rc = dev_alloc(nb, &d.outer_nsteps, (size_t)R * (d.Cmax + d.CLmax)))) return rc;
    if ((rc = dev_alloc(nb, &d.list_nsteps, (size_t)R * (d.Cmax + d.CLmax)))) return rc;
    {
        const int chunksC = (d.capC / 32 + ITEM_STEPS - 1) / ITEM_STEPS, chunksX = (d.capX / 32 + ITEM_STEPS - 1) / ITEM_STEPS;
        nb->max_items = R * (d.CenvMax * chunksC + d.CXmax * chunksX + d.CLmax * chunksC);
        d.max_items = nb->max_items;
        if ((rc = dev_alloc(nb, &d.items, (size_t)nb->max_items * (ITEM_STEPS + 1)))) return rc;
        nb->n_items = 0;
    }
    if ((rc = dev_alloc(nb, &d.flags, NUM_FLAGS))) return rc;
    ATM_CUDA_CHECK(cudaMemsetAsync(d.flags, 0, sizeof(int) * NUM_FLAGS, stream));
    if ((rc = dev_alloc(nb, &d.buf, 9 * RS))) return rc;
    if ((rc = dev_alloc(nb, &d.eacc, (size_t)R * EACC_SLOTS))) return rc;
    if ((rc = dev_alloc(nb, &d.energies, (size_t)R * ATM_NUM_ENERGY_SLOTS))) return rc;
    ATM_CUDA_CHECK(cudaMemsetAsync(d.buf, 0, 9 * RS * sizeof(unsigned long long), stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.eacc, 0, (size_t)R * EACC_SLOTS * sizeof(unsigned long long), stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.energies, 0, (size_t)R * ATM_NUM_ENERGY_SLOTS * sizeof(double), stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.list_nsteps, 0, (size_t)R * (d.Cmax + d.CLmax) * sizeof(int), stream));

    // radix sort scratch
    int key_bits = 16;
    while ((1ull << (key_bits - 16)) < (unsigned long long)R * d.nbins) key_bits++;
    nb->sort_bits = std::min(64, key_bits);
    nb->sort_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, nb->sort_tmp_bytes, d.keys, nb->keys_alt, d.vals, nb->vals_alt, (int)RU, 0,
                                    nb->sort_bits, stream);
    char *tmp;
    if ((rc = dev_alloc(nb, &tmp, nb->sort_tmp_bytes))) return rc;
    nb->sort_tmp = tmp;
    ATM_CUDA_CHECK(cudaStreamSynchronize(stream));
    return ATM_OK;
}

static int upload_box_if_dirty(atm_handle *h, cudaStream_t stream) {
    NbState *nb = h->nb;
    if (!nb->box_dirty) return ATM_OK;
    std::vector<float4> b(h->R), ib(h->R);
    for (int r = 0; r < h->R; r++) {
        b[r] = make_float4((float)nb->h_box[3 * r], (float)nb->h_box[3 * r + 1], (float)nb->h_box[3 * r + 2], 0.f);
        ib[r] = make_float4(1.0f / b[r].x, 1.0f / b[r].y, 1.0f / b[r].z, 0.f);
    }
    ATM_CUDA_CHECK(cudaMemcpyAsync((void *)nb->d.box, b.data(), sizeof(float4) * h->R, cudaMemcpyHostToDevice, stream));
    ATM_CUDA_CHECK(cudaMemcpyAsync((void *)nb->d.invbox, ib.data(), sizeof(float4) * h->R, cudaMemcpyHostToDevice, stream));
    ATM_CUDA_CHECK(cudaStreamSynchronize(stream));  // the staging vectors die at return
    nb->box_dirty = false;
    return ATM_OK;
}

// pack must have run on the current coordinates; optionally refresh the cluster boxes/centres first
static int launch_prune(atm_handle *h, cudaStream_t stream, bool refresh_boxes) {
    NbState *nb = h->nb;
    NbDev &d = nb->d;
    const int nlists = d.Cmax + d.CLmax;
    if (refresh_boxes) nl_bbox_kernel<<<dim3((d.Cmax + 127) / 128, d.R), 128, 0, stream>>>(d);
    ATM_CUDA_CHECK(cudaMemsetAsync(d.flags + 4, 0, sizeof(int), stream));      // live item count
    ATM_CUDA_CHECK(cudaMemsetAsync(d.flags + 6, 0, sizeof(int) * (NUM_FLAGS - 6), stream));  // pruned-entry counter [6,7], item buckets
    nl_prune_kernel<<<dim3((nlists + PRUNE_WARPS - 1) / PRUNE_WARPS, d.R), 32 * PRUNE_WARPS, 0, stream>>>(d);
    h->launches += (refresh_boxes ? 1 : 0) + 1;
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

}  // namespace atm

using namespace atm;

extern "C" {

static int check_pending_rebuild(atm_handle *h, bool wait);

static int launch_prune_all(atm_handle *h, const void *posq, cudaStream_t stream) {
    NbDev &d = h->nb->d;
    nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, stream>>>(d, (const float4 *)posq);
    h->launches++;
    return launch_prune(h, stream, /*refresh_boxes=*/true);
}

int atm_nb_prune(atm_handle *h, const void *posq, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && posq, ATM_ERR_INVALID, "atm_nb_prune: null argument");
    ATM_REQUIRE(h->nb && h->nb->ready, ATM_ERR_STATE, "atm_nb_prune: Tier 2 not set up");
    int rc;
    if ((rc = check_pending_rebuild(h, false))) return rc;
    ATM_REQUIRE(h->nb->list_valid, ATM_ERR_STATE, "atm_nb_prune: no outer list (call atm_nb_rebuild)");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    NbState *nb = h->nb;
    if ((rc = upload_box_if_dirty(h, stream))) return rc;
    if (stream == nullptr) return launch_prune_all(h, posq, stream);
    if (!nb->prune_graph || nb->prune_graph_posq != posq || nb->prune_graph_generation != nb->alloc_generation) {
        if (nb->prune_graph) { cudaGraphExecDestroy(nb->prune_graph); nb->prune_graph = nullptr; }
        cudaGraph_t graph = nullptr;
        const uint64_t before = h->launches;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = launch_prune_all(h, posq, stream);
            cudaError_t err = cudaStreamEndCapture(stream, &graph);
            nb->prune_graph_launches = (int)(h->launches - before);
            h->launches = before;
            if (rc == ATM_OK && err == cudaSuccess && graph && cudaGraphInstantiate(&nb->prune_graph, graph, 0) == cudaSuccess) {
                nb->prune_graph_posq = posq;
                nb->prune_graph_generation = nb->alloc_generation;
            } else {
                nb->prune_graph = nullptr;
                cudaGetLastError();
            }
            if (graph) cudaGraphDestroy(graph);
        }
    }
    if (nb->prune_graph) {
        ATM_CUDA_CHECK(cudaGraphLaunch(nb->prune_graph, stream));
        h->launches += nb->prune_graph_launches;
        return ATM_OK;
    }
    return launch_prune_all(h, posq, stream);
}

int atm_nb_setup(atm_handle *h, const atm_nonbonded_desc *desc, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && desc, ATM_ERR_INVALID, "atm_nb_setup: null argument");
    ATM_REQUIRE(h->cfg.precision != ATM_PREC_DOUBLE, ATM_ERR_UNSUPPORTED,
                "atm_nb_setup: the fused direct-space path computes in fp32 with fixed-point accumulation "
                "(single/mixed); double precision is Tier 1 only");
    ATM_REQUIRE(desc->charge && desc->sigma && desc->epsilon, ATM_ERR_INVALID, "atm_nb_setup: null parameter array");
    ATM_REQUIRE(desc->cutoff > 0 && desc->skin >= 0 && desc->skin_outer >= 0 && desc->ewald_alpha >= 0, ATM_ERR_INVALID,
                "atm_nb_setup: cutoff must be > 0, skins and ewald_alpha >= 0");
    ATM_REQUIRE(desc->num_exclusions >= 0 && desc->num_exceptions >= 0, ATM_ERR_INVALID, "atm_nb_setup: negative count");
    ATM_REQUIRE(h->N > 0, ATM_ERR_INVALID, "atm_nb_setup: empty system");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    const int N = h->N;
    if (!h->nb) h->nb = new NbState();
    NbState *nb = h->nb;
    nb->ready = false;
    nb->list_valid = false;
    nb->desc = *desc;
    nb->h_qp.resize(N);
    nb->h_par.resize(N);
    const double sq_ke = sqrt(ONE_4PI_EPS0);
    for (int a = 0; a < N; a++) {
        ATM_REQUIRE(desc->epsilon[a] >= 0 && desc->sigma[a] >= 0, ATM_ERR_INVALID, "atm_nb_setup: negative sigma/epsilon at atom %d", a);
        nb->h_qp[a] = (float)(desc->charge[a] * sq_ke);
        nb->h_par[a] = make_float2((float)(0.5 * desc->sigma[a]), (float)(2.0 * sqrt(desc->epsilon[a])));
    }
    // exclusions -> CSR (both directions) + pair list
    nb->h_excl_start.assign(N + 1, 0);
    nb->h_excl_pairs.resize(desc->num_exclusions);
    for (int k = 0; k < desc->num_exclusions; k++) {
        const int a = desc->exclusions[2 * k], b = desc->exclusions[2 * k + 1];
        ATM_REQUIRE(a >= 0 && a < N && b >= 0 && b < N && a != b, ATM_ERR_INVALID, "atm_nb_setup: bad exclusion %d (%d,%d)", k, a, b);
        nb->h_excl_pairs[k] = make_int2(a, b);
        nb->h_excl_start[a + 1]++;
        nb->h_excl_start[b + 1]++;
    }
    for (int a = 0; a < N; a++) nb->h_excl_start[a + 1] += nb->h_excl_start[a];
    nb->h_excl_list.assign(nb->h_excl_start[N], 0);
    {
        std::vector<int> fill(N, 0);
        for (int k = 0; k < desc->num_exclusions; k++) {
            const int a = desc->exclusions[2 * k], b = desc->exclusions[2 * k + 1];
            nb->h_excl_list[nb->h_excl_start[a] + fill[a]++] = b;
            nb->h_excl_list[nb->h_excl_start[b] + fill[b]++] = a;
        }
    }
    nb->h_exc_pairs.resize(desc->num_exceptions);
    nb->h_exc_par.resize(desc->num_exceptions);
    for (int k = 0; k < desc->num_exceptions; k++) {
        const int a = desc->exception_pairs[2 * k], b = desc->exception_pairs[2 * k + 1];
        ATM_REQUIRE(a >= 0 && a < N && b >= 0 && b < N && a != b, ATM_ERR_INVALID, "atm_nb_setup: bad exception %d", k);
        nb->h_exc_pairs[k] = make_int2(a, b);
        nb->h_exc_par[k] = make_float4((float)(ONE_4PI_EPS0 * desc->exception_params[3 * k]), (float)desc->exception_params[3 * k + 1],
                                       (float)(4.0 * desc->exception_params[3 * k + 2]), 0.f);
    }
    // long-range dispersion correction coefficient (OpenMM theory guide, "Lennard-Jones interaction": the mean of
    // eps sig^6 and eps sig^12 over all atom pairs, same-atom pairs included, times 8 pi N^2 / V); atoms are grouped
    // into (sigma, epsilon) classes so the double sum is over classes
    {
        std::map<std::pair<double, double>, long long> classes;
        for (int a = 0; a < N; a++) classes[std::make_pair(desc->sigma[a], desc->epsilon[a])]++;
        std::vector<std::pair<std::pair<double, double>, long long>> cl(classes.begin(), classes.end());
        double s12 = 0.0, s6 = 0.0;
        for (size_t a = 0; a < cl.size(); a++) {
            const double sa = cl[a].first.first, ea = cl[a].first.second, na = (double)cl[a].second;
            const double sa6 = pow(sa, 6.0);
            s12 += 0.5 * na * (na + 1.0) * ea * sa6 * sa6;
            s6 += 0.5 * na * (na + 1.0) * ea * sa6;
            for (size_t b = a + 1; b < cl.size(); b++) {
                const double sg = 0.5 * (sa + cl[b].first.first), ep = sqrt(ea * cl[b].first.second), nn = na * (double)cl[b].second;
                const double sg6 = pow(sg, 6.0);
                s12 += nn * ep * sg6 * sg6;
                s6 += nn * ep * sg6;
            }
        }
        const double npairs = 0.5 * (double)N * ((double)N + 1.0), rc3 = desc->cutoff * desc->cutoff * desc->cutoff;
        nb->disp_coeff_full = 8.0 * (double)N * (double)N * M_PI * (s12 / npairs / (9.0 * rc3 * rc3 * rc3) - s6 / npairs / (3.0 * rc3));
        nb->d.disp_coeff = nb->disp_on ? nb->disp_coeff_full : 0.0;
    }
    // the desc pointers are not kept
    nb->desc.charge = nb->desc.sigma = nb->desc.epsilon = nullptr;
    nb->desc.exclusions = nb->desc.exception_pairs = nullptr;
    nb->desc.exception_params = nullptr;
    nb->d.capC = nb->d.capX = 0;
    nb->ready = true;
    if (nb->box_set) return nb_allocate(h, stream);
    return ATM_OK;
}

int atm_set_box(atm_handle *h, int32_t replica, const double box[9]) {
    ATM_REQUIRE(h && box, ATM_ERR_INVALID, "atm_set_box: null argument");
    ATM_REQUIRE(replica >= -1 && replica < h->R, ATM_ERR_INVALID, "atm_set_box: replica %d out of range", replica);
    ATM_REQUIRE(box[1] == 0 && box[2] == 0 && box[3] == 0 && box[5] == 0 && box[6] == 0 && box[7] == 0, ATM_ERR_UNSUPPORTED,
                "atm_set_box: triclinic boxes are not supported by this build (rectangular only)");
    ATM_REQUIRE(box[0] > 0 && box[4] > 0 && box[8] > 0, ATM_ERR_INVALID, "atm_set_box: non-positive box edge");
    if (!h->nb) h->nb = new NbState();
    NbState *nb = h->nb;
    if (nb->h_box.empty()) nb->h_box.assign((size_t)3 * h->R, 0.0);
    for (int r = 0; r < h->R; r++)
        if (replica < 0 || replica == r) {
            nb->h_box[3 * r] = box[0]; nb->h_box[3 * r + 1] = box[4]; nb->h_box[3 * r + 2] = box[8];
        }
    bool all = true;
    for (int r = 0; r < h->R; r++) all = all && nb->h_box[3 * r] > 0;
    const bool first = !nb->box_set && all;
    nb->box_set = all;
    nb->box_dirty = true;
    if (first && nb->ready) {
        ATM_CUDA_CHECK(cudaSetDevice(h->device));
        return nb_allocate(h, 0);
    }
    return ATM_OK;
}

// every launch of a rebuild, asynchronous
static int launch_rebuild(atm_handle *h, const float4 *posq, cudaStream_t stream) {
    NbState *nb = h->nb;
    NbDev &d = nb->d;
    int rc;
    const int RU = d.R * d.U;
    ATM_CUDA_CHECK(cudaMemsetAsync(d.bin_count, 0, sizeof(int) * (size_t)d.R * d.nbins, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_site, 0xff, sizeof(int) * (size_t)d.R * d.Smax, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_out, 0xff, sizeof(int) * (size_t)d.R * d.Smax, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_ghost, 0xff, sizeof(int) * (size_t)d.R * d.Smax, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.par, 0, sizeof(float2) * (size_t)d.R * d.Smax, stream));  // padding slots are read (masked)
    ATM_CUDA_CHECK(cudaMemsetAsync(d.flags, 0, sizeof(int) * 8, stream));
    nl_keys_kernel<<<(RU + 255) / 256, 256, 0, stream>>>(d, posq);
    nl_scan_kernel<<<d.R, 1024, 0, stream>>>(d);
    size_t tmp_bytes = nb->sort_tmp_bytes;
    cub::DeviceRadixSort::SortPairs(nb->sort_tmp, tmp_bytes, d.keys, nb->keys_alt, d.vals, nb->vals_alt, RU, 0, nb->sort_bits, stream);
    nl_place_kernel<<<(RU + 255) / 256, 256, 0, stream>>>(d, nb->keys_alt, nb->vals_alt);
    if (d.M > 0) nl_link_ghosts_kernel<<<(d.R * d.M + 127) / 128, 128, 0, stream>>>(d);
    nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, stream>>>(d, posq);
    nl_bbox_kernel<<<dim3((d.Cmax + 127) / 128, d.R), 128, 0, stream>>>(d);
    const int nlists = d.Cmax + d.CLmax;
    nl_build_kernel<<<dim3((nlists + BUILD_WARPS - 1) / BUILD_WARPS, d.R), 32 * BUILD_WARPS, 0, stream>>>(d);
    h->launches += 6 + (d.M > 0 ? 1 : 0);  // keys, scan, place, [link], pack, bbox, build
    if ((rc = launch_prune(h, stream, /*refresh_boxes=*/false))) return rc;
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

// Looks at the capacity / geometry flags of a finished rebuild.  Returns ATM_OK, or an error after marking the lists
// invalid.  `grow` tells the caller that the capacities were raised and the structure must be reallocated.
static int inspect_rebuild_flags(atm_handle *h, const int *flags, bool *grow) {
    NbState *nb = h->nb;
    NbDev &d = nb->d;
    *grow = false;
    if (flags[0] & 2) {
        nb->list_valid = false;
        set_error("atm_nb_rebuild: a cluster's extent + list radius exceeds half the box; box too small for this build");
        return ATM_ERR_UNSUPPORTED;
    }
    if (flags[0] & 1) {
        const int need = flags[1];
        d.capC = std::max(d.capC, 32 * ((int)(need * 1.25) / 32 + 1));
        d.capX = std::max(d.capX, 32 * ((int)(need * 1.25) / 32 + 1));
        *grow = true;
    }
    unsigned long long inner_entries = 0;
    memcpy(&inner_entries, &flags[6], 8);
    nb->stats[2] = (int64_t)(inner_entries / d.R);
    nb->items_seen = flags[4];
    return ATM_OK;
}

// Deferred verification of an asynchronous rebuild (never blocks unless `wait`).
static int check_pending_rebuild(atm_handle *h, bool wait) {
    NbState *nb = h->nb;
    if (!nb || !nb->flags_pending) return ATM_OK;
    cudaError_t q = wait ? cudaEventSynchronize(nb->flags_event) : cudaEventQuery(nb->flags_event);
    if (q == cudaErrorNotReady) return ATM_OK;
    ATM_REQUIRE(q == cudaSuccess, ATM_ERR_CUDA, "pair-list verification failed: %s", cudaGetErrorString(q));
    nb->flags_pending = false;
    bool grow = false;
    int rc = inspect_rebuild_flags(h, nb->h_flags, &grow);
    if (rc) return rc;
    if (grow) {
        nb->list_valid = false;
        nb->needs_realloc = true;
        set_error("a pair list outgrew its capacity during the last asynchronous atm_nb_rebuild; steps since then dropped "
                  "interactions -- call atm_nb_rebuild again (capacities were raised)");
        return ATM_ERR_STATE;
    }
    return ATM_OK;
}

int atm_nb_rebuild(atm_handle *h, const void *posq_, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && posq_, ATM_ERR_INVALID, "atm_nb_rebuild: null argument");
    ATM_REQUIRE(h->nb && h->nb->ready && h->nb->box_set, ATM_ERR_STATE, "atm_nb_rebuild: call atm_nb_setup and atm_set_box first");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    NbState *nb = h->nb;
    const float4 *posq = (const float4 *)posq_;
    int rc;
    if ((rc = check_pending_rebuild(h, false)) && rc != ATM_ERR_STATE) return rc;  // a capacity error is cured right here
    if (nb->needs_realloc) {
        if ((rc = nb_allocate(h, stream))) return rc;
        nb->needs_realloc = false;
        nb->verified = false;
    }
    if ((rc = upload_box_if_dirty(h, stream))) return rc;
    for (int r = 0; r < h->R; r++)
        ATM_REQUIRE(2.0 * nb->d.rlist_outer < std::min({nb->h_box[3 * r], nb->h_box[3 * r + 1], nb->h_box[3 * r + 2]}), ATM_ERR_UNSUPPORTED,
                    "atm_nb_rebuild: box edge smaller than 2*(cutoff+skin)");
    if (nb->verified && stream != nullptr) {
        // steady state: fully asynchronous (a cached CUDA graph when the coordinate buffer is unchanged); the capacity
        // flags travel to pinned host memory and are inspected by the next API call that finds them ready
        if (!nb->rebuild_graph || nb->rebuild_graph_posq != posq_ || nb->rebuild_graph_generation != nb->alloc_generation) {
            if (nb->rebuild_graph) { cudaGraphExecDestroy(nb->rebuild_graph); nb->rebuild_graph = nullptr; }
            cudaGraph_t graph = nullptr;
            const uint64_t before = h->launches;
            if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                rc = launch_rebuild(h, posq, stream);
                cudaError_t err = cudaStreamEndCapture(stream, &graph);
                nb->rebuild_graph_launches = (int)(h->launches - before);
                h->launches = before;
                if (rc == ATM_OK && err == cudaSuccess && graph && cudaGraphInstantiate(&nb->rebuild_graph, graph, 0) == cudaSuccess) {
                    nb->rebuild_graph_posq = posq_;
                    nb->rebuild_graph_generation = nb->alloc_generation;
                } else {
                    nb->rebuild_graph = nullptr;
                    cudaGetLastError();
                }
                if (graph) cudaGraphDestroy(graph);
            }
        }
        if (nb->rebuild_graph) {
            ATM_CUDA_CHECK(cudaGraphLaunch(nb->rebuild_graph, stream));
            h->launches += nb->rebuild_graph_launches;
        } else if ((rc = launch_rebuild(h, posq, stream))) {
            return rc;
        }
        ATM_CUDA_CHECK(cudaMemcpyAsync(nb->h_flags, nb->d.flags, sizeof(int) * 8, cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaEventRecord(nb->flags_event, stream));
        nb->flags_pending = true;
        nb->list_valid = true;
        nb->generation++;
        return ATM_OK;
    }
    // first build after (re)allocation: synchronous, verified, grows the capacities until everything fits
    for (int attempt = 0; attempt < 4; attempt++) {
        NbDev &d = nb->d;
        if ((rc = launch_rebuild(h, posq, stream))) return rc;
        int flags[8];
        ATM_CUDA_CHECK(cudaMemcpyAsync(flags, d.flags, sizeof(flags), cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaStreamSynchronize(stream));
        bool grow = false;
        if ((rc = inspect_rebuild_flags(h, flags, &grow))) return rc;
        if (grow) {
            if ((rc = nb_allocate(h, stream))) return rc;
            if ((rc = upload_box_if_dirty(h, stream))) return rc;
            continue;
        }
        int ncl0 = 0;
        ATM_CUDA_CHECK(cudaMemcpy(&ncl0, d.nclusters, sizeof(int), cudaMemcpyDeviceToHost));
        nb->stats[0] = d.U; nb->stats[1] = ncl0; nb->stats[3] = d.capC; nb->stats[4] = d.capX;
        nb->stats[5] = d.M; nb->stats[6] = d.G; nb->stats[7] = d.ncol;
        // launch bound of the force kernel: the item count of this verified build + 12.5 % (the kernel's grid-stride
        // loop covers any excess); the capacity bound max_items would start ~3x as many blocks, most of them empty
        nb->n_items = tight_grid_enabled() ? std::min(nb->max_items, nb->items_seen + nb->items_seen / 8 + 64) : nb->max_items;
        nb->list_valid = true;
        nb->verified = true;
        nb->generation++;
        return ATM_OK;
    }
    set_error("atm_nb_rebuild: neighbour list capacity could not be satisfied after 4 attempts");
    return ATM_ERR_NOMEM;
}

// the launches of one step (no validation, no uploads): shared by atm_step and the graph capture
static int launch_step(atm_handle *h, const atm_step_io *io, cudaStream_t stream, bool profile) {
    NbState *nb = h->nb;
    NbDev &d = nb->d;
    int rc;
    if (io->posq1 || io->posq2) {
        if ((rc = launch_copy_state(h, io->posq, io->posq_corr, io->posq1, io->posq1_corr, io->posq2, io->posq2_corr, stream))) return rc;
        h->launches++;
    }
    nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, stream>>>(d, (const float4 *)io->posq);
    const int warps_per_block = NB_THREADS / 32;
    int item_blocks = (nb->n_items + warps_per_block - 1) / warps_per_block;
    {   // ATM_B200_NB2_WAVES=k: cap the grid at k resident waves and let the grid-stride loop do the rest (experiment)
        static const int waves = [] { const char *e = getenv("ATM_B200_NB2_WAVES"); return e ? atoi(e) : 0; }();
        if (waves > 0) item_blocks = std::min(item_blocks, h->num_sms * NB_MIN_BLOCKS * waves);
    }
    SpecialArgs spa;
    spa.excl = nb->d_excl_pairs; spa.exc = nb->d_exc_pairs; spa.exc_par = nb->d_exc_par;
    spa.n_excl = nb->n_excl; spa.n_exc = nb->n_exc;
    spa.blocks_per_replica = (nb->n_excl + nb->n_exc + NB_THREADS - 1) / NB_THREADS;
    const int nblocks = item_blocks + spa.blocks_per_replica * d.R;
    {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (profile) {
            if (nb->prof_used == nb->prof_events.size()) {
                cudaEvent_t a, b;
                ATM_CUDA_CHECK(cudaEventCreate(&a));
                ATM_CUDA_CHECK(cudaEventCreate(&b));
                nb->prof_events.push_back(std::make_pair(a, b));
            }
            e0 = nb->prof_events[nb->prof_used].first;
            e1 = nb->prof_events[nb->prof_used].second;
            nb->prof_used++;
            ATM_CUDA_CHECK(cudaEventRecord(e0, stream));
        }
        if (nblocks > 0) {
            // dependent launch on the pack kernel (not when profiling: the event pair must bracket nb2 alone)
            ATM_CUDA_CHECK(launch_dependent(io->collect_stats ? nb2_kernel<true> : nb2_kernel<false>, dim3(nblocks), dim3(NB_THREADS), stream,
                                            !profile, d, item_blocks, (int)io->include_energy, spa));
            h->launches++;
        }
        if (profile) ATM_CUDA_CHECK(cudaEventRecord(e1, stream));
    }
    if (d.pme_on) {
        const size_t ng = (size_t)d.gx * d.gy * d.gz, nspec = (size_t)d.gx * d.gy * (d.gz / 2 + 1);
        launch_pme_spread(d, stream);
        pme_finalize_kernel<<<dim3((unsigned)((ng / 2 + 256) / 256), d.R), 256, 0, stream>>>(d);
        ATM_REQUIRE(cufftSetStream(nb->pme_plan_fwd, stream) == CUFFT_SUCCESS, ATM_ERR_CUDA, "cufftSetStream failed");
        ATM_REQUIRE(cufftExecD2Z(nb->pme_plan_fwd, d.pme_grid, (cufftDoubleComplex *)d.pme_spec) == CUFFT_SUCCESS, ATM_ERR_CUDA,
                    "cufftExecD2Z failed");
        pme_convolve_kernel<<<dim3((unsigned)((nspec + 255) / 256), d.R, 2), 256, 0, stream>>>(d);
        ATM_REQUIRE(cufftSetStream(nb->pme_plan_bwd, stream) == CUFFT_SUCCESS, ATM_ERR_CUDA, "cufftSetStream failed");
        ATM_REQUIRE(cufftExecZ2D(nb->pme_plan_bwd, (cufftDoubleComplex *)d.pme_spec, d.pme_grid) == CUFFT_SUCCESS, ATM_ERR_CUDA,
                    "cufftExecZ2D failed");
        launch_pme_gather(d, stream);
        h->launches += 4;  // own kernels (the FFTs are cuFFT library code)
    }
    // scalar stage + merge; a dependent launch only directly behind nb2 (its pdl_wait needs nb2 as the predecessor)
    ATM_CUDA_CHECK(launch_dependent(nb_merge_kernel, dim3((d.Smax + MERGE2_THREADS - 1) / MERGE2_THREADS, d.R), dim3(MERGE2_THREADS), stream,
                                    !profile && !d.pme_on && nblocks > 0, d, (long long *)io->force, (const long long *)io->force_state1_ext,
                                    (const long long *)io->force_state2_ext, io->energy_ext, (int)io->include_energy));
    h->launches += 2;  // pack, merge
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

static int validate_step(atm_handle *h, const atm_step_io *io, const char *who) {
    ATM_REQUIRE(h && io, ATM_ERR_INVALID, "%s: null argument", who);
    ATM_REQUIRE(h->nb && h->nb->ready && h->nb->list_valid, ATM_ERR_STATE, "%s: no valid neighbour structure (call atm_nb_rebuild)", who);
    ATM_REQUIRE(io->posq && io->force, ATM_ERR_INVALID, "%s: posq and force are required", who);
    if (io->posq1 || io->posq2) {
        ATM_REQUIRE(h->R == 1, ATM_ERR_UNSUPPORTED, "%s: inner-context coordinate outputs need num_replicas == 1", who);
        ATM_REQUIRE(io->posq1 && io->posq2, ATM_ERR_INVALID, "%s: posq1 and posq2 must be given together", who);
        ATM_REQUIRE(!(h->cfg.precision == ATM_PREC_MIXED && !(io->posq_corr && io->posq1_corr && io->posq2_corr)), ATM_ERR_INVALID,
                    "%s: mixed precision inner-context outputs need the three posqCorrection buffers", who);
    }
    return ATM_OK;
}

int atm_step(atm_handle *h, const atm_step_io *io, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if (h && h->nb && (rc = check_pending_rebuild(h, false))) return rc;
    if ((rc = validate_step(h, io, "atm_step"))) return rc;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    if ((rc = upload_params_if_dirty(h, stream))) return rc;
    if ((rc = upload_box_if_dirty(h, stream))) return rc;
    return launch_step(h, io, stream, h->nb->profiling);
}

// Same step, replayed from a cached CUDA graph (one driver call per step instead of 5-6 launches).  The graph is
// re-captured when the buffers, the flags or the pair-list generation change.  `stream` must be a real (non-legacy)
// stream because it is put into capture mode.
int atm_step_graph(atm_handle *h, const atm_step_io *io, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if (h && h->nb && (rc = check_pending_rebuild(h, false))) return rc;
    if ((rc = validate_step(h, io, "atm_step_graph"))) return rc;
    ATM_REQUIRE(stream != nullptr, ATM_ERR_INVALID, "atm_step_graph: needs a non-default stream (it is captured)");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    NbState *nb = h->nb;
    if ((rc = upload_params_if_dirty(h, stream))) return rc;
    if ((rc = upload_box_if_dirty(h, stream))) return rc;
    const bool same = nb->graph_exec && memcmp(&nb->graph_io, io, sizeof(*io)) == 0 && nb->graph_generation == nb->alloc_generation;
    if (!same) {
        if (nb->graph_exec) { cudaGraphExecDestroy(nb->graph_exec); nb->graph_exec = nullptr; }
        cudaGraph_t graph = nullptr;
        const uint64_t launches_before = h->launches;
        ATM_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        rc = launch_step(h, io, stream, false);
        cudaError_t err = cudaStreamEndCapture(stream, &graph);
        h->launches = launches_before;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        ATM_REQUIRE(err == cudaSuccess && graph, ATM_ERR_CUDA, "atm_step_graph: capture failed: %s", cudaGetErrorString(err));
        err = cudaGraphInstantiate(&nb->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        ATM_REQUIRE(err == cudaSuccess, ATM_ERR_CUDA, "atm_step_graph: instantiate failed: %s", cudaGetErrorString(err));
        nb->graph_io = *io;
        nb->graph_generation = nb->alloc_generation;
        nb->graph_nodes = 3 + (io->posq1 ? 1 : 0) + (nb->d.pme_on ? 4 : 0);  // pack, nb2 (+special pairs), [PME x4], scalar stage + merge
    }
    ATM_CUDA_CHECK(cudaGraphLaunch(nb->graph_exec, stream));
    h->launches += nb->graph_nodes;
    return ATM_OK;
}

int atm_profile_enable(atm_handle *h, int32_t on) {
    ATM_REQUIRE(h && h->nb, ATM_ERR_STATE, "atm_profile_enable: Tier 2 not set up");
    h->nb->profiling = on != 0;
    return ATM_OK;
}

// Sum of the device time of the nb2 launches recorded since the last read (CUDA events on the launching stream).
int atm_profile_read(atm_handle *h, double *nb2_ms_total, int32_t *nb2_launches) {
    ATM_REQUIRE(h && h->nb && nb2_ms_total && nb2_launches, ATM_ERR_INVALID, "atm_profile_read: null argument");
    NbState *nb = h->nb;
    double total = 0.0;
    for (size_t k = 0; k < nb->prof_used; k++) {
        ATM_CUDA_CHECK(cudaEventSynchronize(nb->prof_events[k].second));
        float ms = 0.f;
        ATM_CUDA_CHECK(cudaEventElapsedTime(&ms, nb->prof_events[k].first, nb->prof_events[k].second));
        total += ms;
    }
    *nb2_ms_total = total;
    *nb2_launches = (int32_t)nb->prof_used;
    nb->prof_used = 0;
    return ATM_OK;
}

int atm_launch_count(atm_handle *h, uint64_t *count) {
    ATM_REQUIRE(h && count, ATM_ERR_INVALID, "atm_launch_count: null argument");
    *count = h->launches;
    return ATM_OK;
}

int atm_energies_device(atm_handle *h, const double **dev_ptr) {
    ATM_REQUIRE(h && dev_ptr, ATM_ERR_INVALID, "atm_energies_device: null argument");
    ATM_REQUIRE(h->nb && h->nb->ready && h->nb->d.energies, ATM_ERR_STATE, "atm_energies_device: Tier 2 not set up");
    *dev_ptr = h->nb->d.energies;
    return ATM_OK;
}

int atm_get_energies(atm_handle *h, double *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && out, ATM_ERR_INVALID, "atm_get_energies: null argument");
    ATM_REQUIRE(h->nb && h->nb->ready && h->nb->d.energies, ATM_ERR_STATE, "atm_get_energies: Tier 2 not set up");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    ATM_CUDA_CHECK(cudaMemcpyAsync(out, h->nb->d.energies, sizeof(double) * (size_t)h->R * ATM_NUM_ENERGY_SLOTS, cudaMemcpyDeviceToHost, stream));
    ATM_CUDA_CHECK(cudaStreamSynchronize(stream));
    {
        int rc = check_pending_rebuild(h, false);
        if (rc) return rc;
    }
    for (int r = 0; r < h->R; r++) h->pert_energy[r] = out[(size_t)r * ATM_NUM_ENERGY_SLOTS + ATM_E_USC];
    return ATM_OK;
}

// B-spline moduli |b(m)|^2 of one dimension (host, double)
static void pme_moduli(int n, int order, std::vector<double> &mod) {
    std::vector<double> th(order, 0.0);
    // spline values at the integers: the recursion at w = 0
    th[1] = 0.0; th[0] = 1.0;
    for (int k = 3; k <= order; k++) {
        const double div = 1.0 / (k - 1.0), w = 0.0;
        th[k - 1] = div * w * th[k - 2];
        for (int j = 1; j <= k - 2; j++) th[k - j - 1] = div * ((w + j) * th[k - j - 2] + (k - j - w) * th[k - j - 1]);
        th[0] = div * (1.0 - w) * th[0];
    }
    mod.assign(n, 0.0);
    for (int m = 0; m < n; m++) {
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < order; k++) {
            const double arg = 2.0 * M_PI * m * k / n;
            sr += th[k] * cos(arg);
            si += th[k] * sin(arg);
        }
        mod[m] = sr * sr + si * si;
    }
    for (int m = 0; m < n; m++)
        if (mod[m] < 1e-7) mod[m] = 0.5 * (mod[(m + n - 1) % n] + mod[(m + 1) % n]);
}

int atm_pme_setup(atm_handle *h, int32_t nx, int32_t ny, int32_t nz, int32_t order) {
    ATM_REQUIRE(h && h->nb && h->nb->ready, ATM_ERR_STATE, "atm_pme_setup: call atm_nb_setup first");
    NbState *nb = h->nb;
    NbDev &d = nb->d;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    ATM_CUDA_CHECK(cudaDeviceSynchronize());
    for (void *p : nb->pme_owned) cudaFree(p);
    nb->pme_owned.clear();
    if (nb->pme_plans) { cufftDestroy(nb->pme_plan_fwd); cufftDestroy(nb->pme_plan_bwd); nb->pme_plans = false; }
    d.pme_on = 0;
    nb->alloc_generation++;  // cached graphs are stale
    if (nx == 0 && ny == 0 && nz == 0) return ATM_OK;  // switched off
    ATM_REQUIRE(order >= 4 && order <= PME_MAX_ORDER, ATM_ERR_INVALID, "atm_pme_setup: spline order must be 4..%d", PME_MAX_ORDER);
    ATM_REQUIRE(nx >= order && ny >= order && nz >= order, ATM_ERR_INVALID, "atm_pme_setup: grid smaller than the spline order");
    ATM_REQUIRE(nb->desc.ewald_alpha > 0, ATM_ERR_INVALID, "atm_pme_setup: needs ewald_alpha > 0");
    const size_t ng = (size_t)nx * ny * nz, nspec = (size_t)nx * ny * (nz / 2 + 1);
    const int R = h->R;
    auto alloc = [&](void **ptr, size_t bytes) -> int {
        if (cudaMalloc(ptr, bytes) != cudaSuccess) { set_error("atm_pme_setup: cudaMalloc of %zu bytes failed", bytes); return ATM_ERR_NOMEM; }
        nb->pme_owned.push_back(*ptr);
        return ATM_OK;
    };
    int rc;
    void *p;
    if ((rc = alloc(&p, sizeof(unsigned long long) * 2 * ng * R))) return rc;
    d.pme_acc = (unsigned long long *)p;
    ATM_CUDA_CHECK(cudaMemset(d.pme_acc, 0, sizeof(unsigned long long) * 2 * ng * R));
    if ((rc = alloc(&p, sizeof(double) * 2 * ng * R))) return rc;
    d.pme_grid = (double *)p;
    if ((rc = alloc(&p, sizeof(double2) * 2 * nspec * R))) return rc;
    d.pme_spec = (double2 *)p;
    std::vector<double> mods, m1;
    for (int n : {nx, ny, nz}) {
        pme_moduli(n, order, m1);
        mods.insert(mods.end(), m1.begin(), m1.end());
    }
    if ((rc = alloc(&p, sizeof(double) * mods.size()))) return rc;
    ATM_CUDA_CHECK(cudaMemcpy(p, mods.data(), sizeof(double) * mods.size(), cudaMemcpyHostToDevice));
    d.pme_mod = (const double *)p;
    int dims[3] = {nx, ny, nz};
    ATM_REQUIRE(cufftPlanMany(&nb->pme_plan_fwd, 3, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 2 * R) == CUFFT_SUCCESS, ATM_ERR_CUDA,
                "atm_pme_setup: cufftPlanMany(D2Z) failed");
    if (cufftPlanMany(&nb->pme_plan_bwd, 3, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 2 * R) != CUFFT_SUCCESS) {
        cufftDestroy(nb->pme_plan_fwd);
        set_error("atm_pme_setup: cufftPlanMany(Z2D) failed");
        return ATM_ERR_CUDA;
    }
    nb->pme_plans = true;
    double self = 0.0, qtot = 0.0;
    for (float q : nb->h_qp) { self += (double)q * (double)q; qtot += (double)q; }
    d.pme_self_sum = self;
    d.pme_qtot2 = qtot * qtot;
    d.gx = nx; d.gy = ny; d.gz = nz;
    d.pme_order = order;
    d.pme_on = 1;
    return ATM_OK;
}

int atm_nb_set_dispersion_correction(atm_handle *h, int32_t on) {
    ATM_REQUIRE(h && h->nb && h->nb->ready, ATM_ERR_STATE, "atm_nb_set_dispersion_correction: call atm_nb_setup first");
    NbState *nb = h->nb;
    nb->disp_on = on != 0;
    nb->d.disp_coeff = nb->disp_on ? nb->disp_coeff_full : 0.0;
    nb->alloc_generation++;  // kernel arguments of the cached step graph are stale
    return ATM_OK;
}

int atm_nb_stats(atm_handle *h, int64_t out[8]) {
    ATM_REQUIRE(h && out, ATM_ERR_INVALID, "atm_nb_stats: null argument");
    ATM_REQUIRE(h->nb && h->nb->list_valid, ATM_ERR_STATE, "atm_nb_stats: no neighbour structure");
    memcpy(out, h->nb->stats, sizeof(int64_t) * 8);
    return ATM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Hooks of the host-buffer pipeline (atm_host.cu): the Tier-2 state is private to this file
// ------------------------------------------------------------------------------------------------
namespace atm {

// Everything a pipeline step must settle on the host BEFORE it enqueues (or replays) the device work of one handle:
// deferred capacity check of the last asynchronous rebuild, parameter / box uploads, box-size check of a rebuild.
// *needs_sync_rebuild: the pair lists must first be (re)built through the synchronous, verified atm_nb_rebuild.
int nb_host_prepare(atm_handle *h, int maintenance, cudaStream_t stream, bool *needs_sync_rebuild) {
    ATM_REQUIRE(h && h->nb && h->nb->ready && h->nb->box_set, ATM_ERR_STATE, "atm_host_pipeline_step: call atm_nb_setup and atm_set_box first");
    NbState *nb = h->nb;
    int rc = check_pending_rebuild(h, false);
    if (rc && !(maintenance == 2 && rc == ATM_ERR_STATE)) return rc;  // a capacity error is cured by the rebuild requested now
    *needs_sync_rebuild = nb->needs_realloc || !nb->verified || !nb->list_valid;
    ATM_REQUIRE(maintenance == 2 || !*needs_sync_rebuild, ATM_ERR_STATE,
                "atm_host_pipeline_step: no valid neighbour structure (the first step must ask for a rebuild)");
    if ((rc = upload_params_if_dirty(h, stream))) return rc;
    if (!*needs_sync_rebuild && (rc = upload_box_if_dirty(h, stream))) return rc;
    if (maintenance == 2)
        for (int r = 0; r < h->R; r++)
            ATM_REQUIRE(2.0 * nb->d.rlist_outer < std::min({nb->h_box[3 * r], nb->h_box[3 * r + 1], nb->h_box[3 * r + 2]}), ATM_ERR_UNSUPPORTED,
                        "atm_host_pipeline_step: box edge smaller than 2*(cutoff+skin)");
    return ATM_OK;
}

// The device work of one handle for one step on device staging buffers: [rebuild | prune] + the step itself.  No
// validation, no uploads, no synchronisation: safe inside a stream capture.
int nb_host_enqueue(atm_handle *h, const void *posq, long long *force, int include_energy, int maintenance, cudaStream_t stream) {
    NbState *nb = h->nb;
    int rc;
    if (maintenance == 2) {
        if ((rc = launch_rebuild(h, (const float4 *)posq, stream))) return rc;
        ATM_CUDA_CHECK(cudaMemcpyAsync(nb->h_flags, nb->d.flags, sizeof(int) * 8, cudaMemcpyDeviceToHost, stream));
    } else if (maintenance == 1) {
        if ((rc = launch_prune_all(h, posq, stream))) return rc;
    }
    atm_step_io io{};
    io.posq = posq;
    io.force = (int64_t *)force;
    io.include_energy = include_energy;
    return launch_step(h, &io, stream, false);
}

// Bookkeeping after a replayed asynchronous rebuild: `stream` is ordered behind the copy of the capacity flags.
int nb_host_rebuild_enqueued(atm_handle *h, cudaStream_t stream) {
    NbState *nb = h->nb;
    ATM_CUDA_CHECK(cudaEventRecord(nb->flags_event, stream));
    nb->flags_pending = true;
    nb->list_valid = true;
    nb->generation++;
    return ATM_OK;
}

uint64_t nb_alloc_generation(const atm_handle *h) { return h->nb ? h->nb->alloc_generation : 0; }
const double *nb_energies_device(const atm_handle *h) { return h->nb ? h->nb->d.energies : nullptr; }

}  // namespace atm
