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
//
// Files of this translation unit: atm_nb_types.cuh (constants, NbDev / NbState, device helpers), atm_nb_lists.cuh (pair-list
// kernels), atm_nb_force.cuh (nb2 + merge), atm_nb_pme.cuh (PME kernels); this file holds the host side: allocation, the
// launch sequences of a rebuild / prune / step, the CUDA-graph caches and the extern "C" entry points.
#include <cub/device/device_radix_sort.cuh>
#include <cufft.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "atm_common.cuh"

#include "atm_nb_types.cuh"
#include "atm_nb_lists.cuh"
#include "atm_nb_force.cuh"
#include "atm_nb_pme.cuh"

namespace atm {

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

static void launch_pme_spread_tile(const NbDev &d, size_t smem, cudaStream_t stream) {
    const dim3 grid(d.pme_ntx * d.pme_nty, d.R);
    switch (d.pme_order) {
        case 4: pme_spread_tile_kernel<4><<<grid, PME_SPREAD_THREADS, smem, stream>>>(d); break;
        case 5: pme_spread_tile_kernel<5><<<grid, PME_SPREAD_THREADS, smem, stream>>>(d); break;
        case 6: pme_spread_tile_kernel<6><<<grid, PME_SPREAD_THREADS, smem, stream>>>(d); break;
        case 7: pme_spread_tile_kernel<7><<<grid, PME_SPREAD_THREADS, smem, stream>>>(d); break;
        default: pme_spread_tile_kernel<8><<<grid, PME_SPREAD_THREADS, smem, stream>>>(d); break;
    }
}

// the gather kernel lives on L1 hits of the two potential meshes and uses no shared memory: ask for the largest L1
static cudaError_t pme_gather_prefer_l1(int order) {
    const void *fn;
    switch (order) {
        case 4: fn = (const void *)pme_gather_f_kernel<4>; break;
        case 5: fn = (const void *)pme_gather_f_kernel<5>; break;
        case 6: fn = (const void *)pme_gather_f_kernel<6>; break;
        case 7: fn = (const void *)pme_gather_f_kernel<7>; break;
        default: fn = (const void *)pme_gather_f_kernel<8>; break;
    }
    return cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
}

static cudaError_t pme_spread_tile_smem_attr(int order, size_t smem) {
    const void *fn;
    switch (order) {
        case 4: fn = (const void *)pme_spread_tile_kernel<4>; break;
        case 5: fn = (const void *)pme_spread_tile_kernel<5>; break;
        case 6: fn = (const void *)pme_spread_tile_kernel<6>; break;
        case 7: fn = (const void *)pme_spread_tile_kernel<7>; break;
        default: fn = (const void *)pme_spread_tile_kernel<8>; break;
    }
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

static void launch_pme_gather_f(const NbDev &d, cudaStream_t stream) {
    const dim3 grid((d.Smax + 127) / 128, d.R);
    switch (d.pme_order) {
        case 4: pme_gather_f_kernel<4><<<grid, 128, 0, stream>>>(d); break;
        case 5: pme_gather_f_kernel<5><<<grid, 128, 0, stream>>>(d); break;
        case 6: pme_gather_f_kernel<6><<<grid, 128, 0, stream>>>(d); break;
        case 7: pme_gather_f_kernel<7><<<grid, 128, 0, stream>>>(d); break;
        default: pme_gather_f_kernel<8><<<grid, 128, 0, stream>>>(d); break;
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
    for (StepGraph &g : h->nb->step_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->nb->prune_graph_alt) cudaGraphExecDestroy(h->nb->prune_graph_alt);
    if (h->nb->side_stream) { cudaStreamSynchronize(h->nb->side_stream); cudaStreamDestroy(h->nb->side_stream); }
    if (h->nb->ev_fork) cudaEventDestroy(h->nb->ev_fork);
    if (h->nb->ev_join) cudaEventDestroy(h->nb->ev_join);
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

static void use_inner(NbDev &d, const InnerBuf &b) {
    d.jlist = b.jlist;
    d.list_nsteps = b.list_nsteps;
    d.items = b.items;
    d.iflags = b.iflags;
}

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
    nb->grow_pending = false;
    nb->overflowed = false;
    if (!nb->h_flags) ATM_CUDA_CHECK(cudaMallocHost(&nb->h_flags, sizeof(int) * (NUM_HOST_FLAGS + 8)));
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
    // work items carry the replica in 8 bits and the list offset in a 32-bit int (atm_nb_lists.cuh, nl_prune_kernel)
    ATM_REQUIRE(R <= 256, ATM_ERR_UNSUPPORTED, "more than 256 replicas per handle (%d): split them over several handles", R);
    {
        const unsigned long long per_replica_entries =
            (unsigned long long)d.CenvMax * d.capC + (unsigned long long)d.CXmax * d.capX + (unsigned long long)d.CLmax * d.capC;
        ATM_REQUIRE(per_replica_entries * (unsigned long long)R < (1ull << 31), ATM_ERR_UNSUPPORTED,
                    "pair-list storage of %llu entries exceeds 2^31: use fewer replicas per handle", per_replica_entries * (unsigned long long)R);
    }

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
    ATM_CUDA_CHECK(cudaMemsetAsync(d.bin_count, 0, sizeof(int) * (size_t)R * d.nbins, stream));   // the own sort keeps it zeroed between rebuilds
    {   // own sort front end: a fixed-capacity segment per bin, sorted inside one warp.  Capacity = twice the mean population
        // of an environment column (+32), a power of two; beyond 1024 entries (8 KB of shared memory per warp) the CUB
        // radix sort front end is used instead.  ATM_B200_CUB_SORT=1 forces the CUB path (A/B, bit-identical results).
        static const bool force_cub = [] { const char *e = getenv("ATM_B200_CUB_SORT"); return e && e[0] == '1'; }();
        const double mean = (double)d.U / std::max(1, d.ncol);
        int cap = 64;
        while (cap < 2.0 * mean + 32.0) cap <<= 1;
        d.bin_cap = (cap > 1024 || force_cub || nb->use_cub_sort) ? 0 : cap;
        d.binbuf = nullptr;
        if (d.bin_cap > 0 && (rc = dev_alloc(nb, &d.binbuf, (size_t)R * d.nbins * d.bin_cap))) return rc;
    }
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
    // slots the own sort front end does not write (a truncated bin, clusters beyond the last one) must never hold
    // garbage indices: start from "empty"
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_site, 0xff, sizeof(int) * RS, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_out, 0xff, sizeof(int) * RS, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_ghost, 0xff, sizeof(int) * RS, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_src, 0, sizeof(int) * RS, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.par, 0, sizeof(float2) * RS, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.xs, 0, sizeof(float4) * RS, stream));
    if ((rc = dev_alloc(nb, &d.cc, RC))) return rc;
    if ((rc = dev_alloc(nb, &d.ch, RC))) return rc;
    if ((rc = dev_alloc(nb, &d.cmeta, RC))) return rc;
    const size_t per_replica = (size_t)d.CenvMax * d.capC + (size_t)d.CXmax * d.capX + (size_t)d.CLmax * d.capC;
    nb->jlist_entries = per_replica * R;
    if ((rc = dev_alloc(nb, &d.jlist_outer, nb->jlist_entries))) return rc;
    if ((rc = dev_alloc(nb, &d.outer_nsteps, (size_t)R * (d.Cmax + d.CLmax)))) return rc;
    {
        const int chunksC = (d.capC / 32 + ITEM_STEPS - 1) / ITEM_STEPS, chunksX = (d.capX / 32 + ITEM_STEPS - 1) / ITEM_STEPS;
        nb->max_items = R * (d.CenvMax * chunksC + d.CXmax * chunksX + d.CLmax * chunksC);
        d.max_items = nb->max_items;
        nb->n_items = 0;
    }
    for (InnerBuf &b : nb->inner) {   // two copies of the pruned list (concurrent prune)
        if ((rc = dev_alloc(nb, &b.jlist, nb->jlist_entries))) return rc;
        if ((rc = dev_alloc(nb, &b.list_nsteps, (size_t)R * (d.Cmax + d.CLmax)))) return rc;
        if ((rc = dev_alloc(nb, &b.items, (size_t)nb->max_items * (ITEM_STEPS + 1)))) return rc;
        if ((rc = dev_alloc(nb, &b.iflags, NUM_FLAGS))) return rc;
        ATM_CUDA_CHECK(cudaMemsetAsync(b.iflags, 0, sizeof(int) * NUM_FLAGS, stream));
        ATM_CUDA_CHECK(cudaMemsetAsync(b.list_nsteps, 0, (size_t)R * (d.Cmax + d.CLmax) * sizeof(int), stream));
    }
    nb->cur = 0;
    use_inner(d, nb->inner[0]);
    if ((rc = dev_alloc(nb, &nb->xs_side, RS))) return rc;
    if (!nb->side_stream) {
        // ATM_B200_SIDE_PRIORITY=1: the side stream gets the highest priority, so that the prune's blocks are placed
        // ahead of the force kernel's pending ones instead of filling its tail (experiment switch)
        static const bool high = [] { const char *e = getenv("ATM_B200_SIDE_PRIORITY"); return e && e[0] == '1'; }();
        int least = 0, greatest = 0;
        ATM_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        ATM_CUDA_CHECK(cudaStreamCreateWithPriority(&nb->side_stream, cudaStreamNonBlocking, high ? greatest : least));
        ATM_CUDA_CHECK(cudaEventCreateWithFlags(&nb->ev_fork, cudaEventDisableTiming));
        ATM_CUDA_CHECK(cudaEventCreateWithFlags(&nb->ev_join, cudaEventDisableTiming));
    }
    if ((rc = dev_alloc(nb, &d.flags, NUM_FLAGS))) return rc;
    ATM_CUDA_CHECK(cudaMemsetAsync(d.flags, 0, sizeof(int) * NUM_FLAGS, stream));
    if ((rc = dev_alloc(nb, &d.buf, 9 * RS))) return rc;
    if ((rc = dev_alloc(nb, &d.eacc, (size_t)R * EACC_SLOTS))) return rc;
    if ((rc = dev_alloc(nb, &d.energies, (size_t)R * ATM_NUM_ENERGY_SLOTS))) return rc;
    ATM_CUDA_CHECK(cudaMemsetAsync(d.buf, 0, 9 * RS * sizeof(unsigned long long), stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.eacc, 0, (size_t)R * EACC_SLOTS * sizeof(unsigned long long), stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(d.energies, 0, (size_t)R * ATM_NUM_ENERGY_SLOTS * sizeof(double), stream));

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

// Prunes the outer list at the cluster-order coordinates dv.xs into the inner-list copy dv points at.  The cluster
// centres stay those of the last rebuild (they are only the origin of the minimum-image shifts, see nl_bbox_kernel).
static int launch_prune(atm_handle *h, cudaStream_t stream, const NbDev &dv) {
    const int nlists = dv.Cmax + dv.CLmax;
    ATM_CUDA_CHECK(cudaMemsetAsync(dv.iflags, 0, sizeof(int) * NUM_FLAGS, stream));   // live items, kept entries, item buckets
    nl_prune_kernel<<<dim3((nlists + PRUNE_WARPS - 1) / PRUNE_WARPS, dv.R), 32 * PRUNE_WARPS, 0, stream>>>(dv);
    h->launches += 1;
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

}  // namespace atm

using namespace atm;

extern "C" {

static int check_pending_rebuild(atm_handle *h, bool wait);

static int launch_prune_all(atm_handle *h, const void *posq, cudaStream_t stream) {
    NbDev &d = h->nb->d;
    nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, stream>>>(d, (const float4 *)posq, 1);
    h->launches++;
    return launch_prune(h, stream, d);
}

int atm_nb_prune(atm_handle *h, const void *posq, void *stream_) {
    ATM_NVTX_RANGE("atm_nb_prune");
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
    if (nb->prune_graph_posq != posq || nb->prune_graph_generation != nb->alloc_generation) {   // both cached graphs are stale
        if (nb->prune_graph) { cudaGraphExecDestroy(nb->prune_graph); nb->prune_graph = nullptr; }
        if (nb->prune_graph_alt) { cudaGraphExecDestroy(nb->prune_graph_alt); nb->prune_graph_alt = nullptr; }
        nb->prune_graph_posq = posq;
        nb->prune_graph_generation = nb->alloc_generation;
    }
    cudaGraphExec_t &pg = nb->cur == 0 ? nb->prune_graph : nb->prune_graph_alt;   // the graph that prunes into the copy in use
    if (!pg) {
        cudaGraph_t graph = nullptr;
        const uint64_t before = h->launches;
        if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = launch_prune_all(h, posq, stream);
            cudaError_t err = cudaStreamEndCapture(stream, &graph);
            nb->prune_graph_launches = (int)(h->launches - before);
            h->launches = before;
            if (!(rc == ATM_OK && err == cudaSuccess && graph && cudaGraphInstantiate(&pg, graph, 0) == cudaSuccess)) {
                pg = nullptr;
                cudaGetLastError();
            }
            if (graph) cudaGraphDestroy(graph);
        }
    }
    if (pg) {
        ATM_CUDA_CHECK(cudaGraphLaunch(pg, stream));
        h->launches += nb->prune_graph_launches;
        return ATM_OK;
    }
    return launch_prune_all(h, posq, stream);
}

int atm_nb_setup(atm_handle *h, const atm_nonbonded_desc *desc, void *stream_) {
    ATM_NVTX_RANGE("atm_nb_setup");
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
    nb->cur = 0;   // a rebuild always leaves the pruned list in copy 0 (its cached graph holds that copy's pointers)
    use_inner(d, nb->inner[0]);
    ATM_CUDA_CHECK(cudaMemsetAsync(d.flags, 0, sizeof(int) * NUM_HOST_FLAGS, stream));
    if (d.bin_cap > 0) {
        // own front end: bin (histogram + scatter) -> scan -> in-warp sort + place -> pack + ghost links + bounding boxes
        nl_bin_kernel<<<(RU + 255) / 256, 256, 0, stream>>>(d, posq);
        nl_scan_kernel<<<d.R, 1024, 0, stream>>>(d, 1);
        nl_sort_place_kernel<<<(d.R * d.nbins + SORT_WARPS - 1) / SORT_WARPS, 32 * SORT_WARPS, sizeof(unsigned long long) * SORT_WARPS * d.bin_cap, stream>>>(d);
        nl_pack_bbox_kernel<<<dim3((d.Cmax + 127) / 128, d.R), 128, 0, stream>>>(d, posq);
        h->launches += 4;
    } else {
        ATM_CUDA_CHECK(cudaMemsetAsync(d.bin_count, 0, sizeof(int) * (size_t)d.R * d.nbins, stream));
        ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_site, 0xff, sizeof(int) * (size_t)d.R * d.Smax, stream));
        ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_out, 0xff, sizeof(int) * (size_t)d.R * d.Smax, stream));
        ATM_CUDA_CHECK(cudaMemsetAsync(d.slot_ghost, 0xff, sizeof(int) * (size_t)d.R * d.Smax, stream));
        ATM_CUDA_CHECK(cudaMemsetAsync(d.par, 0, sizeof(float2) * (size_t)d.R * d.Smax, stream));  // padding slots are read (masked)
        nl_keys_kernel<<<(RU + 255) / 256, 256, 0, stream>>>(d, posq);
        nl_scan_kernel<<<d.R, 1024, 0, stream>>>(d, 0);
        size_t tmp_bytes = nb->sort_tmp_bytes;
        cub::DeviceRadixSort::SortPairs(nb->sort_tmp, tmp_bytes, d.keys, nb->keys_alt, d.vals, nb->vals_alt, RU, 0, nb->sort_bits, stream);
        nl_place_kernel<<<(RU + 255) / 256, 256, 0, stream>>>(d, nb->keys_alt, nb->vals_alt);
        if (d.M > 0) nl_link_ghosts_kernel<<<(d.R * d.M + 127) / 128, 128, 0, stream>>>(d);
        nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, stream>>>(d, posq, 1);
        nl_bbox_kernel<<<dim3((d.Cmax + 127) / 128, d.R), 128, 0, stream>>>(d);
        h->launches += 5 + (d.M > 0 ? 1 : 0);  // keys, scan, place, [link], pack, bbox
    }
    const int nlists = d.Cmax + d.CLmax;
    nl_build_kernel<<<dim3((nlists + BUILD_WARPS - 1) / BUILD_WARPS, d.R), 32 * BUILD_WARPS, 0, stream>>>(d);
    h->launches += 1;
    if ((rc = launch_prune(h, stream, d))) return rc;
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
    if (flags[0] & 4) {
        // a bin outgrew its sort segment (a column more than twice as populous as the mean): the structure is incomplete
        // (results were poisoned like a list overflow); from now on the CUB radix sort front end is used
        nb->use_cub_sort = true;
        *grow = true;
    }
    if (flags[0] & 1) {
        const int need = flags[1];
        d.capC = std::max(d.capC, 32 * ((int)(need * 1.25) / 32 + 1));
        d.capX = std::max(d.capX, 32 * ((int)(need * 1.25) / 32 + 1));
        *grow = true;
    } else {
        // head-room: a list within 20 % of its capacity raises that capacity (by half) at the NEXT rebuild, while the
        // lists in use are still complete
        if (!(flags[0] & 4) && flags[FLAG_MAXLEN_C] > (int)(0.8 * d.capC)) { nb->grow_capC = 32 * ((int)(flags[FLAG_MAXLEN_C] * 1.5) / 32 + 1); nb->grow_pending = true; }
        if (!(flags[0] & 4) && flags[FLAG_MAXLEN_X] > (int)(0.8 * d.capX)) { nb->grow_capX = 32 * ((int)(flags[FLAG_MAXLEN_X] * 1.5) / 32 + 1); nb->grow_pending = true; }
    }
    unsigned long long inner_entries = 0;
    memcpy(&inner_entries, &flags[NUM_HOST_FLAGS + 6], 8);   // the inner-list counters travel behind the rebuild flags
    nb->stats[2] = (int64_t)(inner_entries / d.R);
    nb->items_seen = flags[NUM_HOST_FLAGS + 4];
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
        nb->overflowed = true;
        set_error("a pair list outgrew its capacity during the last asynchronous atm_nb_rebuild; the energies and forces of "
                  "the steps since then are NaN -- call atm_nb_rebuild again (capacities were raised) and repeat them");
        return ATM_ERR_STATE;
    }
    return ATM_OK;
}

int atm_nb_check(atm_handle *h, int32_t wait) {
    ATM_REQUIRE(h, ATM_ERR_INVALID, "atm_nb_check: null handle");
    if (!h->nb) return ATM_OK;
    if (h->nb->overflowed && !h->nb->flags_pending) {
        set_error("the pair lists of the last atm_nb_rebuild are incomplete (capacity overflow): call atm_nb_rebuild again");
        return ATM_ERR_STATE;
    }
    int rc = check_pending_rebuild(h, wait != 0);
    if (rc || !wait || !h->nb->d.pme_on || !h->nb->d.pme_f32 || !h->nb->list_valid) return rc;
    // the PME spread relies on no site having left its place in the structure by more than half the outer skin; the
    // gather kernel raises flag bit 3 when one has (the steps since then returned NaN)
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    ATM_CUDA_CHECK(cudaDeviceSynchronize());
    int f0 = 0;
    ATM_CUDA_CHECK(cudaMemcpy(&f0, h->nb->d.flags, sizeof(int), cudaMemcpyDeviceToHost));
    if (f0 & 8) {
        set_error("a site has moved further since the last atm_nb_rebuild than half the outer skin (%.3f nm) beyond its cluster: the "
                  "energies and forces of the steps since then are NaN -- call atm_nb_rebuild and repeat them",
                  0.5 * std::max(h->nb->desc.skin, h->nb->desc.skin_outer));
        return ATM_ERR_STATE;
    }
    return ATM_OK;
}

int atm_nb_rebuild(atm_handle *h, const void *posq_, void *stream_) {
    ATM_NVTX_RANGE("atm_nb_rebuild");
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && posq_, ATM_ERR_INVALID, "atm_nb_rebuild: null argument");
    ATM_REQUIRE(h->nb && h->nb->ready && h->nb->box_set, ATM_ERR_STATE, "atm_nb_rebuild: call atm_nb_setup and atm_set_box first");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    NbState *nb = h->nb;
    const float4 *posq = (const float4 *)posq_;
    int rc;
    if ((rc = check_pending_rebuild(h, false)) && rc != ATM_ERR_STATE) return rc;  // a capacity error is cured right here
    if (nb->needs_realloc || nb->grow_pending) {
        if (nb->grow_pending) {
            nb->d.capC = std::max(nb->d.capC, nb->grow_capC);
            nb->d.capX = std::max(nb->d.capX, nb->grow_capX);
        }
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
            nb->cur = 0;   // what launch_rebuild did when the graph was captured: the pruned list is in copy 0
            use_inner(nb->d, nb->inner[0]);
        } else if ((rc = launch_rebuild(h, posq, stream))) {
            return rc;
        }
        ATM_CUDA_CHECK(cudaMemcpyAsync(nb->h_flags, nb->d.flags, sizeof(int) * NUM_HOST_FLAGS, cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaMemcpyAsync(nb->h_flags + NUM_HOST_FLAGS, nb->d.iflags, sizeof(int) * 8, cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaEventRecord(nb->flags_event, stream));
        nb->flags_pending = true;
        nb->list_valid = true;
        nb->generation++;
        return ATM_OK;
    }
    // first build after (re)allocation: synchronous, verified, grows the capacities until everything fits
    for (int attempt = 0; attempt < 6; attempt++) {
        NbDev &d = nb->d;
        if ((rc = launch_rebuild(h, posq, stream))) return rc;
        int flags[NUM_HOST_FLAGS + 8];
        ATM_CUDA_CHECK(cudaMemcpyAsync(flags, d.flags, sizeof(int) * NUM_HOST_FLAGS, cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaMemcpyAsync(flags + NUM_HOST_FLAGS, d.iflags, sizeof(int) * 8, cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaStreamSynchronize(stream));
        bool grow = false;
        if ((rc = inspect_rebuild_flags(h, flags, &grow))) return rc;
        if (nb->grow_pending) {   // the head-room rule applies to the verified build as well: steady state starts with margin
            d.capC = std::max(d.capC, nb->grow_capC);
            d.capX = std::max(d.capX, nb->grow_capX);
            grow = true;
        }
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
    set_error("atm_nb_rebuild: neighbour list capacity could not be satisfied after 6 attempts");
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
    if (io->concurrent_prune) {
        // Re-prune from THESE coordinates on the side stream while the step runs on the copy in use: its own pack into
        // xs_side, then the prune into the other copy.  The caller's stream joins at the end of the step (the side
        // stream reads posq), and the next step switches copies (flip_inner, called by the entry points per launch).
        NbDev ds = d;
        ds.xs = nb->xs_side;
        use_inner(ds, nb->inner[nb->cur ^ 1]);
        ATM_CUDA_CHECK(cudaEventRecord(nb->ev_fork, stream));
        ATM_CUDA_CHECK(cudaStreamWaitEvent(nb->side_stream, nb->ev_fork, 0));
        nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, nb->side_stream>>>(ds, (const float4 *)io->posq, 0);
        h->launches++;
        if ((rc = launch_prune(h, nb->side_stream, ds))) return rc;
        ATM_CUDA_CHECK(cudaEventRecord(nb->ev_join, nb->side_stream));
    }
    nb_pack_kernel<<<dim3((d.Smax + 255) / 256, d.R), 256, 0, stream>>>(d, (const float4 *)io->posq, 1);
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
    {   // ATM_B200_SPECIAL_FIRST=1: special-pair blocks at the START of the grid instead of the end (A/B switch; measured
        // slower: nb2 0.3296 vs 0.3238 ms at 22 replicas, 0.0550 vs 0.0535 at 3 -- the short blocks delay the first wave
        // of long work items more than they shorten the tail)
        static const int first = [] { const char *e = getenv("ATM_B200_SPECIAL_FIRST"); return (e && e[0] == '1') ? 1 : 0; }();
        spa.special_first = first;
    }
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
    if (d.pme_on && d.pme_f32) {
        launch_pme_spread_tile(d, nb->pme_tile_smem, stream);
        ATM_REQUIRE(cufftSetStream(nb->pme_plan_fwd, stream) == CUFFT_SUCCESS, ATM_ERR_CUDA, "cufftSetStream failed");
        ATM_REQUIRE(cufftExecR2C(nb->pme_plan_fwd, d.pme_gridf, (cufftComplex *)d.pme_specf) == CUFFT_SUCCESS, ATM_ERR_CUDA,
                    "cufftExecR2C failed");
        pme_convolve_f_kernel<<<dim3((unsigned)((d.gx * d.gy + PME_CONV_ROWS - 1) / PME_CONV_ROWS), d.R), 32 * PME_CONV_WARPS,
                                sizeof(double) * (d.gz / 2 + 1), stream>>>(d);
        ATM_REQUIRE(cufftSetStream(nb->pme_plan_bwd, stream) == CUFFT_SUCCESS, ATM_ERR_CUDA, "cufftSetStream failed");
        ATM_REQUIRE(cufftExecC2R(nb->pme_plan_bwd, (cufftComplex *)d.pme_specf, d.pme_gridf) == CUFFT_SUCCESS, ATM_ERR_CUDA,
                    "cufftExecC2R failed");
        pme_blend_kernel<<<dim3((unsigned)(((size_t)d.gx * d.gy * (pme_blend_stride(d.gz, d.pme_order) / 4) + 255) / 256), d.R), 256, 0, stream>>>(
            d, io->energy_ext, (int)io->include_energy);
        launch_pme_gather_f(d, stream);
        h->launches += 4;  // own kernels (the FFTs are cuFFT library code)
    } else if (d.pme_on) {
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
    if (io->concurrent_prune) ATM_CUDA_CHECK(cudaStreamWaitEvent(stream, nb->ev_join, 0));
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

// after a step with a concurrent prune has been ENQUEUED (or its graph replayed): later work uses the other copy
static void flip_inner(NbState *nb) {
    nb->cur ^= 1;
    use_inner(nb->d, nb->inner[nb->cur]);
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
    ATM_NVTX_RANGE("atm_step");
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if (h && h->nb && (rc = check_pending_rebuild(h, false))) return rc;
    if ((rc = validate_step(h, io, "atm_step"))) return rc;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    if ((rc = upload_params_if_dirty(h, stream))) return rc;
    if ((rc = upload_box_if_dirty(h, stream))) return rc;
    ATM_REQUIRE(!(io->concurrent_prune && stream == nullptr), ATM_ERR_INVALID, "atm_step: a concurrent prune needs a non-default stream");
    if ((rc = launch_step(h, io, stream, h->nb->profiling))) return rc;
    if (io->concurrent_prune) flip_inner(h->nb);
    return ATM_OK;
}

// Same step, replayed from a cached CUDA graph (one driver call per step instead of 5-6 launches).  The graph is
// re-captured when the buffers, the flags or the pair-list generation change.  `stream` must be a real (non-legacy)
// stream because it is put into capture mode.
int atm_step_graph(atm_handle *h, const atm_step_io *io, void *stream_) {
    ATM_NVTX_RANGE("atm_step_graph");
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc;
    if (h && h->nb && (rc = check_pending_rebuild(h, false))) return rc;
    if ((rc = validate_step(h, io, "atm_step_graph"))) return rc;
    ATM_REQUIRE(stream != nullptr, ATM_ERR_INVALID, "atm_step_graph: needs a non-default stream (it is captured)");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    NbState *nb = h->nb;
    if ((rc = upload_params_if_dirty(h, stream))) return rc;
    if ((rc = upload_box_if_dirty(h, stream))) return rc;
    // cached graphs: one per (io block, inner-list copy in use); all of them die with the allocation they were captured on
    if (!nb->step_graphs.empty() && nb->step_graphs[0].generation != nb->alloc_generation) {
        for (StepGraph &g : nb->step_graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        nb->step_graphs.clear();
    }
    StepGraph *sg = nullptr;
    for (StepGraph &g : nb->step_graphs)
        if (g.cur == nb->cur && memcmp(&g.io, io, sizeof(*io)) == 0) sg = &g;
    if (!sg) {
        if (nb->step_graphs.size() >= 16) {   // a caller cycling through many buffers: start over
            for (StepGraph &g : nb->step_graphs)
                if (g.exec) cudaGraphExecDestroy(g.exec);
            nb->step_graphs.clear();
        }
        cudaGraph_t graph = nullptr;
        const uint64_t launches_before = h->launches;
        ATM_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        rc = launch_step(h, io, stream, false);
        cudaError_t err = cudaStreamEndCapture(stream, &graph);
        const int nodes = (int)(h->launches - launches_before);
        h->launches = launches_before;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        ATM_REQUIRE(err == cudaSuccess && graph, ATM_ERR_CUDA, "atm_step_graph: capture failed: %s", cudaGetErrorString(err));
        StepGraph g;
        err = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        ATM_REQUIRE(err == cudaSuccess, ATM_ERR_CUDA, "atm_step_graph: instantiate failed: %s", cudaGetErrorString(err));
        g.io = *io;
        g.cur = nb->cur;
        g.generation = nb->alloc_generation;
        g.nodes = nodes;   // own kernels of one replay: [copy-state], [side pack, prune], pack, nb2, [PME x4], merge
        nb->step_graphs.push_back(g);
        sg = &nb->step_graphs.back();
    }
    ATM_CUDA_CHECK(cudaGraphLaunch(sg->exec, stream));
    h->launches += sg->nodes;
    if (io->concurrent_prune) flip_inner(nb);
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
    ATM_NVTX_RANGE("atm_pme_setup");
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
    ATM_REQUIRE(nx >= order && ny >= order && nz >= std::max(order, 8), ATM_ERR_INVALID, "atm_pme_setup: grid smaller than the spline order (or fewer than 8 points along z)");
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
    // ATM_B200_PME_F64=1: the double-precision mesh pipeline of round 1 (global fixed-point spread, D2Z / Z2D) for A/B runs
    static const bool f64 = [] { const char *e = getenv("ATM_B200_PME_F64"); return e && e[0] == '1'; }();
    d.pme_f32 = f64 ? 0 : 1;
    d.pme_acc = nullptr; d.pme_grid = nullptr; d.pme_spec = nullptr; d.pme_gridf = nullptr; d.pme_specf = nullptr; d.pme_blend = nullptr;
    if (d.pme_f32) {
        if ((rc = alloc(&p, sizeof(float) * 2 * ng * R))) return rc;
        d.pme_gridf = (float *)p;
        if ((rc = alloc(&p, sizeof(float2) * 2 * nspec * R))) return rc;
        d.pme_specf = (float2 *)p;
        if ((rc = alloc(&p, sizeof(float) * (size_t)R * nx * ny * pme_blend_stride(nz, order)))) return rc;
        d.pme_blend = (float *)p;
        // spread tiles: about 12 x 12 cells in xy (all of z), smaller when the z extent would not fit in shared memory
        int T = 12;
        {   // ATM_B200_PME_TILE=t: edge of a spread brick in mesh cells (experiment; 12 measured best, profiles/r2x_pme_tile_ab.log)
            static const int t_env = [] { const char *e = getenv("ATM_B200_PME_TILE"); return e ? atoi(e) : 0; }();
            if (t_env >= 4 && t_env <= 32) T = t_env;
        }
        const int nzp = pme_tile_stride(nz, order);   // a tile row: padding for supports that start below z = 0, then the column
        while (T > 4 && sizeof(int) * (size_t)T * T * nzp > 160 * 1024) T--;
        d.pme_ntx = (nx + T - 1) / T;
        d.pme_nty = (ny + T - 1) / T;
        const int tw = (nx + d.pme_ntx - 1) / d.pme_ntx, th = (ny + d.pme_nty - 1) / d.pme_nty;
        d.pme_tile_cells = tw * th * nzp;
        nb->pme_tile_smem = sizeof(int) * ((size_t)d.pme_tile_cells + PME_LIST_CAP);
        ATM_REQUIRE(nb->pme_tile_smem <= 200 * 1024 + sizeof(int) * PME_LIST_CAP, ATM_ERR_UNSUPPORTED, "atm_pme_setup: %d mesh points along z do not fit a shared-memory tile", nz);
        // the opt-in limit is per kernel and device, not per handle: always the largest tile a handle may ask for
        ATM_CUDA_CHECK(pme_spread_tile_smem_attr(order, (size_t)200 * 1024 + sizeof(int) * PME_LIST_CAP));
        ATM_CUDA_CHECK(pme_gather_prefer_l1(order));
    } else {
        if ((rc = alloc(&p, sizeof(unsigned long long) * 2 * ng * R))) return rc;
        d.pme_acc = (unsigned long long *)p;
        ATM_CUDA_CHECK(cudaMemset(d.pme_acc, 0, sizeof(unsigned long long) * 2 * ng * R));
        if ((rc = alloc(&p, sizeof(double) * 2 * ng * R))) return rc;
        d.pme_grid = (double *)p;
        if ((rc = alloc(&p, sizeof(double2) * 2 * nspec * R))) return rc;
        d.pme_spec = (double2 *)p;
    }
    std::vector<double> mods, m1;
    for (int n : {nx, ny, nz}) {
        pme_moduli(n, order, m1);
        mods.insert(mods.end(), m1.begin(), m1.end());
    }
    if ((rc = alloc(&p, sizeof(double) * mods.size()))) return rc;
    ATM_CUDA_CHECK(cudaMemcpy(p, mods.data(), sizeof(double) * mods.size(), cudaMemcpyHostToDevice));
    d.pme_mod = (const double *)p;
    int dims[3] = {nx, ny, nz};
    ATM_REQUIRE(cufftPlanMany(&nb->pme_plan_fwd, 3, dims, nullptr, 1, 0, nullptr, 1, 0, d.pme_f32 ? CUFFT_R2C : CUFFT_D2Z, 2 * R) == CUFFT_SUCCESS,
                ATM_ERR_CUDA, "atm_pme_setup: cufftPlanMany(forward) failed");
    if (cufftPlanMany(&nb->pme_plan_bwd, 3, dims, nullptr, 1, 0, nullptr, 1, 0, d.pme_f32 ? CUFFT_C2R : CUFFT_Z2D, 2 * R) != CUFFT_SUCCESS) {
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
    *needs_sync_rebuild = nb->needs_realloc || !nb->verified || !nb->list_valid || (maintenance == 2 && nb->grow_pending);
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
int nb_host_enqueue(atm_handle *h, const void *posq, long long *force, int include_energy, int maintenance, cudaStream_t stream,
                    const long long *force_state1_ext, const long long *force_state2_ext, const double *energy_ext) {
    NbState *nb = h->nb;
    int rc;
    if (maintenance == 2) {
        if ((rc = launch_rebuild(h, (const float4 *)posq, stream))) return rc;
        ATM_CUDA_CHECK(cudaMemcpyAsync(nb->h_flags, nb->d.flags, sizeof(int) * NUM_HOST_FLAGS, cudaMemcpyDeviceToHost, stream));
        ATM_CUDA_CHECK(cudaMemcpyAsync(nb->h_flags + NUM_HOST_FLAGS, nb->d.iflags, sizeof(int) * 8, cudaMemcpyDeviceToHost, stream));
    } else if (maintenance == 1) {
        if ((rc = launch_prune_all(h, posq, stream))) return rc;
    }
    atm_step_io io{};
    io.posq = posq;
    io.force = (int64_t *)force;
    io.include_energy = include_energy;
    io.concurrent_prune = maintenance == 3;
    io.force_state1_ext = (const int64_t *)force_state1_ext;
    io.force_state2_ext = (const int64_t *)force_state2_ext;
    io.energy_ext = energy_ext;
    return launch_step(h, &io, stream, false);
}

int nb_host_inner_copy(const atm_handle *h) { return h->nb ? h->nb->cur : 0; }
void nb_host_flip_inner(atm_handle *h) { flip_inner(h->nb); }

// Bookkeeping after a replayed asynchronous rebuild: `stream` is ordered behind the copy of the capacity flags.
int nb_host_rebuild_enqueued(atm_handle *h, cudaStream_t stream) {
    NbState *nb = h->nb;
    ATM_CUDA_CHECK(cudaEventRecord(nb->flags_event, stream));
    nb->flags_pending = true;
    nb->list_valid = true;
    nb->generation++;
    nb->cur = 0;   // a (replayed) rebuild leaves the pruned list in copy 0
    use_inner(nb->d, nb->inner[0]);
    return ATM_OK;
}

uint64_t nb_alloc_generation(const atm_handle *h) { return h->nb ? h->nb->alloc_generation : 0; }
const double *nb_energies_device(const atm_handle *h) { return h->nb ? h->nb->d.energies : nullptr; }

}  // namespace atm
