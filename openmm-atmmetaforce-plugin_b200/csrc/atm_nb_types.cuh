// atm_nb_types.cuh -- Tier 2 shared definitions: compile-time constants, the by-value kernel argument block (NbDev),
// the host-side state of a handle (NbState) and the small device helpers every Tier-2 kernel uses.
// Included by atm_nb.cu only (one translation unit; see the file comment there for the formulation).
#pragma once

#include "atm_common.cuh"
#include <cufft.h>

#include <utility>
#include <vector>

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
constexpr int NUM_FLAGS = ITEM_BUCKET0 + ITEM_STEPS + 8;   // rebuild flags, then one bucket counter per item length
// flags[0] bit 0: a list outgrew its capacity, bit 1: box too small, bit 2: a bin outgrew its sort segment; [1] longest overflowing list; [2,3] outer entries
// (64 bit); [4] live work items; [5] outer work items; [6,7] pruned entries (64 bit); [8], [9] longest list of the
// environment / ligand-ghost capacity class at the last build
constexpr int FLAG_MAXLEN_C = 8, FLAG_MAXLEN_X = 9;
constexpr int NUM_HOST_FLAGS = 16;   // what a rebuild copies to pinned host memory
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
    unsigned long long *binbuf;   // own sort: [R * nbins][bin_cap] unsorted (z16 << 32 | site) entries of each bin
    int bin_cap;                  // power of two <= 1024; 0 = the CUB radix-sort front end is used instead
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
    int *flags;   // rebuild flags (layout below)
    int *iflags;  // counters of the INNER list in use: [4] live work items, [6,7] kept entries, [ITEM_BUCKET0 + n] items of n steps
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
    // single-precision mesh pipeline (default): Q1 and dQ = Q2 - Q1 / phi1 and dphi as floats, see atm_nb_pme.cuh
    int pme_f32, pme_ntx, pme_nty; // tiles of mesh cells in x and y owned by one spread block each
    int pme_tile_cells;            // ints of shared memory reserved for a tile (the list of contributing sites follows)
    float *pme_gridf;              // [R][2][ng]
    float2 *pme_specf;             // [R][2][gx][gy][gz/2+1]
    float *pme_blend;              // [R][gx][gy][stride] blended potential phi1 + sp dphi in the gather's padded row layout
    double pme_self_sum;           // sum of (q sqrt(ke))^2 over all atoms
    double pme_qtot2;              // (sum of q sqrt(ke))^2: neutralising-background term -pi Q^2 / (2 V alpha^2)
    double disp_coeff;             // long-range dispersion correction = disp_coeff / V (0 = off)
};

// The pruned (inner) pair list exists twice: a prune that runs CONCURRENTLY with a step (on the handle's side stream)
// writes the copy that is not in use; the next step switches over.
struct InnerBuf {
    unsigned int *jlist = nullptr;
    int *list_nsteps = nullptr;
    int4 *items = nullptr;
    int *iflags = nullptr;
};

struct StepGraph {   // one cached CUDA graph of a step: keyed by the io block, the inner-list copy in use and the allocation
    atm_step_io io{};
    int cur = 0;
    uint64_t generation = 0;
    cudaGraphExec_t exec = nullptr;
    int nodes = 0;
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
    size_t pme_tile_smem = 0;       // dynamic shared memory of one spread tile
    uint64_t generation = 0;        // bumped by every rebuild
    uint64_t alloc_generation = 0;  // bumped by every (re)allocation: buffers and grid bounds change, graphs are stale
    bool verified = false, needs_realloc = false, flags_pending = false;
    int grow_capC = 0, grow_capX = 0;
    bool use_cub_sort = false;      // a bin outgrew bin_cap once, or the bins are too populous for the in-warp sort
    bool grow_pending = false;      // a list came within 20 % of its capacity: reallocate (larger) at the next rebuild
    bool overflowed = false;        // the last asynchronous rebuild truncated a list (results since then are NaN-poisoned)
    int *h_flags = nullptr;         // pinned
    cudaEvent_t flags_event = nullptr;
    cudaGraphExec_t rebuild_graph = nullptr, prune_graph = nullptr;
    const void *rebuild_graph_posq = nullptr, *prune_graph_posq = nullptr;
    uint64_t rebuild_graph_generation = 0, prune_graph_generation = 0;
    int rebuild_graph_launches = 0, prune_graph_launches = 0;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    std::vector<StepGraph> step_graphs;
    // concurrent prune (atm_step_io.concurrent_prune)
    InnerBuf inner[2];
    int cur = 0;                    // which copy the force kernel reads
    float4 *xs_side = nullptr;      // cluster-order coordinates packed by the side stream
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraphExec_t prune_graph_alt = nullptr;   // synchronous prune into copy 1
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
// same, for an address the compiler cannot prove to be global (a pointer pinned in a register, see pin64 in
// atm_nb_force.cuh): names the state space so that no generic-address ATOM + QSPC sequence is emitted
__device__ __forceinline__ void red_add_fixed_global(unsigned long long addr, float f) {
    const long long v = __float2ll_rn(f * 4294967296.0f);
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

}  // namespace atm
