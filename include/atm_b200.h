/*
 * atm_b200.h -- C ABI of the Blackwell (sm_100a) back-end for the ATM Meta-Force per-step hot path.
 *
 * One shared library (libatm_b200.so) replaces the reference's four kernel back-ends
 * (platforms/common, platforms/cuda, platforms/hip, platforms/opencl).  Every entry point is
 * `extern "C"`, takes plain pointers and sizes, returns an int status (0 = ok) and never throws.
 * Device pointers are BORROWED (owned by the caller, e.g. by OpenMM's ComputeContext) unless stated;
 * `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  No entry point
 * synchronises the device unless its comment says so.  One handle per OpenMM Context; calls on one
 * handle must be serialised by the caller (OpenMM contexts are single-threaded), different handles
 * are independent.
 *
 * "ref:" comments cite the reference interface each entry point replaces (paths relative to the
 * reference repository root, Gallicchio-Lab/openmm-atmmetaforce-plugin v0.3.1).
 *
 * Two tiers:
 *   Tier 1 -- drop-in replacements of the reference's two device kernels and its host scalar stage;
 *             OpenMM's inner contexts still evaluate U1,F1 / U2,F2.
 *   Tier 2 -- the fused Blackwell path: the library evaluates the direct-space NonbondedForce of BOTH
 *             inner states itself (one launch, env-env pairs shared), keeps U1/U2/u/W/dW/du on the
 *             device and merges the forces, batched over R replicas that share one System.
 */
#ifndef ATM_B200_H_
#define ATM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATM_B200_VERSION "0.3.1-b200.1"

typedef struct atm_handle atm_handle;

enum atm_status {
    ATM_OK = 0,
    ATM_ERR_INVALID = 1,      /* bad argument (message in atm_last_error) */
    ATM_ERR_CUDA = 2,         /* a CUDA runtime call failed */
    ATM_ERR_STATE = 3,        /* call sequence error (e.g. step before nb_setup) */
    ATM_ERR_UNSUPPORTED = 4,  /* valid request this build does not implement (e.g. triclinic box) */
    ATM_ERR_NOMEM = 5
};

/* ref: platform property "Precision" (example/abfe/abfe.py:109-110); decides the element type of posq:
 * single/mixed -> float4 (+ float4 posqCorrection in mixed), double -> double4. */
enum atm_precision { ATM_PREC_SINGLE = 0, ATM_PREC_MIXED = 1, ATM_PREC_DOUBLE = 2 };

/* Index of each global parameter in a `double p[9]` block.
 * ref: openmmapi/include/ATMMetaForce.h:151-218 (the nine Context parameter names). */
enum atm_param {
    ATM_LAMBDA1 = 0, ATM_LAMBDA2 = 1, ATM_ALPHA = 2, ATM_U0 = 3, ATM_W0 = 4,
    ATM_UMAX = 5, ATM_UBCORE = 6, ATM_ACORE = 7, ATM_DIRECTION = 8, ATM_NUM_PARAMS = 9
};

/* Slots of the per-replica energy record kept on the device (doubles). */
enum atm_energy_slot {
    ATM_E_U1 = 0,      /* state-1 energy of the variable force groups (direct space + external part) */
    ATM_E_U2 = 1,      /* state-2 energy */
    ATM_E_U = 2,       /* raw perturbation energy, +-(U2-U1) by direction, formed from the state-specific pairs only */
    ATM_E_USC = 3,     /* soft-core perturbation energy == ATMMetaForce::getPerturbationEnergy() */
    ATM_E_EBIAS = 4,   /* W(u_sc) */
    ATM_E_ENERGY = 5,  /* e0 + W : what calcForcesAndEnergy returns */
    ATM_E_SP = 6,      /* weight of F2 in the merged force */
    ATM_E_NPAIRS = 7,  /* pairs inside the cutoff evaluated by the last nb2 launch (only with collect_stats) */
    ATM_E_NPAIRS_C = 8,   /* ... of which shared by both states */
    ATM_E_NPAIRS_S1 = 9,  /* ... state-1 only */
    ATM_E_NPAIRS_S2 = 10, /* ... state-2 only */
    ATM_E_UREC1 = 11,     /* PME reciprocal energy of state 1 (0 when atm_pme_setup was not called) */
    ATM_E_UREC2 = 12,     /* ... of state 2 */
    ATM_E_USELF = 13,     /* Ewald self energy (included in U1 and U2 when PME is on) */
    ATM_NUM_ENERGY_SLOTS = 16
};

typedef struct {
    int32_t num_particles;        /* N = force.getNumParticles() (ref: CommonATMMetaForceKernels.cpp:80) */
    int32_t padded_num_particles; /* P = cc.getPaddedNumAtoms(); 0 -> 32*ceil(N/32) (ref: :83) */
    int32_t precision;            /* enum atm_precision */
    int32_t num_replicas;         /* R >= 1 coordinate sets sharing this System (Tier 2 batches them) */
    int32_t device;               /* CUDA device ordinal; -1 = current device */
} atm_config;

/* Thread-local message of the last failing call on this thread. */
const char *atm_last_error(void);
const char *atm_version(void);

/* ref: CudaATMMetaForceKernelFactory::createKernelImpl (platforms/cuda/src/CudaATMMetaForceKernelFactory.cpp:38-43)
 *      + CommonCalcATMMetaForceKernel ctor/dtor.  Allocates only library-owned scratch. */
int atm_create(const atm_config *cfg, atm_handle **out);
int atm_destroy(atm_handle *h);

/* Builds and uploads the float4 displacement table in device (slot) order.
 *   atom_index : [N] slot -> atom (cc.getAtomIndex()); NULL = identity
 *   dxyz       : [N][3] displacement of ATOM a in nm (entry i of the force is atom i; the 'particle'
 *                field is ignored exactly as the reference does)
 * ref: CommonCalcATMMetaForceKernel::initialize (platforms/common/src/CommonATMMetaForceKernels.cpp:78-109),
 *      ReorderListener::execute (:53-72), copyParametersToContext (:229-251).
 * Also (Tier 2) re-derives the displacement groups.  Asynchronous on `stream` (host staging is copied). */
int atm_set_displacements(atm_handle *h, const int32_t *atom_index, const double *dxyz, void *stream);

/* The nine global parameters of one replica (replica = -1: all).
 * ref: context.getParameter(...) reads in execute (CommonATMMetaForceKernels.cpp:164-178). */
int atm_set_parameters(atm_handle *h, int32_t replica, const double p[ATM_NUM_PARAMS]);
int atm_get_parameters(atm_handle *h, int32_t replica, double p[ATM_NUM_PARAMS]);

/* ------------------------------------------------------------------ Tier 1 */

/* posq1 = posq, posq2 = posq + displ (component-wise in `real`, .w gets + 0), corrections copied verbatim.
 * Element type per cfg.precision.  posq_corr/posq1_corr/posq2_corr may be NULL (single, double).
 * Only slots i < N are written.  One sweep, one launch.
 * ref: kernel CopyState (platforms/common/src/kernels/atmmetaforce.cc:19-52), launched from
 *      CommonCalcATMMetaForceKernel::copyState (CommonATMMetaForceKernels.cpp:206-212). */
int atm_copy_state(atm_handle *h, const void *posq, const void *posq_corr, void *posq1, void *posq1_corr,
                   void *posq2, void *posq2_corr, void *stream);

/* Optional extra (north_star "with periodic wrap"): posq2w = posq2 wrapped into [0,L) per axis of a
 * rectangular box; posq2 itself stays the bit-exact unwrapped value.  float4 only. */
int atm_wrap_positions(atm_handle *h, const void *posq_in, void *posq_out, const double box[9], void *stream);

/* force[c*P+i] += llrint(sp*f2[c*P+i] + (1-sp)*f1[c*P+i]), c = 0..2, i < N, on the 2^32 fixed-point
 * SoA buffers; blend in double.
 * ref: kernel HybridForce (platforms/common/src/kernels/atmmetaforce.cc:1-17). */
int atm_hybrid_force(atm_handle *h, int64_t *force, const int64_t *force_state1, const int64_t *force_state2,
                     double sp, void *stream);

/* Host scalar stage: out = {u_sc, fp, ebias, bfp, energy, sp, bfp*fp}.
 * ref: SoftCoreF + softplus in CommonCalcATMMetaForceKernel::execute (CommonATMMetaForceKernels.cpp:19-30,182-199). */
int atm_softcore_softplus(const double p[ATM_NUM_PARAMS], double state1_energy, double state2_energy, double out[7]);

/* The reference's execute(): scalar stage on the host from the replica's parameters, then the
 * hybrid-force launch.  *energy = includeEnergy ? e0 + ebias : 0; the perturbation energy is cached.
 * ref: CommonCalcATMMetaForceKernel::execute (CommonATMMetaForceKernels.cpp:154-204). */
int atm_execute(atm_handle *h, int32_t replica, double state1_energy, double state2_energy, int64_t *force,
                const int64_t *force_state1, const int64_t *force_state2, int32_t include_energy, double *energy,
                void *stream);

/* ref: CalcATMMetaForceKernel::getPerturbationEnergy (openmmapi/include/ATMMetaForceKernels.h:57). */
int atm_get_perturbation_energy(atm_handle *h, int32_t replica, double *u_sc);

/* ------------------------------------------------------------------ Tier 2 */

/* What the inner contexts evaluate for the variable force groups: the NonbondedForce direct space.
 * ref: ATMMetaForceImpl::copysystem / initialize (openmmapi/src/ATMMetaForceImpl.cpp:51-65,75-81) decide
 * which forces make up U1/U2; the arithmetic itself is OpenMM's NonbondedForce (external to the reference). */
typedef struct {
    const double *charge;            /* [N] e, by atom */
    const double *sigma;             /* [N] nm */
    const double *epsilon;           /* [N] kJ/mol */
    int32_t num_exclusions;
    const int32_t *exclusions;       /* [num_exclusions][2] atom pairs, each once */
    int32_t num_exceptions;
    const int32_t *exception_pairs;  /* [num_exceptions][2], subset of the exclusions */
    const double *exception_params;  /* [num_exceptions][3] chargeProd (e^2), sigma (nm), epsilon (kJ/mol) */
    double cutoff;                   /* nm */
    double ewald_alpha;              /* 1/nm; 0 = plain Coulomb inside the cutoff */
    double skin;                     /* nm, padding of the pruned (inner) pair list used every step */
    double skin_outer;               /* nm, padding of the outer list the inner one is pruned from (0 = same as skin) */
} atm_nonbonded_desc;

int atm_nb_setup(atm_handle *h, const atm_nonbonded_desc *desc, void *stream);

/* Optional: evaluate the PME reciprocal space of BOTH states inside atm_step as well (smooth PME,
 * B-splines of `order` 4..8 (OpenMM: 5), mesh nx*ny*nz with nz >= 8, cuFFT for the transforms; the environment charge
 * is spread once for the two states).  U1/U2 then contain direct + reciprocal + self energy, i.e. the complete
 * NonbondedForce of both inner contexts (the long-range dispersion correction: atm_nb_set_dispersion_correction).  nx = ny = nz = 0 switches it off.
 * Meshes and transforms are single precision, energies double (the environment variable ATM_B200_PME_F64=1, read at
 * the first set-up, selects the double-precision mesh pipeline instead).  The spread finds sites through the pair-list
 * structure of the last atm_nb_rebuild: a site that has since moved further than half the outer skin beyond its cluster
 * makes the step return NaN energies (like a pair-list overflow) until the next rebuild -- the same movement invalidates
 * the pair list itself.
 * Call after atm_nb_setup.  SYNCHRONISES the device. */
int atm_pme_setup(atm_handle *h, int32_t nx, int32_t ny, int32_t nz, int32_t order);

/* NonbondedForce::setUseDispersionCorrection (OpenMM default: on; ref: the inner contexts evaluate the cloned
 * NonbondedForce as is, openmmapi/src/ATMMetaForceImpl.cpp:51-65).  When on, the isotropic long-range Lennard-Jones
 * tail energy  8 pi N^2/V (<eps sig^12>/(9 rc^9) - <eps sig^6>/(3 rc^3))  is added to U1 and U2 by the scalar stage
 * (identical in both states, so u, W and the forces are unaffected; off by default here).  With PME on, the
 * neutralising-background term -pi Q^2/(2 V alpha^2) of a charged cell is always included.  Call after atm_nb_setup. */
int atm_nb_set_dispersion_correction(atm_handle *h, int32_t on);

/* Box vectors (row-major a,b,c in nm) of one replica (-1 = all).  Rectangular boxes only in this build.
 * ref: the box mirror in copyState (CommonATMMetaForceKernels.cpp:214-219). */
int atm_set_box(atm_handle *h, int32_t replica, const double box[9]);

/* (Re)builds the cluster pair lists of every replica for both coordinate states from posq ([R][P] float4, slot
 * order): spatial sort, clusters, the OUTER list (radius cutoff+skin_outer, with exclusion masks) and the pruned INNER
 * list (radius cutoff+skin).  Must be called before the first atm_step, whenever any atom has moved by more than
 * skin_outer/2 since the last rebuild, and after atm_set_displacements / reordering.
 * The first build after a (re)allocation SYNCHRONISES `stream` (it verifies and, if needed, grows the list
 * capacities).  Later builds on a non-default stream are fully asynchronous (one cached CUDA graph); their capacity
 * flags are copied to pinned host memory and inspected by the next API call that finds them complete -- a list that
 * outgrew its capacity then surfaces as ATM_ERR_STATE from that call (and is cured by calling atm_nb_rebuild again). */
int atm_nb_rebuild(atm_handle *h, const void *posq, void *stream);

/* Verification of the last asynchronous rebuild (wait != 0: block until its capacity flags have arrived).
 * ATM_OK: the lists are complete.  ATM_ERR_STATE: a list outgrew its capacity -- every step since that rebuild has
 * returned NaN energies and NaN-poisoned forces (never silently truncated sums); atm_nb_rebuild again reallocates
 * with larger capacities (synchronously) and the caller repeats those steps.  A list that merely came within 20 % of
 * its capacity makes the NEXT atm_nb_rebuild reallocate ahead of time; that is not an error.
 * With the PME option on, wait != 0 also synchronises the device and reports (ATM_ERR_STATE) when a site has moved
 * further since the last rebuild than the structure tolerates (see atm_pme_setup): the steps since then returned NaN.
 * No reference counterpart: OpenMM's own neighbour list handles this inside its inner contexts. */
int atm_nb_check(atm_handle *h, int32_t wait);

/* Re-prunes the INNER list from the outer one at the current coordinates (cheap, asynchronous, capturable).
 * Needed whenever any atom has moved by more than skin/2 since the last prune or rebuild. */
int atm_nb_prune(atm_handle *h, const void *posq, void *stream);

typedef struct {
    const void *posq;          /* [R][P] float4 positions+charge, slot order (borrowed) */
    const void *posq_corr;     /* [R][P] float4 or NULL */
    int64_t *force;            /* [R][3P] outer-context force buffers the merged force is ADDED to */
    /* optional per-state contributions computed elsewhere (e.g. OpenMM's PME reciprocal in the inner contexts): */
    const int64_t *force_state1_ext; /* [R][3P] or NULL */
    const int64_t *force_state2_ext; /* [R][3P] or NULL */
    const double *energy_ext;        /* device [R][2] {U1_ext, U2_ext} or NULL */
    /* optional inner-context coordinate outputs (exactly what atm_copy_state writes), or NULL: */
    void *posq1, *posq1_corr, *posq2, *posq2_corr;
    int32_t include_energy;    /* also evaluate the shared (env-env) pair energies so that U1,U2,E are valid;
                                  0 = forces and u only (u, u_sc, W, sp remain exact) */
    int32_t collect_stats;     /* also count the pairs inside the cutoff (ATM_E_NPAIRS); costs a few percent */
    int32_t concurrent_prune;  /* 1 = re-prune the inner list from THESE coordinates on the handle's side stream while
                                  the step runs on the list in use; `stream` joins at the end of the step and the NEXT
                                  step uses the new list (the inner skin must cover one more step than with
                                  atm_nb_prune).  Needs a non-default stream. */
    int32_t reserved;
} atm_step_io;

/* One pass of the hot path for all R replicas: copy-state -> two-state direct space -> device scalar
 * stage -> merge.  No host synchronisation; safe to capture into a CUDA graph.
 * ref: ATMMetaForceImpl::calcForcesAndEnergy (openmmapi/src/ATMMetaForceImpl.cpp:90-128). */
int atm_step(atm_handle *h, const atm_step_io *io, void *stream);

/* atm_step replayed from a cached CUDA graph (re-captured when buffers / flags / the pair list change).
 * `stream` must be a non-default stream. */
int atm_step_graph(atm_handle *h, const atm_step_io *io, void *stream);

/* Measurement hooks.  With profiling on, atm_step brackets the nb2 launch with CUDA events on the launching
 * stream; atm_profile_read synchronises those events and returns their summed time and count since the last read.
 * atm_launch_count: kernels of this library launched through the handle so far. */
int atm_profile_enable(atm_handle *h, int32_t on);
int atm_profile_read(atm_handle *h, double *nb2_ms_total, int32_t *nb2_launches);
int atm_launch_count(atm_handle *h, uint64_t *count);

/* Device pointer to the [R][ATM_NUM_ENERGY_SLOTS] energy records (valid after atm_step completes). */
int atm_energies_device(atm_handle *h, const double **dev_ptr);
/* Copies the records to the host; SYNCHRONISES `stream`. */
int atm_get_energies(atm_handle *h, double *out, void *stream);

/* Diagnostics of the last rebuild: out = {sites per replica, clusters of replica 0, list entries per replica (mean),
 * env list capacity, ligand/ghost list capacity, displaced atoms M, displacement groups G, xy columns}. */
int atm_nb_stats(atm_handle *h, int64_t out[8]);

/* ------------------------------------------------------------------ runtime utilities for hosts without the CUDA toolkit */

/* A C++ / Python host that drives the library through this header alone (no cuda_runtime.h) still needs a stream to
 * issue work on and page-locked memory for the host-buffer step.  Thin wrappers of cudaStreamCreateWithFlags
 * (non-blocking) / cudaStreamDestroy / cudaStreamSynchronize / cudaHostAlloc / cudaFreeHost on `device` (-1 = current).
 * The reference gets all of these from OpenMM's CudaContext (cu.getCurrentStream(), CudaContext::getPinnedBuffer). */
int atm_stream_create(int32_t device, void **stream);
int atm_stream_destroy(void *stream);
int atm_stream_synchronize(void *stream);
int atm_host_alloc(size_t bytes, void **ptr);
int atm_host_free(void *ptr);

/* ------------------------------------------------------------------ the step with HOST buffers on both sides */

/* The reference's callers move coordinates in and forces / energies out through host memory every time they touch a
 * Context (context.setPositions(...), context.getState(getEnergy=True, getForces=True); ref: python/tests/test_abfe.py:115-146,
 * example/abfe/abfe.py:140-160).  atm_host_pipeline_step is that round trip for the Tier-2 path as ONE call:
 *   per handle: H2D coordinates -> [rebuild | prune] -> atm_step -> D2H forces and energy records,
 * every handle on a stream of its own, forked from and joined back into `stream`, the whole fork/join replayed from one
 * cached CUDA graph per maintenance kind.  Several handles ("chunks" of the replicas resident on this GPU) overlap the
 * copies of one chunk with the kernels of the others.  The pipeline owns the device staging buffers; the caller needs
 * no device memory at all.  Handles must outlive the pipeline and must not be stepped concurrently through atm_step. */
typedef struct atm_host_pipeline atm_host_pipeline;

/* What comes back in force_host: the fixed-point long force buffer as OpenMM keeps it (24 B per atom), the same values as
 * float32 kJ/mol/nm (12 B per atom: half the PCIe traffic; a host caller converts to floating point anyway, as
 * Context.getState does), or nothing (a caller that only samples energies). */
enum { ATM_FORCE_I64 = 0, ATM_FORCE_F32 = 1, ATM_FORCE_NONE = 2 };

typedef struct {
    const void *posq_host;     /* [R][P] float4 (x, y, z, q), slot order; pinned host memory (cudaHostAlloc / cudaHostRegister) */
    void *force_host;          /* [R][3P] pinned host memory, SoA x|y|z blocks per replica: int64 (2^32 fixed point) or float32
                                  according to force_format; may be NULL with ATM_FORCE_NONE */
    double *energies_host;     /* [R][ATM_NUM_ENERGY_SLOTS] pinned host memory, or NULL */
    int32_t include_energy;    /* as atm_step_io.include_energy */
    int32_t force_format;      /* ATM_FORCE_I64 (0, default) | ATM_FORCE_F32 | ATM_FORCE_NONE */
    int32_t posq_format;       /* ATM_POSQ_F4 (0, default): posq_host is [R][P] float4 as above;
                                  ATM_POSQ_F3: posq_host is [R][P] packed float3 (x, y, z), 12 B per slot instead of 16 -- the
                                  direct-space path takes the charges from atm_nb_setup, never from posq.w */
    int32_t reserved;
    /* optional per-state contributions of OTHER forces of the variable force groups, evaluated by the caller at the state-1
     * and state-2 coordinates (the reference's inner contexts evaluate every Force in a variable group, ref:
     * openmmapi/src/ATMMetaForceImpl.cpp:51-65,113-116): uploaded and handed to the step as atm_step_io.force_state{1,2}_ext /
     * energy_ext.  Pinned host memory; NULL = none. */
    const int64_t *force_state1_ext_host;   /* [R][3P] 2^32 fixed point, SoA x|y|z blocks per replica */
    const int64_t *force_state2_ext_host;   /* [R][3P] */
    const double *energy_ext_host;          /* [R][2] {U1_ext, U2_ext} */
} atm_host_io;
enum { ATM_POSQ_F4 = 0, ATM_POSQ_F3 = 1 };

int atm_host_pipeline_create(int32_t num_handles, atm_handle *const *handles, atm_host_pipeline **out);
int atm_host_pipeline_destroy(atm_host_pipeline *p);
/* ios: one entry per handle, in the order given to _create.  maintenance: 0 = none, 1 = atm_nb_prune first,
 * 2 = atm_nb_rebuild first (the very first step must pass 2; that one synchronises, see atm_nb_rebuild),
 * 3 = re-prune CONCURRENTLY with this step (atm_step_io.concurrent_prune: the new list serves the next step).
 * Asynchronous: the host buffers hold the results once `stream` (non-default) has been synchronised. */
int atm_host_pipeline_step(atm_host_pipeline *p, const atm_host_io *ios, int32_t maintenance, void *stream);
/* atm_nb_check(wait = 1) over every handle of the pipeline: call after synchronising a maintenance = 2 step and, on
 * ATM_ERR_STATE, repeat that step (its rebuild then takes the synchronous, capacity-growing path). */
int atm_host_pipeline_check(atm_host_pipeline *p);

/* ------------------------------------------------------------------ Hamiltonian replica exchange (host) */

/* Deterministic Metropolis sweep over neighbouring lambda-states, identical on every rank.
 *   state_params  : [num_states][9]
 *   u12           : [num_replicas][2] {U1, U2} of every replica's current coordinates (all-gathered)
 *   replica_state : [num_replicas] in/out, state index held by each replica (a permutation when
 *                   num_replicas == num_states)
 *   beta          : 1/kT in mol/kJ;  (seed, cycle) key the counter-based RNG
 * No reference counterpart (the reference has no replica exchange). */
int atm_hrex_sweep(int32_t num_states, const double *state_params, int32_t num_replicas, const double *u12,
                   int32_t *replica_state, double beta, uint64_t seed, uint64_t cycle, int32_t *num_accepted);

/* The same exchange cycle WITHOUT a host round trip (no synchronisation, capturable): the replica -> state table lives
 * on the device (replicated on every rank), the sweep runs in one small kernel behind the all-gather, and the parameter
 * rows of the local replicas are rewritten in place.  Bit-identical decisions to atm_hrex_sweep.
 *   setup    : state_params [num_states][9]; replica_state [num_replicas] initial table; local_replicas [R] global id
 *              of each local replica (-1 = unused slot); gather_slot [num_replicas] row of replica g in the gathered
 *              array ([gathered_rows][2] doubles).  SYNCHRONISES `stream` (host arrays are copied).
 *   pack     : send[k] = {U1, U2} of local replica k (rows >= R are zero) -- the all-gather input.
 *   exchange : gathered = the all-gathered rows (device); cycle keys the RNG together with the seed.
 *   state    : copies the table and {accepted swaps, cycles, error flag (a non-finite energy skipped a cycle)} to the
 *              host; SYNCHRONISES `stream`.  atm_get_parameters / atm_set_parameters refresh the host mirror of the
 *              parameter rows from the device first. */
int atm_hrex_device_setup(atm_handle *h, int32_t num_states, const double *state_params, int32_t num_replicas,
                          const int32_t *replica_state, const int32_t *local_replicas, const int32_t *gather_slot,
                          int32_t gathered_rows, double beta, uint64_t seed, void *stream);
int atm_hrex_device_pack(atm_handle *h, double *send, int32_t rows, void *stream);
int atm_hrex_device_exchange(atm_handle *h, const double *gathered, uint64_t cycle, void *stream);
int atm_hrex_device_state(atm_handle *h, int32_t *replica_state, int64_t counters[3], void *stream);

/* The communicator of the replica layer (no reference counterpart): one rank per GPU of a node,
 * NCCL over NVLink.  NCCL is resolved at RUN time (the libnccl.so.2 the process already carries, e.g. torch's, else the
 * system library); a build without NCCL still loads and only these entry points report ATM_ERR_UNSUPPORTED.
 *   atm_re_unique_id      rank 0 fills 128 bytes (an ncclUniqueId) and shares them by any side channel (MPI, a file,
 *                         torch.distributed.broadcast_object_list ...)
 *   atm_re_comm_create    every rank, collectively: ncclCommInitRank.  world == 1 needs neither id nor NCCL.
 *   atm_re_comm_from_nccl borrows an ncclComm_t the host created itself (not destroyed by _destroy)
 *   atm_hrex_device_cycle one whole exchange cycle on `stream`: pack kernel -> ncclAllGather of 2 doubles per replica
 *                         slot -> sweep kernel (new lambda states, parameter rows rewritten in place).  No host
 *                         synchronisation; capturable into a CUDA graph; every rank must call it with the same cycle. */
typedef struct atm_re_comm atm_re_comm;
int atm_re_unique_id(void *id128);
int atm_re_comm_create(const void *id128, int32_t world, int32_t rank, int32_t device, atm_re_comm **out);
int atm_re_comm_from_nccl(void *nccl_comm, int32_t rank, atm_re_comm **out);
int atm_re_comm_destroy(atm_re_comm *c);
int atm_hrex_device_cycle(atm_handle *h, atm_re_comm *comm, uint64_t cycle, void *stream);

/* Reduced energy beta*E_s(x) of coordinates with energies (U1,U2) under state parameters p (host). */
double atm_hrex_reduced_energy(const double p[ATM_NUM_PARAMS], double U1, double U2, double beta);

#ifdef __cplusplus
}
#endif
#endif /* ATM_B200_H_ */
