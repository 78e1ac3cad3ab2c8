// atm_common.cuh -- shared declarations of the Blackwell ATM back-end (handle, error plumbing, scalar stage).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <string>
#include <vector>

#include "atm_b200.h"

#include <nvtx3/nvToolsExt.h>

namespace atm {

void set_error(const char *fmt, ...);

#define ATM_CUDA_CHECK(call)                                                                     \
    do {                                                                                         \
        cudaError_t err__ = (call);                                                              \
        if (err__ != cudaSuccess) {                                                              \
            atm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return ATM_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define ATM_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            atm::set_error(__VA_ARGS__); \
            return code;                 \
        }                                \
    } while (0)

constexpr double ONE_4PI_EPS0 = 138.935456;  // kJ nm / (mol e^2), OpenMM 7.x
constexpr double FORCE_SCALE = 4294967296.0; // 2^32 fixed point of OpenMM's long force buffers

// ---------------------------------------------------------------------------------------------
// Scalar stage (soft-core + softplus), identical on host and device, double precision.
// Semantics: reference CommonATMMetaForceKernels.cpp:19-30 (SoftCoreF) and :182-199.
// ---------------------------------------------------------------------------------------------
struct Scalars {
    double u, e0, usc, fp, ebias, bfp, energy, sp;
};

__host__ __device__ inline double softcore(double u, double umax, double a, double ub, double &fp) {
    if (u <= ub) {
        fp = 1.0;
        return u;
    }
    double g = (u - ub) / (a * (umax - ub));
    double zeta = 1.0 + 2.0 * g * (g + 1.0);
    double z = pow(zeta, a);
    double s = 4.0 * (2.0 * g + 1.0) / zeta;
    fp = s * z / ((1.0 + z) * (1.0 + z));
    return (umax - ub) * (z - 1.0) / (z + 1.0) + ub;
}

// du = U2 - U1 is passed separately so that the fused path can form it from the state-specific pairs only.
__host__ __device__ inline Scalars scalar_stage(const double *p, double U1, double U2, double du) {
    Scalars s;
    const double dir = p[ATM_DIRECTION];
    s.u = dir > 0 ? du : -du;
    s.e0 = dir > 0 ? U1 : U2;
    s.usc = softcore(s.u, p[ATM_UMAX], p[ATM_ACORE], p[ATM_UBCORE], s.fp);
    const double alpha = p[ATM_ALPHA];
    double ee = 1.0 + exp(-alpha * (s.usc - p[ATM_U0]));
    s.ebias = 0.0;
    if (alpha > 0) s.ebias = ((p[ATM_LAMBDA2] - p[ATM_LAMBDA1]) / alpha) * log(ee);
    s.ebias += p[ATM_LAMBDA2] * s.usc + p[ATM_W0];
    s.bfp = (p[ATM_LAMBDA2] - p[ATM_LAMBDA1]) / ee + p[ATM_LAMBDA1];
    s.energy = s.e0 + s.ebias;
    s.sp = dir > 0 ? s.bfp * s.fp : 1.0 - s.bfp * s.fp;
    return s;
}

// NVTX range around an entry point (host side; a no-op unless a profiler injects the NVTX library): the calls show up
// as named spans above their kernels in a timeline (tracing).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};
#define ATM_NVTX_RANGE(name) atm::NvtxRange atm_nvtx_range__(name)

struct NbState;  // Tier-2 state (atm_nb.cu)

}  // namespace atm

struct atm_handle {
    atm_config cfg;
    int N, P, R;
    int device;
    int num_sms;
    // displacement table, slot order, float4[P] (shared by all replicas)
    float4 *d_displ;
    std::vector<float> h_displ;        // staging (float4 per slot)
    std::vector<int32_t> atom_index;   // slot -> atom
    std::vector<double> displ_by_atom; // [N][3]
    bool have_displ;
    // parameters [R][9] host + device mirror
    std::vector<double> params;
    double *d_params;
    bool params_dirty;
    bool params_device_newer;          // the device rows were rewritten by the on-device exchange: refresh the host mirror first
    void *hrex;                        // atm::HrexState (atm_hrex.cu) or NULL
    std::vector<double> pert_energy;   // cached u_sc per replica (Tier-1 execute)
    uint64_t launches;                 // kernels of this library launched through this handle
    atm::NbState *nb;
};

namespace atm {
// Tier-2 hooks implemented in atm_nb.cu
void nb_destroy(atm_handle *h);
int nb_on_displacements_changed(atm_handle *h, cudaStream_t stream);
int upload_params_if_dirty(atm_handle *h, cudaStream_t stream);
int refresh_params_from_device(atm_handle *h);
void hrex_destroy(atm_handle *h);

// hooks of the host-buffer pipeline (atm_host.cu), implemented in atm_nb.cu
int nb_host_prepare(atm_handle *h, int maintenance, cudaStream_t stream, bool *needs_sync_rebuild);
int nb_host_enqueue(atm_handle *h, const void *posq, long long *force, int include_energy, int maintenance, cudaStream_t stream,
                    const long long *force_state1_ext = nullptr, const long long *force_state2_ext = nullptr,
                    const double *energy_ext = nullptr);
int nb_host_rebuild_enqueued(atm_handle *h, cudaStream_t stream);
int nb_host_inner_copy(const atm_handle *h);     // which copy of the pruned list is in use (0 / 1)
void nb_host_flip_inner(atm_handle *h);          // after a maintenance = 3 step has been enqueued / replayed
uint64_t nb_alloc_generation(const atm_handle *h);
const double *nb_energies_device(const atm_handle *h);

// Tier-1 launchers implemented in atm_copy_merge.cu
int launch_copy_state(atm_handle *h, const void *posq, const void *corr, void *posq1, void *corr1, void *posq2,
                      void *corr2, cudaStream_t stream);
int launch_hybrid_force(atm_handle *h, int64_t *force, const int64_t *f1, const int64_t *f2, double sp,
                        cudaStream_t stream);
int launch_wrap(atm_handle *h, const void *in, void *out, const double box[9], cudaStream_t stream);
}  // namespace atm
