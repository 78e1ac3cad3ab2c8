// atm_capi.cu -- C ABI (include/atm_b200.h): handle life cycle, displacement table, parameters, the Tier-1
// entry points and the host-side replica-exchange decision.  Tier-2 entry points live in atm_nb.cu.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "atm_common.cuh"

namespace atm {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int upload_params_if_dirty(atm_handle *h, cudaStream_t stream) {
    if (!h->params_dirty) return ATM_OK;
    // params live in pageable host memory owned by the handle: cudaMemcpyAsync stages it before returning.
    ATM_CUDA_CHECK(cudaMemcpyAsync(h->d_params, h->params.data(), sizeof(double) * h->params.size(),
                                   cudaMemcpyHostToDevice, stream));
    h->params_dirty = false;
    return ATM_OK;
}

// The on-device replica exchange rewrites parameter rows behind the host mirror's back; host reads / edits pull them
// back first (synchronises the device: rare, bookkeeping only).
int refresh_params_from_device(atm_handle *h) {
    if (!h->params_device_newer) return ATM_OK;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    ATM_CUDA_CHECK(cudaDeviceSynchronize());
    ATM_CUDA_CHECK(cudaMemcpy(h->params.data(), h->d_params, sizeof(double) * h->params.size(), cudaMemcpyDeviceToHost));
    h->params_device_newer = false;
    return ATM_OK;
}

}  // namespace atm

using namespace atm;

extern "C" {

const char *atm_last_error(void) { return g_error; }
const char *atm_version(void) { return ATM_B200_VERSION; }

int atm_create(const atm_config *cfg, atm_handle **out) {
    ATM_REQUIRE(cfg != nullptr && out != nullptr, ATM_ERR_INVALID, "atm_create: null argument");
    ATM_REQUIRE(cfg->num_particles >= 0, ATM_ERR_INVALID, "atm_create: num_particles < 0");
    ATM_REQUIRE(cfg->precision >= ATM_PREC_SINGLE && cfg->precision <= ATM_PREC_DOUBLE, ATM_ERR_INVALID,
                "atm_create: unknown precision %d", cfg->precision);
    ATM_REQUIRE(cfg->num_replicas >= 1, ATM_ERR_INVALID, "atm_create: num_replicas must be >= 1");
    int P = cfg->padded_num_particles;
    if (P == 0) P = 32 * ((cfg->num_particles + 31) / 32);
    ATM_REQUIRE(P >= cfg->num_particles, ATM_ERR_INVALID, "atm_create: padded_num_particles %d < num_particles %d", P,
                cfg->num_particles);
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0) {
        set_error("atm_create: no CUDA device available (%s); this back-end has no CPU fallback",
                  err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
        return ATM_ERR_CUDA;
    }
    int dev = cfg->device;
    if (dev < 0) ATM_CUDA_CHECK(cudaGetDevice(&dev));
    ATM_REQUIRE(dev < ndev, ATM_ERR_INVALID, "atm_create: device %d out of range (%d devices)", dev, ndev);
    ATM_CUDA_CHECK(cudaSetDevice(dev));

    atm_handle *h = new (std::nothrow) atm_handle();
    ATM_REQUIRE(h != nullptr, ATM_ERR_NOMEM, "atm_create: out of host memory");
    h->cfg = *cfg;
    h->cfg.padded_num_particles = P;
    h->N = cfg->num_particles;
    h->P = P;
    h->R = cfg->num_replicas;
    h->device = dev;
    h->d_displ = nullptr;
    h->d_params = nullptr;
    h->params_device_newer = false;
    h->hrex = nullptr;
    h->have_displ = false;
    h->launches = 0;
    h->nb = nullptr;
    cudaDeviceProp prop;
    if ((err = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess) {
        set_error("atm_create: cudaGetDeviceProperties failed: %s", cudaGetErrorString(err));
        atm_destroy(h);   // every failure after `new` releases the handle
        return ATM_ERR_CUDA;
    }
    h->num_sms = prop.multiProcessorCount;
    h->params.assign((size_t)h->R * ATM_NUM_PARAMS, 0.0);
    for (int r = 0; r < h->R; r++) h->params[(size_t)r * ATM_NUM_PARAMS + ATM_DIRECTION] = 1.0;
    h->params_dirty = true;
    h->pert_energy.assign(h->R, 0.0);
    if (cudaMalloc(&h->d_displ, sizeof(float4) * (size_t)std::max(P, 1)) != cudaSuccess ||
        cudaMalloc(&h->d_params, sizeof(double) * h->params.size()) != cudaSuccess) {
        set_error("atm_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        atm_destroy(h);
        return ATM_ERR_CUDA;
    }
    if ((err = cudaMemset(h->d_displ, 0, sizeof(float4) * (size_t)std::max(P, 1))) != cudaSuccess) {
        set_error("atm_create: cudaMemset failed: %s", cudaGetErrorString(err));
        atm_destroy(h);
        return ATM_ERR_CUDA;
    }
    h->atom_index.resize(h->N);
    for (int i = 0; i < h->N; i++) h->atom_index[i] = i;
    h->displ_by_atom.assign((size_t)h->N * 3, 0.0);
    *out = h;
    return ATM_OK;
}

int atm_destroy(atm_handle *h) {
    if (h == nullptr) return ATM_OK;
    cudaSetDevice(h->device);
    nb_destroy(h);
    hrex_destroy(h);
    if (h->d_displ) cudaFree(h->d_displ);
    if (h->d_params) cudaFree(h->d_params);
    delete h;
    return ATM_OK;
}

int atm_set_displacements(atm_handle *h, const int32_t *atom_index, const double *dxyz, void *stream) {
    ATM_NVTX_RANGE("atm_set_displacements");
    ATM_REQUIRE(h != nullptr && (dxyz != nullptr || h->N == 0), ATM_ERR_INVALID, "atm_set_displacements: null argument");
    const int N = h->N, P = h->P;
    if (atom_index != nullptr) {
        std::vector<char> seen(N, 0);
        for (int s = 0; s < N; s++) {
            int a = atom_index[s];
            ATM_REQUIRE(a >= 0 && a < N && !seen[a], ATM_ERR_INVALID,
                        "atm_set_displacements: atom_index is not a permutation (slot %d -> %d)", s, a);
            seen[a] = 1;
        }
        std::copy(atom_index, atom_index + N, h->atom_index.begin());
    } else {
        for (int i = 0; i < N; i++) h->atom_index[i] = i;
    }
    if (N > 0) std::copy(dxyz, dxyz + (size_t)3 * N, h->displ_by_atom.begin());
    // float4 table in slot order, zero padded; d rounded double -> float exactly once.
    h->h_displ.assign((size_t)4 * P, 0.0f);
    for (int s = 0; s < N; s++) {
        const int a = h->atom_index[s];
        h->h_displ[4 * (size_t)s + 0] = (float)dxyz[3 * (size_t)a + 0];
        h->h_displ[4 * (size_t)s + 1] = (float)dxyz[3 * (size_t)a + 1];
        h->h_displ[4 * (size_t)s + 2] = (float)dxyz[3 * (size_t)a + 2];
    }
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    if (P > 0)
        ATM_CUDA_CHECK(cudaMemcpyAsync(h->d_displ, h->h_displ.data(), sizeof(float) * 4 * (size_t)P,
                                       cudaMemcpyHostToDevice, (cudaStream_t)stream));
    h->have_displ = true;
    return nb_on_displacements_changed(h, (cudaStream_t)stream);
}

int atm_set_parameters(atm_handle *h, int32_t replica, const double p[ATM_NUM_PARAMS]) {
    ATM_REQUIRE(h != nullptr && p != nullptr, ATM_ERR_INVALID, "atm_set_parameters: null argument");
    ATM_REQUIRE(replica >= -1 && replica < h->R, ATM_ERR_INVALID, "atm_set_parameters: replica %d out of range", replica);
    int rc0 = refresh_params_from_device(h);
    if (rc0) return rc0;
    for (int r = 0; r < h->R; r++)
        if (replica < 0 || replica == r) std::copy(p, p + ATM_NUM_PARAMS, h->params.begin() + (size_t)r * ATM_NUM_PARAMS);
    h->params_dirty = true;
    return ATM_OK;
}

int atm_get_parameters(atm_handle *h, int32_t replica, double p[ATM_NUM_PARAMS]) {
    ATM_REQUIRE(h != nullptr && p != nullptr, ATM_ERR_INVALID, "atm_get_parameters: null argument");
    ATM_REQUIRE(replica >= 0 && replica < h->R, ATM_ERR_INVALID, "atm_get_parameters: replica %d out of range", replica);
    int rc0 = refresh_params_from_device(h);
    if (rc0) return rc0;
    std::copy(h->params.begin() + (size_t)replica * ATM_NUM_PARAMS,
              h->params.begin() + (size_t)(replica + 1) * ATM_NUM_PARAMS, p);
    return ATM_OK;
}

int atm_copy_state(atm_handle *h, const void *posq, const void *posq_corr, void *posq1, void *posq1_corr, void *posq2,
                   void *posq2_corr, void *stream) {
    ATM_NVTX_RANGE("atm_copy_state");
    ATM_REQUIRE(h != nullptr, ATM_ERR_INVALID, "atm_copy_state: null handle");
    if (h->N == 0) return ATM_OK;
    ATM_REQUIRE(posq && posq1 && posq2, ATM_ERR_INVALID, "atm_copy_state: null position buffer");
    ATM_REQUIRE((posq_corr == nullptr) == (posq1_corr == nullptr) && (posq_corr == nullptr) == (posq2_corr == nullptr),
                ATM_ERR_INVALID, "atm_copy_state: correction buffers must be all given or all NULL");
    ATM_REQUIRE(!(h->cfg.precision == ATM_PREC_MIXED && posq_corr == nullptr), ATM_ERR_INVALID,
                "atm_copy_state: mixed precision needs the posqCorrection buffers");
    ATM_REQUIRE(!(h->cfg.precision == ATM_PREC_DOUBLE && posq_corr != nullptr), ATM_ERR_INVALID,
                "atm_copy_state: double precision has no posqCorrection buffers");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    if (h->N > 0) h->launches++;
    return launch_copy_state(h, posq, posq_corr, posq1, posq1_corr, posq2, posq2_corr, (cudaStream_t)stream);
}

int atm_wrap_positions(atm_handle *h, const void *posq_in, void *posq_out, const double box[9], void *stream) {
    ATM_REQUIRE(h && posq_in && posq_out && box, ATM_ERR_INVALID, "atm_wrap_positions: null argument");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    h->launches++;
    return launch_wrap(h, posq_in, posq_out, box, (cudaStream_t)stream);
}

int atm_hybrid_force(atm_handle *h, int64_t *force, const int64_t *f1, const int64_t *f2, double sp, void *stream) {
    ATM_NVTX_RANGE("atm_hybrid_force");
    ATM_REQUIRE(h != nullptr, ATM_ERR_INVALID, "atm_hybrid_force: null handle");
    if (h->N == 0) return ATM_OK;
    ATM_REQUIRE(force && f1 && f2, ATM_ERR_INVALID, "atm_hybrid_force: null force buffer");
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    h->launches++;
    return launch_hybrid_force(h, force, f1, f2, sp, (cudaStream_t)stream);
}

int atm_softcore_softplus(const double p[ATM_NUM_PARAMS], double U1, double U2, double out[7]) {
    ATM_REQUIRE(p && out, ATM_ERR_INVALID, "atm_softcore_softplus: null argument");
    Scalars s = scalar_stage(p, U1, U2, U2 - U1);
    out[0] = s.usc; out[1] = s.fp; out[2] = s.ebias; out[3] = s.bfp; out[4] = s.energy; out[5] = s.sp;
    out[6] = s.bfp * s.fp;
    return ATM_OK;
}

int atm_execute(atm_handle *h, int32_t replica, double U1, double U2, int64_t *force, const int64_t *f1,
                const int64_t *f2, int32_t include_energy, double *energy, void *stream) {
    ATM_NVTX_RANGE("atm_execute");
    ATM_REQUIRE(h != nullptr, ATM_ERR_INVALID, "atm_execute: null handle");
    ATM_REQUIRE(replica >= 0 && replica < h->R, ATM_ERR_INVALID, "atm_execute: replica %d out of range", replica);
    int rc0 = refresh_params_from_device(h);  // the on-device exchange may have rewritten the rows
    if (rc0) return rc0;
    Scalars s = scalar_stage(h->params.data() + (size_t)replica * ATM_NUM_PARAMS, U1, U2, U2 - U1);
    h->pert_energy[replica] = s.usc;
    if (energy) *energy = include_energy ? s.energy : 0.0;
    return atm_hybrid_force(h, force, f1, f2, s.sp, stream);
}

int atm_get_perturbation_energy(atm_handle *h, int32_t replica, double *u_sc) {
    ATM_REQUIRE(h && u_sc, ATM_ERR_INVALID, "atm_get_perturbation_energy: null argument");
    ATM_REQUIRE(replica >= 0 && replica < h->R, ATM_ERR_INVALID, "atm_get_perturbation_energy: replica out of range");
    *u_sc = h->pert_energy[replica];
    return ATM_OK;
}

// ------------------------------------------------------------------ replica exchange (host, deterministic)

double atm_hrex_reduced_energy(const double p[ATM_NUM_PARAMS], double U1, double U2, double beta) {
    Scalars s = scalar_stage(p, U1, U2, U2 - U1);
    return beta * s.energy;
}

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int atm_hrex_sweep(int32_t num_states, const double *state_params, int32_t num_replicas, const double *u12,
                   int32_t *replica_state, double beta, uint64_t seed, uint64_t cycle, int32_t *num_accepted) {
    ATM_REQUIRE(state_params && u12 && replica_state, ATM_ERR_INVALID, "atm_hrex_sweep: null argument");
    ATM_REQUIRE(num_states >= 1 && num_replicas >= 1, ATM_ERR_INVALID, "atm_hrex_sweep: empty problem");
    std::vector<int> holder(num_states, -1);  // state -> replica (first holder)
    for (int r = 0; r < num_replicas; r++) {
        int s = replica_state[r];
        ATM_REQUIRE(s >= 0 && s < num_states, ATM_ERR_INVALID, "atm_hrex_sweep: replica %d holds invalid state %d", r, s);
        for (int c = 0; c < 2; c++)
            ATM_REQUIRE(isfinite(u12[2 * r + c]), ATM_ERR_INVALID,
                        "atm_hrex_sweep: non-finite energy for replica %d (U%d)", r, c + 1);
        if (holder[s] < 0) holder[s] = r;
    }
    int accepted = 0;
    // neighbouring states (k, k+1), k of alternating parity per cycle
    for (int k = (int)(cycle & 1); k + 1 < num_states; k += 2) {
        int a = holder[k], b = holder[k + 1];
        if (a < 0 || b < 0) continue;
        const double *pk = state_params + (size_t)k * ATM_NUM_PARAMS, *pk1 = state_params + (size_t)(k + 1) * ATM_NUM_PARAMS;
        double e_ka = atm_hrex_reduced_energy(pk, u12[2 * a], u12[2 * a + 1], beta);
        double e_k1b = atm_hrex_reduced_energy(pk1, u12[2 * b], u12[2 * b + 1], beta);
        double e_kb = atm_hrex_reduced_energy(pk, u12[2 * b], u12[2 * b + 1], beta);
        double e_k1a = atm_hrex_reduced_energy(pk1, u12[2 * a], u12[2 * a + 1], beta);
        double delta = (e_kb + e_k1a) - (e_ka + e_k1b);
        uint64_t bits = splitmix64(splitmix64(seed ^ splitmix64(cycle)) + (uint64_t)k);
        double rnd = (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
        bool accept = !(delta > 0.0) || rnd < exp(-delta);
        if (delta != delta) accept = false;  // NaN guard
        if (accept) {
            replica_state[a] = k + 1;
            replica_state[b] = k;
            accepted++;
        }
    }
    if (num_accepted) *num_accepted = accepted;
    return ATM_OK;
}

}  // extern "C"
