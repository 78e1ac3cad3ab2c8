// atm_hrex.cu -- Hamiltonian replica exchange without a host round trip.
//
// The exchange step of the replica layer (the reference has no counterpart) needs, per cycle, the
// (U1, U2) of EVERY replica on EVERY rank, one Metropolis sweep over neighbouring lambda states that all ranks repeat
// bit-identically (counter-based RNG), and the new parameter rows of the local replicas.  Done on the host this costs a
// device synchronisation, a D2H copy, the sweep and an H2D upload per cycle -- about a millisecond of idle GPU, 9 % of
// the step time at 3 replicas per GPU.  Here the whole cycle stays on the stream:
//
//   hrex_pack_kernel      local (U1, U2) -> the all-gather send buffer            (device, async)
//   ncclAllGather         issued by the caller (torch.distributed) on the same stream
//   hrex_exchange_kernel  one block: sweep over all states, replica->state table updated in place (replicated on every
//                         rank), parameter rows of the local replicas rewritten in the handle's device parameter block
//
// The host reads the bookkeeping (state permutation, acceptance count) only when it asks for it.
// The sweep is the same function the host entry point atm_hrex_sweep runs (atm_capi.cu); tests compare the two.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "atm_common.cuh"

namespace atm {

struct HrexState {
    int num_states = 0, num_replicas = 0, gathered_rows = 0;
    double beta = 0.0;
    uint64_t seed = 0;
    double *d_schedule = nullptr;    // [num_states][9]
    int *d_replica_state = nullptr;  // [num_replicas], replicated on every rank
    int *d_gather_slot = nullptr;    // [num_replicas] row of replica g in the gathered array
    int *d_local_global = nullptr;   // [R] global replica id of local replica k, -1 = unused slot
    long long *d_counters = nullptr; // accepted swaps, cycles, error flag (non-finite energy)
    double *d_send = nullptr;        // [rows_per_rank][2] (atm_hrex_device_cycle)
    double *d_gathered = nullptr;    // [gathered_rows][2]
};

// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time: the library has no link-time dependency on it.  A process that already carries a
// libnccl.so.2 (e.g. through torch) keeps using THAT instance; otherwise the system library is loaded.  Only the five
// entry points below are used; their signatures are NCCL's public C API (nccl.h), with ncclComm_t / ncclUniqueId
// treated as an opaque pointer / 128 opaque bytes.
// ------------------------------------------------------------------------------------------------
struct NcclId128 { char b[128]; };   // ncclUniqueId is passed BY VALUE to ncclCommInitRank
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *id) = nullptr;
    int (*CommInitRank)(void **comm, int nranks, NcclId128 id, int rank) = nullptr;
    int (*CommDestroy)(void *comm) = nullptr;
    int (*CommCount)(void *comm, int *count) = nullptr;
    int (*AllGather)(const void *send, void *recv, size_t count, int dtype, void *comm, cudaStream_t stream) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

static NcclApi &nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the instance the process already uses, if any
        if (!a.lib) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return a;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
        a.CommCount = reinterpret_cast<decltype(a.CommCount)>(dlsym(a.lib, "ncclCommCount"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommCount && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble / ncclFloat64 in nccl.h

__host__ __device__ inline uint64_t hrex_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void hrex_pack_kernel(const double *__restrict__ energies, int R, double *__restrict__ send, int rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rows) return;
    send[2 * k] = k < R ? energies[(size_t)k * ATM_NUM_ENERGY_SLOTS + ATM_E_U1] : 0.0;
    send[2 * k + 1] = k < R ? energies[(size_t)k * ATM_NUM_ENERGY_SLOTS + ATM_E_U2] : 0.0;
}

constexpr int HREX_THREADS = 128;

__global__ void __launch_bounds__(HREX_THREADS)
hrex_exchange_kernel(int num_states, int num_replicas, const double *__restrict__ schedule, int *__restrict__ replica_state,
                     const int *__restrict__ gather_slot, const double *__restrict__ gathered, double beta, uint64_t seed,
                     uint64_t cycle, const int *__restrict__ local_global, int R, double *__restrict__ params,
                     long long *__restrict__ counters) {
    extern __shared__ int s_holder[];  // state -> first replica holding it
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    for (int s = tid; s < num_states; s += HREX_THREADS) s_holder[s] = 0x7fffffff;
    __syncthreads();
    for (int r = tid; r < num_replicas; r += HREX_THREADS) {
        const double u1 = gathered[2 * gather_slot[r]], u2 = gathered[2 * gather_slot[r] + 1];
        const int s = replica_state[r];
        if (!isfinite(u1) || !isfinite(u2) || s < 0 || s >= num_states) s_bad = 1;
        else atomicMin(&s_holder[s], r);
    }
    __syncthreads();
    if (s_bad) {  // leave everything as it is; the host sees the flag at its next read
        if (tid == 0) { counters[2] = 1; counters[1] += 1; }
        return;
    }
    // neighbouring states (k, k+1), k of alternating parity per cycle: the pairs of one cycle are disjoint
    for (int k = (int)(cycle & 1) + 2 * tid; k + 1 < num_states; k += 2 * HREX_THREADS) {
        const int a = s_holder[k], b = s_holder[k + 1];
        if (a == 0x7fffffff || b == 0x7fffffff) continue;
        const double *pk = schedule + (size_t)k * ATM_NUM_PARAMS, *pk1 = pk + ATM_NUM_PARAMS;
        const double ua1 = gathered[2 * gather_slot[a]], ua2 = gathered[2 * gather_slot[a] + 1];
        const double ub1 = gathered[2 * gather_slot[b]], ub2 = gathered[2 * gather_slot[b] + 1];
        const double e_ka = beta * scalar_stage(pk, ua1, ua2, ua2 - ua1).energy;
        const double e_k1b = beta * scalar_stage(pk1, ub1, ub2, ub2 - ub1).energy;
        const double e_kb = beta * scalar_stage(pk, ub1, ub2, ub2 - ub1).energy;
        const double e_k1a = beta * scalar_stage(pk1, ua1, ua2, ua2 - ua1).energy;
        const double delta = (e_kb + e_k1a) - (e_ka + e_k1b);
        const uint64_t bits = hrex_splitmix64(hrex_splitmix64(seed ^ hrex_splitmix64(cycle)) + (uint64_t)k);
        const double rnd = (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
        bool accept = !(delta > 0.0) || rnd < exp(-delta);
        if (delta != delta) accept = false;
        if (accept) {
            replica_state[a] = k + 1;
            replica_state[b] = k;
            atomicAdd((unsigned long long *)&counters[0], 1ull);
        }
    }
    __syncthreads();
    if (tid == 0) counters[1] += 1;
    // parameter rows of the replicas resident on this rank
    for (int t = tid; t < R * ATM_NUM_PARAMS; t += HREX_THREADS) {
        const int k = t / ATM_NUM_PARAMS, c = t - k * ATM_NUM_PARAMS;
        const int g = local_global[k];
        if (g >= 0) params[t] = schedule[(size_t)replica_state[g] * ATM_NUM_PARAMS + c];
    }
}

void hrex_destroy(atm_handle *h) {
    HrexState *x = (HrexState *)h->hrex;
    if (!x) return;
    cudaFree(x->d_schedule); cudaFree(x->d_replica_state); cudaFree(x->d_gather_slot); cudaFree(x->d_local_global);
    cudaFree(x->d_counters);
    cudaFree(x->d_send); cudaFree(x->d_gathered);
    delete x;
    h->hrex = nullptr;
}

}  // namespace atm

using namespace atm;

extern "C" {

int atm_hrex_device_setup(atm_handle *h, int32_t num_states, const double *state_params, int32_t num_replicas,
                          const int32_t *replica_state, const int32_t *local_replicas, const int32_t *gather_slot,
                          int32_t gathered_rows, double beta, uint64_t seed, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && state_params && replica_state && local_replicas && gather_slot, ATM_ERR_INVALID, "atm_hrex_device_setup: null argument");
    ATM_REQUIRE(num_states >= 1 && num_replicas >= 1 && gathered_rows >= 1, ATM_ERR_INVALID, "atm_hrex_device_setup: empty problem");
    ATM_REQUIRE(num_states <= 8192, ATM_ERR_UNSUPPORTED, "atm_hrex_device_setup: more than 8192 states");
    for (int g = 0; g < num_replicas; g++) {
        ATM_REQUIRE(replica_state[g] >= 0 && replica_state[g] < num_states, ATM_ERR_INVALID,
                    "atm_hrex_device_setup: replica %d holds invalid state %d", g, replica_state[g]);
        ATM_REQUIRE(gather_slot[g] >= 0 && gather_slot[g] < gathered_rows, ATM_ERR_INVALID, "atm_hrex_device_setup: bad gather slot of replica %d", g);
    }
    for (int k = 0; k < h->R; k++)
        ATM_REQUIRE(local_replicas[k] >= -1 && local_replicas[k] < num_replicas, ATM_ERR_INVALID, "atm_hrex_device_setup: bad local replica id %d", local_replicas[k]);
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    hrex_destroy(h);
    HrexState *x = new HrexState();
    h->hrex = x;
    x->num_states = num_states; x->num_replicas = num_replicas; x->gathered_rows = gathered_rows;
    x->beta = beta; x->seed = seed;
    ATM_CUDA_CHECK(cudaMalloc(&x->d_schedule, sizeof(double) * num_states * ATM_NUM_PARAMS));
    ATM_CUDA_CHECK(cudaMalloc(&x->d_replica_state, sizeof(int) * num_replicas));
    ATM_CUDA_CHECK(cudaMalloc(&x->d_gather_slot, sizeof(int) * num_replicas));
    ATM_CUDA_CHECK(cudaMalloc(&x->d_local_global, sizeof(int) * h->R));
    ATM_CUDA_CHECK(cudaMalloc(&x->d_counters, sizeof(long long) * 4));
    ATM_CUDA_CHECK(cudaMalloc(&x->d_send, sizeof(double) * 2 * gathered_rows));
    ATM_CUDA_CHECK(cudaMalloc(&x->d_gathered, sizeof(double) * 2 * gathered_rows));
    ATM_CUDA_CHECK(cudaMemsetAsync(x->d_send, 0, sizeof(double) * 2 * gathered_rows, stream));
    ATM_CUDA_CHECK(cudaMemcpyAsync(x->d_schedule, state_params, sizeof(double) * num_states * ATM_NUM_PARAMS, cudaMemcpyHostToDevice, stream));
    ATM_CUDA_CHECK(cudaMemcpyAsync(x->d_replica_state, replica_state, sizeof(int) * num_replicas, cudaMemcpyHostToDevice, stream));
    ATM_CUDA_CHECK(cudaMemcpyAsync(x->d_gather_slot, gather_slot, sizeof(int) * num_replicas, cudaMemcpyHostToDevice, stream));
    ATM_CUDA_CHECK(cudaMemcpyAsync(x->d_local_global, local_replicas, sizeof(int) * h->R, cudaMemcpyHostToDevice, stream));
    ATM_CUDA_CHECK(cudaMemsetAsync(x->d_counters, 0, sizeof(long long) * 4, stream));
    ATM_CUDA_CHECK(cudaStreamSynchronize(stream));  // the host arrays are the caller's
    return ATM_OK;
}

int atm_hrex_device_pack(atm_handle *h, double *send, int32_t rows, void *stream_) {
    ATM_NVTX_RANGE("atm_hrex_device_pack");
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && send && rows >= 1, ATM_ERR_INVALID, "atm_hrex_device_pack: bad argument");
    const double *en = nullptr;
    int rc = atm_energies_device(h, &en);
    if (rc) return rc;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    hrex_pack_kernel<<<(rows + 63) / 64, 64, 0, stream>>>(en, h->R, send, rows);
    h->launches++;
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

int atm_hrex_device_exchange(atm_handle *h, const double *gathered, uint64_t cycle, void *stream_) {
    ATM_NVTX_RANGE("atm_hrex_device_exchange");
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && h->hrex, ATM_ERR_STATE, "atm_hrex_device_exchange: call atm_hrex_device_setup first");
    ATM_REQUIRE(gathered, ATM_ERR_INVALID, "atm_hrex_device_exchange: null argument");
    HrexState *x = (HrexState *)h->hrex;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    int rc;
    if ((rc = upload_params_if_dirty(h, stream))) return rc;  // pending host edits first: the kernel overwrites the local rows
    hrex_exchange_kernel<<<1, HREX_THREADS, sizeof(int) * x->num_states, stream>>>(
        x->num_states, x->num_replicas, x->d_schedule, x->d_replica_state, x->d_gather_slot, gathered, x->beta, x->seed, cycle,
        x->d_local_global, h->R, h->d_params, x->d_counters);
    h->launches++;
    h->params_device_newer = true;
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

// ---- the communicator of the replica layer and the whole cycle in one call
}  // extern "C"

struct atm_re_comm {
    void *comm = nullptr;   // ncclComm_t
    int world = 1, rank = 0;
    bool owned = false;
};

extern "C" {

int atm_re_unique_id(void *id128) {
    ATM_REQUIRE(id128, ATM_ERR_INVALID, "atm_re_unique_id: null argument");
    NcclApi &n = nccl();
    ATM_REQUIRE(n.ok, ATM_ERR_UNSUPPORTED, "atm_re_unique_id: no usable libnccl.so.2 in this process or on the library path");
    const int rc = n.GetUniqueId(id128);
    ATM_REQUIRE(rc == 0, ATM_ERR_CUDA, "ncclGetUniqueId: %s", n.GetErrorString(rc));
    return ATM_OK;
}

int atm_re_comm_create(const void *id128, int32_t world, int32_t rank, int32_t device, atm_re_comm **out) {
    ATM_REQUIRE(out, ATM_ERR_INVALID, "atm_re_comm_create: null argument");
    *out = nullptr;
    ATM_REQUIRE(world >= 1 && rank >= 0 && rank < world, ATM_ERR_INVALID, "atm_re_comm_create: bad rank %d of %d", rank, world);
    atm_re_comm *c = new atm_re_comm();
    c->world = world;
    c->rank = rank;
    if (world > 1) {
        ATM_REQUIRE(id128, ATM_ERR_INVALID, "atm_re_comm_create: a unique id is needed for more than one rank");
        NcclApi &n = nccl();
        if (!n.ok) { delete c; set_error("atm_re_comm_create: no usable libnccl.so.2 in this process or on the library path"); return ATM_ERR_UNSUPPORTED; }
        if (device >= 0) ATM_CUDA_CHECK(cudaSetDevice(device));
        NcclId128 id;
        memcpy(id.b, id128, sizeof(id.b));
        const int rc = n.CommInitRank(&c->comm, world, id, rank);
        if (rc != 0) { delete c; set_error("ncclCommInitRank: %s", n.GetErrorString(rc)); return ATM_ERR_CUDA; }
        c->owned = true;
    }
    *out = c;
    return ATM_OK;
}

int atm_re_comm_from_nccl(void *nccl_comm, int32_t rank, atm_re_comm **out) {
    ATM_REQUIRE(out && nccl_comm, ATM_ERR_INVALID, "atm_re_comm_from_nccl: null argument");
    NcclApi &n = nccl();
    ATM_REQUIRE(n.ok, ATM_ERR_UNSUPPORTED, "atm_re_comm_from_nccl: no usable libnccl.so.2");
    int count = 0;
    const int rc = n.CommCount(nccl_comm, &count);
    ATM_REQUIRE(rc == 0, ATM_ERR_CUDA, "ncclCommCount: %s", n.GetErrorString(rc));
    atm_re_comm *c = new atm_re_comm();
    c->comm = nccl_comm;
    c->world = count;
    c->rank = rank;
    *out = c;
    return ATM_OK;
}

int atm_re_comm_destroy(atm_re_comm *c) {
    if (!c) return ATM_OK;
    if (c->owned && c->comm) nccl().CommDestroy(c->comm);
    delete c;
    return ATM_OK;
}

// pack -> all-gather over NVLink -> sweep, all on `stream`, no host synchronisation, capturable into a CUDA graph.
// comm == NULL or a one-rank communicator: the local rows are the gathered rows.
int atm_hrex_device_cycle(atm_handle *h, atm_re_comm *comm, uint64_t cycle, void *stream_) {
    ATM_NVTX_RANGE("atm_hrex_device_cycle");
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && h->hrex, ATM_ERR_STATE, "atm_hrex_device_cycle: call atm_hrex_device_setup first");
    HrexState *x = (HrexState *)h->hrex;
    const int world = comm ? comm->world : 1;
    ATM_REQUIRE(x->gathered_rows % world == 0, ATM_ERR_INVALID, "atm_hrex_device_cycle: %d gathered rows do not divide over %d ranks",
                x->gathered_rows, world);
    const int rows = x->gathered_rows / world;
    int rc;
    if ((rc = atm_hrex_device_pack(h, x->d_send, rows, stream))) return rc;
    const double *gathered = x->d_send;
    if (world > 1) {
        NcclApi &n = nccl();
        ATM_REQUIRE(n.ok && comm->comm, ATM_ERR_STATE, "atm_hrex_device_cycle: the communicator has no NCCL handle");
        const int nrc = n.AllGather(x->d_send, x->d_gathered, (size_t)2 * rows, NCCL_FLOAT64, comm->comm, stream);
        ATM_REQUIRE(nrc == 0, ATM_ERR_CUDA, "ncclAllGather: %s", n.GetErrorString(nrc));
        gathered = x->d_gathered;
    }
    return atm_hrex_device_exchange(h, gathered, cycle, stream);
}

int atm_hrex_device_state(atm_handle *h, int32_t *replica_state, int64_t counters[3], void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(h && h->hrex, ATM_ERR_STATE, "atm_hrex_device_state: call atm_hrex_device_setup first");
    HrexState *x = (HrexState *)h->hrex;
    ATM_CUDA_CHECK(cudaSetDevice(h->device));
    long long c[4] = {0, 0, 0, 0};
    if (replica_state)
        ATM_CUDA_CHECK(cudaMemcpyAsync(replica_state, x->d_replica_state, sizeof(int) * x->num_replicas, cudaMemcpyDeviceToHost, stream));
    ATM_CUDA_CHECK(cudaMemcpyAsync(c, x->d_counters, sizeof(c), cudaMemcpyDeviceToHost, stream));
    ATM_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (counters) { counters[0] = c[0]; counters[1] = c[1]; counters[2] = c[2]; }
    return ATM_OK;
}

}  // extern "C"
