// atm_host.cu -- the hot path with HOST buffers on both sides (atm_host_pipeline_*, include/atm_b200.h).
//
// The reference's Context API moves coordinates in and forces/energies out through host memory
// (context.setPositions / getState; ref: python/tests/test_abfe.py:115-146).  This file is the native form of that
// call for the Tier-2 path: one or more handles ("chunks" of the replicas resident on this GPU) each get
//     H2D coordinates -> zero the force staging buffer -> [rebuild | prune] -> pack, nb2, merge -> D2H forces, energies
// on their own stream, forked from and joined back into the caller's stream, so the copies of one chunk overlap the
// kernels of the others (PCIe is full duplex: the bound is max(compute, D2H), not their sum).  The whole fork/join is
// captured ONCE per maintenance variant into a CUDA graph; a step is one cudaGraphLaunch (about 10 us of host time,
// whatever the number of chunks -- the Python loop this replaces cost ~65 us per chunk and was the e2e bound at N >= 2).
#include <string.h>

#include <algorithm>
#include <vector>

#include "atm_common.cuh"

using namespace atm;

namespace {

struct Chunk {
    atm_handle *h = nullptr;
    float4 *posq = nullptr;        // device staging [R][P]
    long long *force = nullptr;    // device staging [R][3P]
    float *force_f32 = nullptr;    // device staging [R][3P] for force_format 1 (allocated on first use)
    float *posq3 = nullptr;        // device staging [R][P][3] for posq_format 1 (allocated on first use)
    long long *ext[2] = {nullptr, nullptr};   // device staging [R][3P] of force_state{1,2}_ext_host (allocated on first use)
    double *eext = nullptr;        // device staging [R][2] of energy_ext_host
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint64_t generation = 0;       // alloc generation the cached graphs were captured against
};

}  // namespace

struct atm_host_pipeline {
    int device = 0;
    std::vector<Chunk> chunks;
    cudaEvent_t fork = nullptr;
    // maintenance 0 / 1 / 2 / 3 (3 = concurrent prune), times the copy of the pruned list the chunks read (0 / 1)
    cudaGraphExec_t exec[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<uint64_t> exec_launches_chunk[8];            // own kernels per replay, per chunk
    std::vector<atm_host_io> ios;                            // the buffers the cached graphs were captured with
};

static void drop_graphs(atm_host_pipeline *p) {
    for (int v = 0; v < 8; v++)
        if (p->exec[v]) { cudaGraphExecDestroy(p->exec[v]); p->exec[v] = nullptr; }
}

namespace atm {
// force_format 1: the merged fixed-point force as float32 kJ/mol/nm, same [R][3][P] layout: half the D2H bytes
__global__ void force_to_f32_kernel(const long long *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)((double)in[i] * (1.0 / 4294967296.0));
}
// posq_format 1: packed float3 coordinates -> the float4 staging buffer the kernels read (w unused by the Tier-2 path)
__global__ void posq3_to_posq4_kernel(const float *__restrict__ in, float4 *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 0.f);
}
}  // namespace atm

// enqueue one step of every chunk (fork from `stream`, join back into it); capturable
static int enqueue_all(atm_host_pipeline *p, const atm_host_io *ios, int maintenance, cudaStream_t stream) {
    ATM_CUDA_CHECK(cudaEventRecord(p->fork, stream));
    for (size_t c = 0; c < p->chunks.size(); c++) {
        Chunk &k = p->chunks[c];
        atm_handle *h = k.h;
        const size_t np = (size_t)h->R * h->P;
        int rc;
        ATM_CUDA_CHECK(cudaStreamWaitEvent(k.stream, p->fork, 0));
        if (ios[c].posq_format == ATM_POSQ_F3) {
            ATM_CUDA_CHECK(cudaMemcpyAsync(k.posq3, ios[c].posq_host, sizeof(float) * 3 * np, cudaMemcpyHostToDevice, k.stream));
            posq3_to_posq4_kernel<<<(unsigned)((np + 255) / 256), 256, 0, k.stream>>>(k.posq3, k.posq, np);
            h->launches++;
        } else {
            ATM_CUDA_CHECK(cudaMemcpyAsync(k.posq, ios[c].posq_host, sizeof(float4) * np, cudaMemcpyHostToDevice, k.stream));
        }
        ATM_CUDA_CHECK(cudaMemsetAsync(k.force, 0, sizeof(long long) * 3 * np, k.stream));
        const int64_t *ext_host[2] = {ios[c].force_state1_ext_host, ios[c].force_state2_ext_host};
        for (int s = 0; s < 2; s++)
            if (ext_host[s])
                ATM_CUDA_CHECK(cudaMemcpyAsync(k.ext[s], ext_host[s], sizeof(long long) * 3 * np, cudaMemcpyHostToDevice, k.stream));
        if (ios[c].energy_ext_host)
            ATM_CUDA_CHECK(cudaMemcpyAsync(k.eext, ios[c].energy_ext_host, sizeof(double) * 2 * (size_t)h->R, cudaMemcpyHostToDevice, k.stream));
        if ((rc = nb_host_enqueue(h, k.posq, k.force, ios[c].include_energy, maintenance, k.stream, ext_host[0] ? k.ext[0] : nullptr,
                                  ext_host[1] ? k.ext[1] : nullptr, ios[c].energy_ext_host ? k.eext : nullptr)))
            return rc;
        if (ios[c].force_format == ATM_FORCE_F32) {
            force_to_f32_kernel<<<(unsigned)((3 * np + 255) / 256), 256, 0, k.stream>>>(k.force, k.force_f32, 3 * np);
            h->launches++;
            ATM_CUDA_CHECK(cudaMemcpyAsync(ios[c].force_host, k.force_f32, sizeof(float) * 3 * np, cudaMemcpyDeviceToHost, k.stream));
        } else if (ios[c].force_format == ATM_FORCE_I64) {
            ATM_CUDA_CHECK(cudaMemcpyAsync(ios[c].force_host, k.force, sizeof(long long) * 3 * np, cudaMemcpyDeviceToHost, k.stream));
        }   // ATM_FORCE_NONE: energies only
        if (ios[c].energies_host)
            ATM_CUDA_CHECK(cudaMemcpyAsync(ios[c].energies_host, nb_energies_device(h), sizeof(double) * (size_t)h->R * ATM_NUM_ENERGY_SLOTS,
                                           cudaMemcpyDeviceToHost, k.stream));
        ATM_CUDA_CHECK(cudaEventRecord(k.done, k.stream));
        ATM_CUDA_CHECK(cudaStreamWaitEvent(stream, k.done, 0));
    }
    return ATM_OK;
}

extern "C" {

int atm_stream_create(int32_t device, void **stream) {
    ATM_REQUIRE(stream, ATM_ERR_INVALID, "atm_stream_create: null argument");
    *stream = nullptr;
    int count = 0;
    ATM_REQUIRE(cudaGetDeviceCount(&count) == cudaSuccess && count > 0, ATM_ERR_CUDA,
                "atm_stream_create: no CUDA device available (this library has no CPU fallback)");
    if (device >= 0) ATM_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t s = nullptr;
    ATM_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return ATM_OK;
}

int atm_stream_destroy(void *stream) {
    if (!stream) return ATM_OK;
    ATM_CUDA_CHECK(cudaStreamDestroy((cudaStream_t)stream));
    return ATM_OK;
}

int atm_stream_synchronize(void *stream) {
    ATM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return ATM_OK;
}

int atm_host_alloc(size_t bytes, void **ptr) {
    ATM_REQUIRE(ptr && bytes > 0, ATM_ERR_INVALID, "atm_host_alloc: null argument / zero size");
    *ptr = nullptr;
    int count = 0;
    ATM_REQUIRE(cudaGetDeviceCount(&count) == cudaSuccess && count > 0, ATM_ERR_CUDA,
                "atm_host_alloc: no CUDA device available (this library has no CPU fallback)");
    cudaError_t err = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (err != cudaSuccess) {
        *ptr = nullptr;
        set_error("atm_host_alloc: %s", cudaGetErrorString(err));
        return err == cudaErrorMemoryAllocation ? ATM_ERR_NOMEM : ATM_ERR_CUDA;
    }
    return ATM_OK;
}

int atm_host_free(void *ptr) {
    if (!ptr) return ATM_OK;
    ATM_CUDA_CHECK(cudaFreeHost(ptr));
    return ATM_OK;
}

int atm_host_pipeline_create(int32_t num_handles, atm_handle *const *handles, atm_host_pipeline **out) {
    ATM_REQUIRE(out && handles && num_handles > 0, ATM_ERR_INVALID, "atm_host_pipeline_create: null argument / no handles");
    *out = nullptr;
    for (int c = 0; c < num_handles; c++) {
        ATM_REQUIRE(handles[c] && handles[c]->nb, ATM_ERR_STATE, "atm_host_pipeline_create: handle %d has no Tier-2 set-up (atm_nb_setup)", c);
        ATM_REQUIRE(handles[c]->device == handles[0]->device, ATM_ERR_INVALID, "atm_host_pipeline_create: handles live on different devices");
        ATM_REQUIRE(handles[c]->cfg.precision != ATM_PREC_DOUBLE, ATM_ERR_UNSUPPORTED, "atm_host_pipeline_create: float4 coordinates only");
        for (int b = 0; b < c; b++)
            ATM_REQUIRE(handles[b] != handles[c], ATM_ERR_INVALID, "atm_host_pipeline_create: handle %d listed twice", c);
    }
    ATM_CUDA_CHECK(cudaSetDevice(handles[0]->device));
    atm_host_pipeline *p = new atm_host_pipeline();
    p->device = handles[0]->device;
    p->chunks.resize(num_handles);
    cudaError_t err = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming);
    // earlier chunks get the higher stream priority: the block scheduler then drains the chunks roughly first-in
    // first-out (chunk c's D2H overlaps the kernels of chunk c+1) instead of sharing the SMs evenly among all chunks,
    // which would finish them together and leave every D2H copy for the end
    int prio_least = 0, prio_greatest = 0;
    if (err == cudaSuccess) err = cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    for (int c = 0; c < num_handles && err == cudaSuccess; c++) {
        Chunk &k = p->chunks[c];
        k.h = handles[c];
        const size_t np = (size_t)k.h->R * k.h->P;
        if ((err = cudaMalloc(&k.posq, sizeof(float4) * np)) != cudaSuccess) break;
        if ((err = cudaMalloc(&k.force, sizeof(long long) * 3 * np)) != cudaSuccess) break;
        if ((err = cudaStreamCreateWithPriority(&k.stream, cudaStreamNonBlocking, std::min(prio_least, prio_greatest + c))) != cudaSuccess) break;
        err = cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming);
    }
    if (err != cudaSuccess) {
        set_error("atm_host_pipeline_create: %s", cudaGetErrorString(err));
        atm_host_pipeline_destroy(p);
        return err == cudaErrorMemoryAllocation ? ATM_ERR_NOMEM : ATM_ERR_CUDA;
    }
    *out = p;
    return ATM_OK;
}

int atm_host_pipeline_destroy(atm_host_pipeline *p) {
    if (!p) return ATM_OK;
    cudaSetDevice(p->device);
    drop_graphs(p);
    for (Chunk &k : p->chunks) {
        if (k.stream) { cudaStreamSynchronize(k.stream); cudaStreamDestroy(k.stream); }
        if (k.done) cudaEventDestroy(k.done);
        if (k.posq) cudaFree(k.posq);
        if (k.force) cudaFree(k.force);
        if (k.force_f32) cudaFree(k.force_f32);
        if (k.posq3) cudaFree(k.posq3);
        if (k.ext[0]) cudaFree(k.ext[0]);
        if (k.ext[1]) cudaFree(k.ext[1]);
        if (k.eext) cudaFree(k.eext);
    }
    if (p->fork) cudaEventDestroy(p->fork);
    delete p;
    return ATM_OK;
}

int atm_host_pipeline_step(atm_host_pipeline *p, const atm_host_io *ios, int32_t maintenance, void *stream_) {
    ATM_NVTX_RANGE("atm_host_pipeline_step");
    cudaStream_t stream = (cudaStream_t)stream_;
    ATM_REQUIRE(p && ios, ATM_ERR_INVALID, "atm_host_pipeline_step: null argument");
    ATM_REQUIRE(maintenance >= 0 && maintenance <= 3, ATM_ERR_INVALID,
                "atm_host_pipeline_step: maintenance must be 0 (none), 1 (prune), 2 (rebuild) or 3 (prune concurrently with the step)");
    ATM_REQUIRE(stream != nullptr, ATM_ERR_INVALID, "atm_host_pipeline_step: needs a non-default stream (it is captured)");
    const size_t nc = p->chunks.size();
    for (size_t c = 0; c < nc; c++) {
        ATM_REQUIRE(ios[c].posq_host, ATM_ERR_INVALID, "atm_host_pipeline_step: chunk %d: posq_host is required", (int)c);
        ATM_REQUIRE(ios[c].force_format >= ATM_FORCE_I64 && ios[c].force_format <= ATM_FORCE_NONE, ATM_ERR_INVALID,
                    "atm_host_pipeline_step: chunk %d: unknown force_format %d", (int)c, (int)ios[c].force_format);
        ATM_REQUIRE(ios[c].force_host || ios[c].force_format == ATM_FORCE_NONE, ATM_ERR_INVALID,
                    "atm_host_pipeline_step: chunk %d: force_host is required unless force_format is ATM_FORCE_NONE", (int)c);
        ATM_REQUIRE(ios[c].force_format != ATM_FORCE_NONE || ios[c].energies_host, ATM_ERR_INVALID,
                    "atm_host_pipeline_step: chunk %d: nothing to return (no forces, no energies)", (int)c);
        ATM_REQUIRE(ios[c].posq_format == ATM_POSQ_F4 || ios[c].posq_format == ATM_POSQ_F3, ATM_ERR_INVALID,
                    "atm_host_pipeline_step: chunk %d: unknown posq_format %d", (int)c, (int)ios[c].posq_format);
        if (ios[c].posq_format == ATM_POSQ_F3 && !p->chunks[c].posq3) {
            ATM_CUDA_CHECK(cudaSetDevice(p->device));
            ATM_CUDA_CHECK(cudaMalloc(&p->chunks[c].posq3, sizeof(float) * 3 * (size_t)p->chunks[c].h->R * p->chunks[c].h->P));
        }
        const int64_t *ext_host[2] = {ios[c].force_state1_ext_host, ios[c].force_state2_ext_host};
        for (int s = 0; s < 2; s++)
            if (ext_host[s] && !p->chunks[c].ext[s]) {
                ATM_CUDA_CHECK(cudaSetDevice(p->device));
                ATM_CUDA_CHECK(cudaMalloc(&p->chunks[c].ext[s], sizeof(long long) * 3 * (size_t)p->chunks[c].h->R * p->chunks[c].h->P));
            }
        if (ios[c].energy_ext_host && !p->chunks[c].eext) {
            ATM_CUDA_CHECK(cudaSetDevice(p->device));
            ATM_CUDA_CHECK(cudaMalloc(&p->chunks[c].eext, sizeof(double) * 2 * (size_t)p->chunks[c].h->R));
        }
        if (ios[c].force_format == ATM_FORCE_F32 && !p->chunks[c].force_f32) {
            ATM_CUDA_CHECK(cudaSetDevice(p->device));
            ATM_CUDA_CHECK(cudaMalloc(&p->chunks[c].force_f32, sizeof(float) * 3 * (size_t)p->chunks[c].h->R * p->chunks[c].h->P));
        }
    }
    ATM_CUDA_CHECK(cudaSetDevice(p->device));
    int rc;
    bool any_sync = false;
    for (size_t c = 0; c < nc; c++) {
        bool sync_rebuild = false;
        if ((rc = nb_host_prepare(p->chunks[c].h, maintenance, stream, &sync_rebuild))) return rc;
        any_sync |= sync_rebuild;
    }
    if (any_sync) {
        // first build (or a capacity change): the synchronous, verified path of atm_nb_rebuild on the staged coordinates
        for (size_t c = 0; c < nc; c++) {
            Chunk &k = p->chunks[c];
            const size_t np = (size_t)k.h->R * k.h->P;
            if (ios[c].posq_format == ATM_POSQ_F3) {
                ATM_CUDA_CHECK(cudaMemcpyAsync(k.posq3, ios[c].posq_host, sizeof(float) * 3 * np, cudaMemcpyHostToDevice, stream));
                posq3_to_posq4_kernel<<<(unsigned)((np + 255) / 256), 256, 0, stream>>>(k.posq3, k.posq, np);
            } else {
                ATM_CUDA_CHECK(cudaMemcpyAsync(k.posq, ios[c].posq_host, sizeof(float4) * np, cudaMemcpyHostToDevice, stream));
            }
            if ((rc = atm_nb_rebuild(k.h, k.posq, stream))) return rc;
        }
        maintenance = 0;
    }
    bool stale = p->ios.size() != nc || memcmp(p->ios.data(), ios, sizeof(atm_host_io) * nc) != 0;
    for (size_t c = 0; c < nc; c++) stale |= p->chunks[c].generation != nb_alloc_generation(p->chunks[c].h);
    if (stale) {
        drop_graphs(p);
        p->ios.assign(ios, ios + nc);
        for (size_t c = 0; c < nc; c++) p->chunks[c].generation = nb_alloc_generation(p->chunks[c].h);
    }
    const int copy = nb_host_inner_copy(p->chunks[0].h);
    for (size_t c = 1; c < nc; c++)
        ATM_REQUIRE(nb_host_inner_copy(p->chunks[c].h) == copy || maintenance == 2, ATM_ERR_STATE,
                    "atm_host_pipeline_step: the handles of a pipeline must be stepped together (pruned-list copies differ)");
    const int m = maintenance;
    const int v = 2 * maintenance + (maintenance == 2 ? 0 : copy);   // a rebuild always writes copy 0
    if (!p->exec[v]) {
        std::vector<uint64_t> before(nc);
        for (size_t c = 0; c < nc; c++) before[c] = p->chunks[c].h->launches;
        cudaGraph_t graph = nullptr;
        ATM_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_all(p, ios, m, stream);
        cudaError_t err = cudaStreamEndCapture(stream, &graph);
        p->exec_launches_chunk[v].assign(nc, 0);
        for (size_t c = 0; c < nc; c++) {
            p->exec_launches_chunk[v][c] = p->chunks[c].h->launches - before[c];
            p->chunks[c].h->launches = before[c];
        }
        if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
        ATM_REQUIRE(err == cudaSuccess && graph, ATM_ERR_CUDA, "atm_host_pipeline_step: capture failed: %s", cudaGetErrorString(err));
        err = cudaGraphInstantiate(&p->exec[v], graph, 0);
        cudaGraphDestroy(graph);
        if (err != cudaSuccess) p->exec[v] = nullptr;
        ATM_REQUIRE(err == cudaSuccess, ATM_ERR_CUDA, "atm_host_pipeline_step: instantiate failed: %s", cudaGetErrorString(err));
    }
    ATM_CUDA_CHECK(cudaGraphLaunch(p->exec[v], stream));
    for (size_t c = 0; c < nc; c++) {
        p->chunks[c].h->launches += p->exec_launches_chunk[v][c];
        if (m == 2 && (rc = nb_host_rebuild_enqueued(p->chunks[c].h, stream))) return rc;
        if (m == 3) nb_host_flip_inner(p->chunks[c].h);
    }
    return ATM_OK;
}

int atm_host_pipeline_check(atm_host_pipeline *p) {
    ATM_REQUIRE(p, ATM_ERR_INVALID, "atm_host_pipeline_check: null argument");
    int first = ATM_OK;
    for (Chunk &k : p->chunks) {
        const int rc = atm_nb_check(k.h, 1);
        if (rc != ATM_OK && first == ATM_OK) first = rc;   // every handle is checked: each raises its own capacities
    }
    return first;
}

}  // extern "C"
