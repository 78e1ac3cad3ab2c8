// atm_copy_merge.cu -- Tier-1 bandwidth kernels: CopyState and HybridForce for sm_100a.
//
// Both are pure streaming kernels (no reuse): the bound is HBM bandwidth once the working set exceeds the
// 126 MB L2, launch latency below that.  Design: 128-bit accesses, every load of a thread's UNROLL atoms
// issued before the first store (memory-level parallelism), one sweep over the atoms (the reference
// kernel sweeps twice and reads posq twice, platforms/common/src/kernels/atmmetaforce.cc:33-51).
//
// Algorithmic bytes per atom (DESIGN.md "Kernels"): copy 64 (single) / 112 (mixed, double); merge 96.
#include "atm_common.cuh"

namespace atm {

constexpr int COPY_THREADS = 256;
constexpr int COPY_UNROLL = 4;

// posq1 = posq ; posq2 = posq + (real)displ with .w + 0 ; corrections verbatim.
// ref semantics: kernels/atmmetaforce.cc:33-51 with real = float.
template <bool MIXED>
__global__ void __launch_bounds__(COPY_THREADS)
copy_state_f32_kernel(int n, const float4 *__restrict__ posq, const float4 *__restrict__ corr,
                      const float4 *__restrict__ displ, float4 *__restrict__ posq1, float4 *__restrict__ corr1,
                      float4 *__restrict__ posq2, float4 *__restrict__ corr2) {
    const int base = blockIdx.x * (COPY_THREADS * COPY_UNROLL) + threadIdx.x;
    float4 p[COPY_UNROLL], d[COPY_UNROLL], c[COPY_UNROLL];
#pragma unroll
    for (int u = 0; u < COPY_UNROLL; u++) {
        const int i = base + u * COPY_THREADS;
        if (i < n) {
            p[u] = __ldg(posq + i);
            d[u] = __ldg(displ + i);
            if (MIXED) c[u] = __ldg(corr + i);
        }
    }
#pragma unroll
    for (int u = 0; u < COPY_UNROLL; u++) {
        const int i = base + u * COPY_THREADS;
        if (i < n) {
            float4 q;
            q.x = __fadd_rn(p[u].x, d[u].x);
            q.y = __fadd_rn(p[u].y, d[u].y);
            q.z = __fadd_rn(p[u].z, d[u].z);
            q.w = __fadd_rn(p[u].w, 0.0f);
            posq1[i] = p[u];
            posq2[i] = q;
            if (MIXED) {
                corr1[i] = c[u];
                corr2[i] = c[u];
            }
        }
    }
}

// real = double: posq is double4 (32 B) handled as two double2 halves; the table stays float4.
__global__ void __launch_bounds__(COPY_THREADS)
copy_state_f64_kernel(int n, const double2 *__restrict__ posq, const float4 *__restrict__ displ,
                      double2 *__restrict__ posq1, double2 *__restrict__ posq2) {
    const int base = blockIdx.x * (COPY_THREADS * COPY_UNROLL) + threadIdx.x;
    double2 lo[COPY_UNROLL], hi[COPY_UNROLL];
    float4 d[COPY_UNROLL];
#pragma unroll
    for (int u = 0; u < COPY_UNROLL; u++) {
        const int i = base + u * COPY_THREADS;
        if (i < n) {
            lo[u] = __ldg(posq + 2 * i);
            hi[u] = __ldg(posq + 2 * i + 1);
            d[u] = __ldg(displ + i);
        }
    }
#pragma unroll
    for (int u = 0; u < COPY_UNROLL; u++) {
        const int i = base + u * COPY_THREADS;
        if (i < n) {
            posq1[2 * i] = lo[u];
            posq1[2 * i + 1] = hi[u];
            double2 a, b;
            a.x = __dadd_rn(lo[u].x, (double)d[u].x);
            a.y = __dadd_rn(lo[u].y, (double)d[u].y);
            b.x = __dadd_rn(hi[u].x, (double)d[u].z);
            b.y = __dadd_rn(hi[u].y, 0.0);
            posq2[2 * i] = a;
            posq2[2 * i + 1] = b;
        }
    }
}

int launch_copy_state(atm_handle *h, const void *posq, const void *corr, void *posq1, void *corr1, void *posq2,
                      void *corr2, cudaStream_t stream) {
    const int n = h->N;
    if (n == 0) return ATM_OK;
    const int per_block = COPY_THREADS * COPY_UNROLL;
    const int grid = (n + per_block - 1) / per_block;
    if (h->cfg.precision == ATM_PREC_DOUBLE) {
        copy_state_f64_kernel<<<grid, COPY_THREADS, 0, stream>>>(n, (const double2 *)posq, h->d_displ,
                                                                  (double2 *)posq1, (double2 *)posq2);
    } else if (corr != nullptr) {
        copy_state_f32_kernel<true><<<grid, COPY_THREADS, 0, stream>>>(
            n, (const float4 *)posq, (const float4 *)corr, h->d_displ, (float4 *)posq1, (float4 *)corr1,
            (float4 *)posq2, (float4 *)corr2);
    } else {
        copy_state_f32_kernel<false><<<grid, COPY_THREADS, 0, stream>>>(
            n, (const float4 *)posq, nullptr, h->d_displ, (float4 *)posq1, nullptr, (float4 *)posq2, nullptr);
    }
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

// posq_out.xyz = posq_in.xyz wrapped into [0, L) of a rectangular box; .w untouched.
__global__ void wrap_kernel(int n, const float4 *__restrict__ in, float4 *__restrict__ out, float3 box, float3 inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(in + i);
    p.x -= box.x * floorf(p.x * inv.x);
    p.y -= box.y * floorf(p.y * inv.y);
    p.z -= box.z * floorf(p.z * inv.z);
    out[i] = p;
}

int launch_wrap(atm_handle *h, const void *in, void *out, const double box[9], cudaStream_t stream) {
    ATM_REQUIRE(box[1] == 0 && box[2] == 0 && box[3] == 0 && box[5] == 0 && box[6] == 0 && box[7] == 0,
                ATM_ERR_UNSUPPORTED, "atm_wrap_positions: only rectangular boxes are supported");
    ATM_REQUIRE(h->cfg.precision != ATM_PREC_DOUBLE, ATM_ERR_UNSUPPORTED, "atm_wrap_positions: float4 only");
    if (h->N == 0) return ATM_OK;
    float3 b = make_float3((float)box[0], (float)box[4], (float)box[8]);
    float3 inv = make_float3(1.0f / b.x, 1.0f / b.y, 1.0f / b.z);
    wrap_kernel<<<(h->N + 255) / 256, 256, 0, stream>>>(h->N, (const float4 *)in, (float4 *)out, b, inv);
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

// ---------------------------------------------------------------------------------------------
// HybridForce: force += llrint(sp*f2 + (1-sp)*f1) on the int64 fixed-point SoA buffers.
// ref semantics: kernels/atmmetaforce.cc:8-16; blend in double (Reference platform arithmetic,
// ReferenceATMMetaForceKernels.cpp:101-104) with explicit _rn intrinsics so that no FMA contraction
// happens and the result is bit-identical to oracle/atm_oracle.c:atm_oracle_hybrid_force_i64.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long blend(long long f1, long long f2, double sp, double sp1) {
    double v = __dadd_rn(__dmul_rn(sp, (double)f2), __dmul_rn(sp1, (double)f1));
    return __double2ll_rn(v);
}

constexpr int MERGE_THREADS = 256;

// Two atoms per thread per component: 128-bit loads/stores (P even keeps every component block 16 B aligned).
__global__ void __launch_bounds__(MERGE_THREADS)
hybrid_force_vec2_kernel(int n, int P, longlong2 *__restrict__ force, const longlong2 *__restrict__ f1,
                         const longlong2 *__restrict__ f2, double sp) {
    const int k = blockIdx.x * MERGE_THREADS + threadIdx.x;  // pair index
    const int npairs = n >> 1;
    const double sp1 = 1.0 - sp;
    if (k < npairs) {
        const int halfP = P >> 1;
        longlong2 a[3], b[3], f[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            a[c] = __ldg(f1 + c * halfP + k);
            b[c] = __ldg(f2 + c * halfP + k);
            f[c] = force[c * halfP + k];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            f[c].x += blend(a[c].x, b[c].x, sp, sp1);
            f[c].y += blend(a[c].y, b[c].y, sp, sp1);
            force[c * halfP + k] = f[c];
        }
    } else if (k == npairs && (n & 1)) {
        const long long *s1 = (const long long *)f1, *s2 = (const long long *)f2;
        long long *fo = (long long *)force;
        const int i = n - 1;
#pragma unroll
        for (int c = 0; c < 3; c++) fo[c * P + i] += blend(s1[c * P + i], s2[c * P + i], sp, sp1);
    }
}

__global__ void __launch_bounds__(MERGE_THREADS)
hybrid_force_scalar_kernel(int n, int P, long long *__restrict__ force, const long long *__restrict__ f1,
                           const long long *__restrict__ f2, double sp) {
    const int i = blockIdx.x * MERGE_THREADS + threadIdx.x;
    const double sp1 = 1.0 - sp;
    if (i < n) {
#pragma unroll
        for (int c = 0; c < 3; c++) force[c * P + i] += blend(f1[c * P + i], f2[c * P + i], sp, sp1);
    }
}

int launch_hybrid_force(atm_handle *h, int64_t *force, const int64_t *f1, const int64_t *f2, double sp,
                        cudaStream_t stream) {
    const int n = h->N, P = h->P;
    if (n == 0) return ATM_OK;
    const bool aligned = ((P & 1) == 0) && (((uintptr_t)force | (uintptr_t)f1 | (uintptr_t)f2) & 15) == 0;
    if (aligned) {
        const int work = (n >> 1) + 1;
        hybrid_force_vec2_kernel<<<(work + MERGE_THREADS - 1) / MERGE_THREADS, MERGE_THREADS, 0, stream>>>(
            n, P, (longlong2 *)force, (const longlong2 *)f1, (const longlong2 *)f2, sp);
    } else {
        hybrid_force_scalar_kernel<<<(n + MERGE_THREADS - 1) / MERGE_THREADS, MERGE_THREADS, 0, stream>>>(
            n, P, (long long *)force, (const long long *)f1, (const long long *)f2, sp);
    }
    ATM_CUDA_CHECK(cudaGetLastError());
    return ATM_OK;
}

}  // namespace atm
