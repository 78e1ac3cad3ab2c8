// f32x2_throughput.cu -- issue rate of Blackwell's packed FP32 instructions (FFMA2 / FMUL2 / FADD2) against their
// scalar forms, per SM sub-partition.  Evidence for DESIGN.md section 3.2: does an FFMA2 occupy the FP32 pipe for one
// issue cycle or two?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_throughput tools/microbench/f32x2_throughput.cu && ./f32x2_throughput
//
// Each warp runs ITER iterations of 8 independent dependency chains (so latency never limits), 16 warps per SM (4 per
// sub-partition); time comes from clock64 on one SM.  Output: warp instructions per cycle per sub-partition, and the
// FP32 lane-operations per cycle per SM that implies.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int CHAINS = 8;

template <int MODE>
__global__ void __launch_bounds__(512) kernel(float *out, long long *cycles, float seed) {
    float2 a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = make_float2(seed + k + threadIdx.x, seed - k);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) {
            if (MODE == 0) a[k] = __ffma2_rn(a[k], m, c);                                    // FFMA2
            if (MODE == 1) { a[k].x = fmaf(a[k].x, m.x, c.x); a[k].y = fmaf(a[k].y, m.y, c.y); }  // 2 x FFMA
            if (MODE == 2) a[k] = __fmul2_rn(a[k], m);                                       // FMUL2
            if (MODE == 3) a[k] = __fadd2_rn(a[k], c);                                       // FADD2
            if (MODE == 4) { a[k].x = a[k].x * m.x; a[k].y = a[k].y * m.y; }                 // 2 x FMUL
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += a[k].x + a[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int instr_per_chain_step, int lane_ops_per_instr) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out;
    long long *cyc, h[1024];
    cudaMalloc(&out, sizeof(float) * sms * 512);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    kernel<MODE><<<sms, 512>>>(out, cyc, 1.0f);   // warm-up
    kernel<MODE><<<sms, 512>>>(out, cyc, 2.0f);
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long worst = 0;
    for (int i = 0; i < sms; i++) worst = h[i] > worst ? h[i] : worst;
    // 512 threads = 16 warps per block, one block per SM, 4 sub-partitions: 4 warps per sub-partition
    const double warp_instr_per_subpartition = 4.0 * ITER * CHAINS * instr_per_chain_step;
    const double ipc = warp_instr_per_subpartition / (double)worst;
    printf("%-10s %8lld cycles   %.3f warp instr / cycle / sub-partition   %.1f fp32 lane-ops / cycle / SM\n", name, worst, ipc,
           ipc * 4 * 32 * lane_ops_per_instr);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("FFMA2", 1, 2);
    run<1>("2xFFMA", 2, 1);
    run<2>("FMUL2", 1, 2);
    run<3>("FADD2", 1, 2);
    run<4>("2xFMUL", 2, 1);
    return 0;
}
