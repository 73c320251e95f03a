// FFMA vs FFMA2 (fma.rn.f32x2, sm_100+) issue rate per SM: 8 independent chains per thread, 16 warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int kPacked>
__global__ void __launch_bounds__(512, 1) rate_kernel(float* out, int iters, long long* cycles) {
    float2 acc[8];
    const float2 a = make_float2(1.0001f + threadIdx.x * 1e-7f, 0.9999f), b = make_float2(1e-6f, -1e-6f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2((float)i, (float)-i);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kPacked) {
                acc[i] = __ffma2_rn(acc[i], a, b);
            } else {
                acc[i].x = fmaf(acc[i].x, a.x, b.x);
                acc[i].y = fmaf(acc[i].y, a.y, b.y);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    cudaMallocManaged(&cyc, sizeof(long long));
    const int iters = 4096;
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 2; ++rep) {
            if (packed) rate_kernel<1><<<148, 512>>>(out, iters, cyc);
            else rate_kernel<0><<<148, 512>>>(out, iters, cyc);
            cudaDeviceSynchronize();
        }
        const double fma_per_thread = 16.0 * iters;                 // scalar FMAs per thread
        const double per_clk_sm = fma_per_thread * 512 / (double)*cyc;
        printf("%s: %lld cycles, %.1f fp32 FMA / clk / SM, %.2f warp-instructions / clk / SM\n", packed ? "FFMA2" : "FFMA ",
               *cyc, per_clk_sm, per_clk_sm / 32 / (packed ? 2 : 1));
    }
    return 0;
}
