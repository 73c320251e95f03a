// Debug helper: throughput of legacy mma.sync.m16n8k16 FP16 (fp32 accumulate) on sm_100a, 8 independent
// accumulators per warp (as a 2x4 register tile would issue them) and with only 4.
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(float* out, int iters) {
    float c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(float* out) {
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<NACC><<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<NACC><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double mmas = (double)iters * NACC * warps;             // per SM
        const double cyc = ms * 1e-3 * 1.965e9;
        printf("acc %d warps/SM %2d: %.3f ms, %.3f mma.sync(m16n8k16 f16) per cycle per SM = %.0f MAC/clk/SM, %.1f TFLOP/s chip\n",
               NACC, warps, ms, mmas / cyc, mmas / cyc * 16 * 8 * 16, mmas * 148 * 16 * 8 * 16 * 2 / (ms * 1e-3) / 1e12);
    }
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
    run<8>(out);
    run<4>(out);
    return 0;
}
