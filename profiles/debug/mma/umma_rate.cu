// umma_rate.cu -- how long does one tcgen05.mma (kind::f16, M = 128, operands in shared memory) take back to back, as a
// function of N and of the operand layout?  (The update kernel's MMA thread measured 105-130 cycles per M128 x N128 x K16
// instruction with the no-swizzle layouts against a math floor of 64.)  One CTA, thread 0 issues R x 4 MMAs, commits, waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../../uav-wrf-les-ppo-lstm_b200/csrc -I../../../include umma_rate.cu -o umma_rate
#include <cstdio>
#include "tc_gemm.cuh"
using namespace plume;

__device__ __forceinline__ uint64_t desc_layout(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return tc::make_smem_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}

// mode 0: A, B K-major interleaved (LBO 128, SBO 1024, K step 256 B)     -- G1
// mode 1: A MN-major interleaved (LBO 2048, SBO 128), B K-major          -- G2
// mode 2: A K-major (LBO 128, SBO 2048), B MN-major (LBO 2048, SBO 128)  -- G3
// mode 3: A, B K-major, 128-byte swizzle (SBO 1024, K step 32 B inside the 128-byte row)
// mode 4: the update kernel's G2 pattern: per K-step (A_lo, B_hi), (A_hi, B_lo), (A_hi, B_hi) with A MN-major from a resident
//         64 KB operand and B K-major from a 32 KB half; spinners > 0: that many extra warps poll the completion barrier
__global__ void __launch_bounds__(544, 1) umma_rate_kernel(int mode, int N, int reps, int spinners, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0 pairs
    if (tid == 0) {
        tc::mbar_init(&done, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);   // (the fill loop above covers the first 96 KB; the rest of the 160 KB is whatever it holds)
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a = tc::smem_u32(smem), b = a + 32768;
        uint32_t idesc = tc::make_idesc_f16(128, N);
        if (mode == 1 || mode == 4) idesc |= 1u << 15;
        if (mode == 2) idesc |= 1u << 16;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint64_t da, db;
                if (mode == 0) {
                    da = desc_layout(a + j * 256, 128, 1024, 0);
                    db = desc_layout(b + j * 256, 128, 1024, 0);
                } else if (mode == 1) {
                    da = desc_layout(a + j * 2 * 2048, 2048, 128, 0);
                    db = desc_layout(b + j * 256, 128, 1024, 0);
                } else if (mode == 2) {
                    da = desc_layout(a + j * 256, 128, 2048, 0);
                    db = desc_layout(b + j * 2 * 2048, 2048, 128, 0);
                } else if (mode == 4) {
                    const uint32_t ah = a + j * 2 * 2048, al = ah + 32768, bh = a + 65536 + j * 256, bl = bh + 16384;
                    tc::mma_f16(tmem, desc_layout(al, 2048, 128, 0), desc_layout(bh, 128, 1024, 0), idesc, (r | j) ? 1u : 0u);
                    tc::mma_f16(tmem, desc_layout(ah, 2048, 128, 0), desc_layout(bl, 128, 1024, 0), idesc, 1u);
                    da = desc_layout(ah, 2048, 128, 0);
                    db = desc_layout(bh, 128, 1024, 0);
                } else {
                    da = desc_layout(a + j * 32, 1, 1024, 2);
                    db = desc_layout(b + j * 32, 1, 1024, 2);
                }
                tc::mma_f16(tmem, da, db, idesc, (r | j) ? 1u : 0u);
            }
        }
        const long long t1 = clock64();
        tc::mma_commit(&done);
        tc::mbar_wait(&done, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    } else if (warp >= 1 && warp <= spinners) {
        tc::mbar_wait(&done, 0);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const char* names[5] = {"K-major / K-major interleaved (G1)", "A MN-major interleaved (G2)", "B MN-major interleaved (G3)",
                            "K-major / K-major 128B swizzle", "G2 pattern: hi/lo operands, 3 MMAs per K-step"};
    for (int mode = 0; mode < 5; ++mode)
        for (int N : {64, 128, 256}) {
            if (mode == 4 && N != 128) continue;
            for (int spinners : {0, 16}) {
                if (spinners && mode != 0 && mode != 4) continue;
                const int reps = 64, per = mode == 4 ? 12 : 4;
                umma_rate_kernel<<<1, 544, 160 * 1024>>>(mode, N, reps, spinners, out);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(cudaGetLastError())); return 1; }
                umma_rate_kernel<<<1, 544, 160 * 1024>>>(mode, N, reps, spinners, out);
                cudaDeviceSynchronize();
                printf("%-48s N = %3d, %2d warps polling: %6.1f cycles per MMA issued, %6.1f until complete (floor %d)\n", names[mode], N,
                       spinners, (double)out[0] / (per * reps), (double)out[1] / (per * reps), N / 2);
            }
        }
    return 0;
}
