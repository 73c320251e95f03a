// tc_gemm_kernels.cu -- C[M][N] = A[M][K] . B[N][K]^T in fp32-grade 3xTF32 on the sm_100a tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM).  The stand-alone GEMM is the unit test of
// the building blocks in tc_gemm.cuh that the PPO update kernels use.
//
// One CTA (128 threads) per 128-row tile; K streamed in 32-wide chunks through a 2-stage shared-
// memory ring: all threads load + split a chunk while the previous chunk's MMAs run asynchronously;
// thread 0 issues the MMAs and commits them to the stage's mbarrier.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace plume {

template <int BN>
__global__ void __launch_bounds__(128, 1) tc_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ C, int M, int K) {
    extern __shared__ __align__(1024) float sm[];
    constexpr int kA = 128 * tc::kChunkK, kB = BN * tc::kChunkK;
    float* a_hi = sm;                  // [2][kA]
    float* a_lo = a_hi + 2 * kA;
    float* b_hi = a_lo + 2 * kA;       // [2][kB]
    float* b_lo = b_hi + 2 * kB;
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    if (tid == 0) {
        tc::mbar_init(&mma_done[0], 1);
        tc::mbar_init(&mma_done[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<BN>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = tc::make_idesc_tf32(128, BN);

    const int chunks = K / tc::kChunkK;
    for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc & 1;
        if (kc >= 2) tc::mbar_wait(&mma_done[s], (uint32_t)(((kc >> 1) - 1) & 1));   // stage s is free again
        tc::load_split_chunk<128, 128>(a_hi + s * kA, a_lo + s * kA, A + (size_t)m0 * K + kc * tc::kChunkK, K, M - m0, tid);
        tc::load_split_chunk<BN, 128>(b_hi + s * kB, b_lo + s * kB, B + kc * tc::kChunkK, K, BN, tid);
        tc::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            tc::mma_chunk_3xtf32(tmem_d, a_hi + s * kA, a_lo + s * kA, b_hi + s * kB, b_lo + s * kB, idesc, kc == 0);
            tc::mma_commit(&mma_done[s]);
        }
    }
    // all MMAs complete in order: wait for the last chunk's commit
    {
        const int last = chunks - 1;
        tc::mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
    }
    tc::tc_fence_after();
    // epilogue: warp w owns TMEM lanes 32w..32w+31 = rows m0+32w+lane
    const int row = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        tc::tmem_ld_wait();
        if (row < M) {
            float4* dst = reinterpret_cast<float4*>(C + (size_t)row * BN + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<BN>(tmem_d);
}

template <int BN>
static int launch_tc_gemm(const float* A, const float* B, float* C, int M, int K, cudaStream_t s) {
    const int smem = (2 * 2 * 128 * tc::kChunkK + 2 * 2 * BN * tc::kChunkK) * (int)sizeof(float) + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("tc_gemm: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    tc_gemm_kernel<BN><<<(M + 127) / 128, 128, smem, s>>>(A, B, C, M, K);
    if (cudaGetLastError() != cudaSuccess) return fail("tc_gemm launch failed");
    return 0;
}

// The same GEMM on kind::f16 with the two-term split (tc_gemm.cuh): 64-wide chunks, main and cross terms in two TMEM
// accumulators (columns [0,BN) and [BN,2BN)), combined in the epilogue.  kScaled = false keeps lo unscaled and
// everything in one accumulator (the form the backward GEMMs of the update kernel use).
template <int BN, bool kScaled>
__global__ void __launch_bounds__(128, 1) tc_gemm_f16_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ C, int M, int K) {
    extern __shared__ __align__(1024) float sm[];
    constexpr int kA = 128 * tc::kChunkK, kB = BN * tc::kChunkK;      // floats: a 64-half chunk row is 128 bytes too
    float* a_hi = sm;
    float* a_lo = a_hi + 2 * kA;
    float* b_hi = a_lo + 2 * kA;
    float* b_lo = b_hi + 2 * kB;
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t kCols = kScaled ? 2 * BN : BN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    if (tid == 0) {
        tc::mbar_init(&mma_done[0], 1);
        tc::mbar_init(&mma_done[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<kCols>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = tc::make_idesc_f16(128, BN);
    const float ls = kScaled ? tc::kLoScale : 1.0f;

    const int chunks = K / tc::kChunkKH;
    for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc & 1;
        if (kc >= 2) tc::mbar_wait(&mma_done[s], (uint32_t)(((kc >> 1) - 1) & 1));
        tc::load_split_chunk_f16<128, 128>(a_hi + s * kA, a_lo + s * kA, A + (size_t)m0 * K + kc * tc::kChunkKH, K, M - m0,
                                           tid, ls);
        tc::load_split_chunk_f16<BN, 128>(b_hi + s * kB, b_lo + s * kB, B + kc * tc::kChunkKH, K, BN, tid, ls);
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            if (kScaled)
                tc::mma_chunk_f16_split(tmem_d, tmem_d + BN, a_hi + s * kA, a_lo + s * kA, b_hi + s * kB, b_lo + s * kB,
                                        idesc, kc == 0);
            else
                tc::mma_chunk_f16(tmem_d, a_hi + s * kA, a_lo + s * kA, b_hi + s * kB, b_lo + s * kB, idesc, kc == 0);
            tc::mma_commit(&mma_done[s]);
        }
    }
    {
        const int last = chunks - 1;
        tc::mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
    }
    tc::tc_fence_after();
    const int row = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        tc::tmem_ld_wait();
        if (kScaled) {
            float w[32];
            tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(BN + c * 32), w);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(w[j], tc::kLoInv, v[j]);
        }
        if (row < M) {
            float4* dst = reinterpret_cast<float4*>(C + (size_t)row * BN + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<kCols>(tmem_d);
}

// The unscaled-lo GEMM with A staged MN-major (unit test of the descriptor form the update kernel's G2 uses): per
// 64-wide K chunk, slot (kg, mg, k & 7) at kg * 2048 + mg * 128 + (k & 7) * 16 bytes holds A[8 mg .. 8 mg + 7][k].
template <int BN>
__global__ void __launch_bounds__(128, 1) tc_gemm_f16_amn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                 float* __restrict__ C, int M, int K) {
    extern __shared__ __align__(1024) float sm[];
    constexpr int kA = 128 * tc::kChunkK, kB = BN * tc::kChunkK;
    float* a_hi = sm;
    float* a_lo = a_hi + 2 * kA;
    float* b_hi = a_lo + 2 * kA;
    float* b_lo = b_hi + 2 * kB;
    __shared__ uint64_t mma_done[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    if (tid == 0) {
        tc::mbar_init(&mma_done[0], 1);
        tc::mbar_init(&mma_done[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<BN>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = tc::make_idesc_f16_a_mn(128, BN);
    const int chunks = K / tc::kChunkKH;
    for (int kc = 0; kc < chunks; ++kc) {
        const int s = kc & 1;
        if (kc >= 2) tc::mbar_wait(&mma_done[s], (uint32_t)(((kc >> 1) - 1) & 1));
        uint4* ah = reinterpret_cast<uint4*>(a_hi + s * kA);
        uint4* al = reinterpret_cast<uint4*>(a_lo + s * kA);
        for (int f = tid; f < 1024; f += 128) {               // slot f = kg * 128 + mg * 8 + (k & 7)
            const int kg = f >> 7, mg = (f >> 3) & 15, k = kc * tc::kChunkKH + kg * 8 + (f & 7);
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = m0 + 8 * mg + i;
                v[i] = row < M ? A[(size_t)row * K + k] : 0.0f;
            }
            uint4 hi, lo;
            tc::split_f16x8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), 1.0f, hi, lo);
            ah[f] = hi;
            al[f] = lo;
        }
        tc::load_split_chunk_f16<BN, 128>(b_hi + s * kB, b_lo + s * kB, B + kc * tc::kChunkKH, K, BN, tid, 1.0f);
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            const uint32_t sah = tc::smem_u32(ah), sal = tc::smem_u32(al);
            const uint32_t sbh = tc::smem_u32(b_hi + s * kB), sbl = tc::smem_u32(b_lo + s * kB);
#pragma unroll
            for (int j = 0; j < tc::kChunkKH / 16; ++j) {
                const uint32_t aoff = j * 2 * 2048, boff = j * 2 * tc::kLBO;       // two core matrices along K per step
                const uint64_t dah = tc::make_smem_desc(sah + aoff, 2048, 128), dal = tc::make_smem_desc(sal + aoff, 2048, 128);
                const uint64_t dbh = tc::make_smem_desc(sbh + boff, tc::kLBO, tc::kSBO);
                const uint64_t dbl = tc::make_smem_desc(sbl + boff, tc::kLBO, tc::kSBO);
                tc::mma_f16(tmem_d, dal, dbh, idesc, (kc == 0 && j == 0) ? 0u : 1u);
                tc::mma_f16(tmem_d, dah, dbl, idesc, 1u);
                tc::mma_f16(tmem_d, dah, dbh, idesc, 1u);
            }
            tc::mma_commit(&mma_done[s]);
        }
    }
    {
        const int last = chunks - 1;
        tc::mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
    }
    tc::tc_fence_after();
    const int row = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        tc::tmem_ld_wait();
        if (row < M) {
            float4* dst = reinterpret_cast<float4*>(C + (size_t)row * BN + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<BN>(tmem_d);
}

template <int BN>
static int launch_tc_gemm_f16_amn(const float* A, const float* B, float* C, int M, int K, cudaStream_t s) {
    const int smem = (2 * 2 * 128 * tc::kChunkK + 2 * 2 * BN * tc::kChunkK) * (int)sizeof(float) + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_gemm_f16_amn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("tc_gemm_f16 (A MN-major): cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    tc_gemm_f16_amn_kernel<BN><<<(M + 127) / 128, 128, smem, s>>>(A, B, C, M, K);
    if (cudaGetLastError() != cudaSuccess) return fail("tc_gemm_f16 (A MN-major) launch failed");
    return 0;
}


// The unscaled-lo GEMM C[M][64] = A[M][K] . B[64][K]^T in the form the update kernel's G3 uses (K a multiple of 128):
// A staged K-major with the strides of the resident dz2 operand (slot (mg, kg, m & 7) at mg * 2048 + kg * 128 +
// (m & 7) * 16 bytes holds A[m][8 kg .. 8 kg + 7]: LBO 128, SBO 2048), B staged MN-major in the layout of a forward
// activation chunk (slot (kg, ng, k & 7) at kg * 1024 + ng * 128 + (k & 7) * 16 bytes holds B[8 ng .. 8 ng + 7][k]:
// instruction-descriptor bit 16, LBO 1024 = next 8 k, SBO 128 = next 8 n) -- and B makes the round trip the update
// kernel's activation stash makes: written to shared memory, cp.async.bulk to global memory (the tile's own rows of C
// serve as the 32 KB scratch), cp.async.bulk back into a second buffer, consumed by the MMAs from there.
__global__ void __launch_bounds__(128, 1) tc_gemm_f16_bmn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                 float* __restrict__ C, int M, int K) {
    extern __shared__ __align__(1024) float sm[];
    uint4* a_hi = reinterpret_cast<uint4*>(sm);              // 2048 slots (32 KB)
    uint4* a_lo = a_hi + 2048;
    uint4* b_src = a_lo + 2048;                              // hi 1024 slots, lo 1024 slots (32 KB): written by the threads
    uint4* b_dst = b_src + 2048;                             // the copy that came back from global memory
    __shared__ uint64_t mma_done, landed;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    if (tid == 0) {
        tc::mbar_init(&mma_done, 1);
        tc::mbar_init(&landed, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<64>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = tc::make_idesc_f16_b_mn(128, 64);
    char* scratch = reinterpret_cast<char*>(C + (size_t)m0 * 64);      // 128 rows x 64 floats = 32 KB
    const int chunks = K / 128;
    for (int kc = 0; kc < chunks; ++kc) {
        if (kc >= 1) tc::mbar_wait(&mma_done, (uint32_t)((kc - 1) & 1));
        for (int f = tid; f < 2048; f += 128) {               // A slot f = mg * 128 + kg * 8 + (m & 7)
            const int mg = f >> 7, kg = (f >> 3) & 15, row = m0 + 8 * mg + (f & 7);
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (row < M) {
                v0 = *reinterpret_cast<const float4*>(A + (size_t)row * K + kc * 128 + 8 * kg);
                v1 = *reinterpret_cast<const float4*>(A + (size_t)row * K + kc * 128 + 8 * kg + 4);
            }
            uint4 hi, lo;
            tc::split_f16x8(v0, v1, 1.0f, hi, lo);
            a_hi[f] = hi;
            a_lo[f] = lo;
        }
        for (int f = tid; f < 1024; f += 128) {               // B slot f = kg * 64 + ng * 8 + (k & 7)
            const int kg = f >> 6, ng = (f >> 3) & 7, k = kc * 128 + 8 * kg + (f & 7);
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = B[(size_t)(8 * ng + i) * K + k];
            uint4 hi, lo;
            tc::split_f16x8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), 1.0f, hi, lo);
            b_src[f] = hi;
            b_src[1024 + f] = lo;
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::bulk_store(scratch, b_src, 32768);
            tc::bulk_commit_group();
            tc::bulk_wait_group_all();
            tc::bulk_load(b_dst, scratch, 32768, &landed);
            tc::mbar_wait(&landed, (uint32_t)(kc & 1));
            tc::tc_fence_after();
            const uint32_t sah = tc::smem_u32(a_hi), sal = tc::smem_u32(a_lo);
            const uint32_t sbh = tc::smem_u32(b_dst), sbl = tc::smem_u32(b_dst + 1024);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t aoff = j * 2 * 128, boff = j * 2 * 1024;
                const uint64_t dah = tc::make_smem_desc(sah + aoff, 128, 2048), dal = tc::make_smem_desc(sal + aoff, 128, 2048);
                const uint64_t dbh = tc::make_smem_desc(sbh + boff, 1024, 128), dbl = tc::make_smem_desc(sbl + boff, 1024, 128);
                tc::mma_f16(tmem_d, dal, dbh, idesc, (kc == 0 && j == 0) ? 0u : 1u);
                tc::mma_f16(tmem_d, dah, dbl, idesc, 1u);
                tc::mma_f16(tmem_d, dah, dbh, idesc, 1u);
            }
            tc::mma_commit(&mma_done);
        }
    }
    tc::mbar_wait(&mma_done, (uint32_t)((chunks - 1) & 1));
    tc::tc_fence_after();
    const int row = m0 + warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        tc::tmem_ld_wait();
        if (row < M) {
            float4* dst = reinterpret_cast<float4*>(C + (size_t)row * 64 + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(tmem_d);
}

static int launch_tc_gemm_f16_bmn(const float* A, const float* B, float* C, int M, int K, cudaStream_t s) {
    const int smem = 4 * 32768 + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_gemm_f16_bmn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("tc_gemm_f16 (B MN-major): cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    tc_gemm_f16_bmn_kernel<<<(M + 127) / 128, 128, smem, s>>>(A, B, C, M, K);
    if (cudaGetLastError() != cudaSuccess) return fail("tc_gemm_f16 (B MN-major) launch failed");
    return 0;
}

template <int BN, bool kScaled>
static int launch_tc_gemm_f16(const float* A, const float* B, float* C, int M, int K, cudaStream_t s) {
    const int smem = (2 * 2 * 128 * tc::kChunkK + 2 * 2 * BN * tc::kChunkK) * (int)sizeof(float) + 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_gemm_f16_kernel<BN, kScaled>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return fail("tc_gemm_f16: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    tc_gemm_f16_kernel<BN, kScaled><<<(M + 127) / 128, 128, smem, s>>>(A, B, C, M, K);
    if (cudaGetLastError() != cudaSuccess) return fail("tc_gemm_f16 launch failed");
    return 0;
}

}  // namespace plume

using namespace plume;

extern "C" int plume_tc_gemm_f16(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K,
                                 int32_t scaled_lo, void* stream) {
    PLUME_CHECK_ARG(A && B && C, "null pointer");
    PLUME_CHECK_ARG(M > 0 && K > 0 && K % tc::kChunkKH == 0, "K must be a positive multiple of 64");
    if (scaled_lo == 3) {             // B staged MN-major through the bulk-copy round trip (N = 64, K % 128 == 0; M % 128 == 0:
        // the tile's rows of C are the scratch)
        PLUME_CHECK_ARG(N == 64 && K % 128 == 0 && M % 128 == 0, "B MN-major form: N = 64, K and M multiples of 128");
        return launch_tc_gemm_f16_bmn(A, B, C, M, K, as_stream(stream));
    }
    if (scaled_lo == 2) {             // A staged MN-major, unscaled lo
        if (N == 128) return launch_tc_gemm_f16_amn<128>(A, B, C, M, K, as_stream(stream));
        if (N == 256) return launch_tc_gemm_f16_amn<256>(A, B, C, M, K, as_stream(stream));
        return fail("plume_tc_gemm_f16: N must be 128 or 256");
    }
    if (N == 128 && scaled_lo) return launch_tc_gemm_f16<128, true>(A, B, C, M, K, as_stream(stream));
    if (N == 128) return launch_tc_gemm_f16<128, false>(A, B, C, M, K, as_stream(stream));
    if (N == 256 && scaled_lo) return launch_tc_gemm_f16<256, true>(A, B, C, M, K, as_stream(stream));
    if (N == 256) return launch_tc_gemm_f16<256, false>(A, B, C, M, K, as_stream(stream));
    return fail("plume_tc_gemm_f16: N must be 128 or 256");
}

extern "C" int plume_tc_gemm(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream) {
    PLUME_CHECK_ARG(A && B && C, "null pointer");
    PLUME_CHECK_ARG(M > 0 && K > 0 && K % tc::kChunkK == 0, "K must be a positive multiple of 32");
    if (N == 128) return launch_tc_gemm<128>(A, B, C, M, K, as_stream(stream));
    if (N == 256) return launch_tc_gemm<256>(A, B, C, M, K, as_stream(stream));
    return fail("plume_tc_gemm: N must be 128 or 256");
}
