// mlp_tc_tile.cuh -- inference forward of the actor-critic MLP (P4, model.py:16-46) for a tile of 32 samples with
// the 256 -> 128 layer (93 % of the FLOP) on the tensor cores.  Same interface as mlp_tile.cuh (which stays the
// training forward of csrc/ppo_kernels.cu): 256 threads, input tile in sm[x], logits / value in sm[out].
//
// Why warp-level mma.sync and not tcgen05: the rollout runs one 32-sample tile per SM and step (4096 envs / 128
// tiles), tcgen05.mma needs M >= 64 rows and a TMEM round trip per step; mma.sync.m16n8k16 takes the two 16-row
// halves of the tile directly from shared memory.  Measured on B200 (profiles/debug/mma): 0.46 m16n8k16 f16 MMAs
// per clock and SM = 950 MAC/clk/SM, against 128 FFMA lanes.
//
// fp32-grade accuracy from fp16 operands -- the two-term split with a scaled low part:
//     x = hi + 2^-11 lo,   hi = fp16(x),   lo = fp16((x - hi) * 2^11)          (|x - hi| <= 2^-11 |x|)
//     x w = hi_x hi_w + 2^-11 (hi_x lo_w + lo_x hi_w) + O(2^-22 x w)
// The scaling keeps lo in the normal fp16 range; the cross terms accumulate in their own fp32 accumulator and are
// folded in once at the end (the same idea as the separate TMEM accumulator of the tcgen05 update kernel).  Three
// MMAs per product: 317 effective MAC/clk/SM, 2.5x the FFMA rate; relative error ~2^-22 per product, i.e. the
// same order as the fp32 FMA chain's own rounding.  Checked against the torch fp32 forward at rel 1e-5
// (tests/test_gpu_policy.py).
//
// Shared-memory plan (float units; the fp16 arrays are addressed as half):
//   W1t [6][256]  P1 [3][256]  P2 [3][128]  Wh [128][8]  bh [8]  x [32][8]  red [2][8][16] (+ spare)  out [32][8]
//   stat [4][32]
//   W2h, W2l [128][264] half   feature.3.weight[o][k] split, row padded to 528 B (conflict-free ldmatrix rows)
//   Ah,  Al  [16][264]  half   layer-1 activations of the half tile in flight, split (the arrays keep room for 32
//                              rows); the region is reused for h2 [16][132] float
#pragma once
#include <cuda_fp16.h>

#include "mlp_tile.cuh"

namespace plume {

constexpr int kTcLd = 264;                        // halves per padded row (256 + 8)

struct MlpTcSmem {
    static constexpr int W1t = 0;
    static constexpr int P1 = W1t + 6 * 256;
    static constexpr int P2 = P1 + 3 * 256;
    static constexpr int Wh = P2 + 3 * 128;
    static constexpr int bh = Wh + 128 * 8;
    static constexpr int x = bh + 8;
    static constexpr int red = x + kTileM * 8;
    static constexpr int out = red + 2 * 8 * kTileM;
    static constexpr int stat = out + kTileM * 8;
    static constexpr int W2h = stat + 4 * kTileM;                 // 128 * 264 halves = 16896 floats
    static constexpr int W2l = W2h + 128 * kTcLd / 2;
    static constexpr int Ah = W2l + 128 * kTcLd / 2;              // 32 * 264 halves = 4224 floats
    static constexpr int Al = Ah + kTileM * kTcLd / 2;
    static constexpr int h2 = Ah;                                 // [32][132] floats, after the MMAs have read A
    static constexpr int total = Al + kTileM * kTcLd / 2;         // 47112 floats = 188 448 B
};
static_assert(MlpTcSmem::W2h % 4 == 0 && MlpTcSmem::Ah % 4 == 0, "ldmatrix rows must be 16-byte aligned");
static_assert(kTileM * kH2Stride <= 2 * kTileM * kTcLd / 2, "h2 must fit in the A operand region");

constexpr float kSplitScale = 2048.0f, kSplitInv = 1.0f / 2048.0f;

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * kSplitScale);
}

// One-time: flat parameters (include/plume_b200.h layout) -> shared memory.
__device__ __forceinline__ void mlp_tc_load_weights(float* sm, const float* __restrict__ p) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 6 * 256; i += kMlpThreads) {            // W1[o][k] -> W1t[k][o]
        const int o = i / 6, k = i - o * 6;
        sm[MlpTcSmem::W1t + k * 256 + o] = p[PLUME_OFF_W1 + i];
    }
    for (int i = tid; i < 256; i += kMlpThreads) {
        sm[MlpTcSmem::P1 + i] = p[PLUME_OFF_B1 + i];
        sm[MlpTcSmem::P1 + 256 + i] = p[PLUME_OFF_G1 + i];
        sm[MlpTcSmem::P1 + 512 + i] = p[PLUME_OFF_BE1 + i];
    }
    __half2* w2h = reinterpret_cast<__half2*>(sm + MlpTcSmem::W2h);
    __half2* w2l = reinterpret_cast<__half2*>(sm + MlpTcSmem::W2l);
    for (int i = tid; i < 128 * 128; i += kMlpThreads) {          // pairs (o, k), (o, k+1): coalesced float2 reads
        const int o = i >> 7, k2 = i & 127;
        const float2 w = *reinterpret_cast<const float2*>(p + PLUME_OFF_W2 + o * 256 + 2 * k2);
        __half h0, l0, h1, l1;
        split_f16(w.x, h0, l0);
        split_f16(w.y, h1, l1);
        w2h[o * (kTcLd / 2) + k2] = __halves2half2(h0, h1);
        w2l[o * (kTcLd / 2) + k2] = __halves2half2(l0, l1);
    }
    for (int i = tid; i < 128; i += kMlpThreads) {
        sm[MlpTcSmem::P2 + i] = p[PLUME_OFF_B2 + i];
        sm[MlpTcSmem::P2 + 128 + i] = p[PLUME_OFF_G2 + i];
        sm[MlpTcSmem::P2 + 256 + i] = p[PLUME_OFF_BE2 + i];
    }
    for (int i = tid; i < 128 * 8; i += kMlpThreads) {
        const int k = i >> 3, o = i & 7;
        float w = 0.0f;
        if (o < 5) w = p[PLUME_OFF_WA + o * 128 + k];
        else if (o == 5) w = p[PLUME_OFF_WC + k];
        sm[MlpTcSmem::Wh + i] = w;
    }
    if (tid < 8) sm[MlpTcSmem::bh + tid] = tid < 5 ? p[PLUME_OFF_BA + tid] : (tid == 5 ? p[PLUME_OFF_BC] : 0.0f);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// barrier of the 256 MLP threads: the whole CTA (__syncthreads) or, where the CTA has more warps than these eight
// (the pipelined rollout kernel), named barrier 1
template <bool kNamed>
__device__ __forceinline__ void mlp_sync() {
    if (kNamed) asm volatile("bar.sync 1, 256;" ::: "memory");
    else __syncthreads();
}

// Forward of HALF a tile: the 16 samples in rows 16*half .. 16*half+15 of sm[x] (rows beyond the valid ones must hold
// finite values, e.g. zeros) -> sm[out][same rows][0..4] = logits, [5] = value.  Executed by the 256 threads of warps
// 0..7.  The caller makes x visible before the call and synchronises before it reads sm[out]; between two calls no
// barrier is needed (the first barrier inside the next call orders the reuse of the A / h2 region).
// A 16-sample unit is what the pipelined rollout kernel steps at a time (one half's policy forward overlaps the other
// half's env step); every other user (policy_kernel, the rollout kernel with the in-loop stop head) runs two halves,
// so all paths compute bit-identical logits.
#ifdef PLUME_ROLLOUT_TIMELINE
#define PLUME_MLP_TL(n) mtl[n] = clock64()
static __device__ int g_mlp_tl_calls = 0;
#else
#define PLUME_MLP_TL(n)
#endif
template <bool kNamed>
__device__ __forceinline__ void mlp_tc_forward_half(float* sm, int half) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef PLUME_ROLLOUT_TIMELINE
    long long mtl[5];
#endif
    PLUME_MLP_TL(0);
    __half* Ah = reinterpret_cast<__half*>(sm + MlpTcSmem::Ah);
    __half* Al = reinterpret_cast<__half*>(sm + MlpTcSmem::Al);
    // ---- layer 1 on the CUDA cores: thread = (sample s, 16 outputs of block ob) -----------------------------
    {
        const int s = lane & 15, ob = 2 * warp + (lane >> 4);
        float2 xr2[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float v = sm[MlpTcSmem::x + (16 * half + s) * 8 + k];
            xr2[k] = make_float2(v, v);
        }
        // four outputs per step: float4 weight loads and packed fp32 pairs (FFMA2: two IEEE-rn FMAs per issue slot);
        // every output sums its six terms in the order k = 0..5, then the bias
        float z[16];
        float part = 0.0f;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const int o = ob * 16 + 4 * j4;
            float2 a01 = make_float2(0.0f, 0.0f), a23 = a01;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(sm + MlpTcSmem::W1t + k * 256 + o);
                a01 = __ffma2_rn(xr2[k], make_float2(w.x, w.y), a01);
                a23 = __ffma2_rn(xr2[k], make_float2(w.z, w.w), a23);
            }
            const float4 b = *reinterpret_cast<const float4*>(sm + MlpTcSmem::P1 + o);
            a01 = __fadd2_rn(a01, make_float2(b.x, b.y));
            a23 = __fadd2_rn(a23, make_float2(b.z, b.w));
            z[4 * j4 + 0] = a01.x;
            z[4 * j4 + 1] = a01.y;
            z[4 * j4 + 2] = a23.x;
            z[4 * j4 + 3] = a23.y;
            part += (a01.x + a01.y) + (a23.x + a23.y);
        }
        // LayerNorm-1 statistics with ONE exchange: every partial carries its sum and its sum of squares about its OWN
        // mean; partials merge with M2 = sum_g [M2_g + n_g (mean_g - mean)^2] (as accurate as the two-pass form)
        const float mean_t = part * (1.0f / 16.0f);
        float m2 = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float d = z[j] - mean_t;
            m2 = fmaf(d, d, m2);
        }
        {
            const float part_p = __shfl_xor_sync(0xffffffffu, part, 16), m2_p = __shfl_xor_sync(0xffffffffu, m2, 16);
            const float sum32 = part + part_p, mean32 = sum32 * (1.0f / 32.0f);
            const float da = mean_t - mean32, db = part_p * (1.0f / 16.0f) - mean32;
            m2 = (m2 + m2_p) + 16.0f * fmaf(da, da, db * db);
            part = sum32;
        }
        if (lane < 16) {
            sm[MlpTcSmem::red + warp * 16 + s] = part;
            sm[MlpTcSmem::red + 128 + warp * 16 + s] = m2;
        }
        mlp_sync<kNamed>();
        PLUME_MLP_TL(1);
        float mean = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) mean += sm[MlpTcSmem::red + w * 16 + s];
        mean *= (1.0f / 256.0f);
        float var = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float dm = sm[MlpTcSmem::red + w * 16 + s] * (1.0f / 32.0f) - mean;
            var += fmaf(32.0f * dm, dm, sm[MlpTcSmem::red + 128 + w * 16 + s]);
        }
        const float rstd = 1.0f / sqrtf(var * (1.0f / 256.0f) + kLnEps);
        // relu(LN) -> fp16 hi / scaled lo, 8 values = one 16-byte store per array (conflict-free per quarter warp)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            __align__(16) __half2 hh[4], ll[4];
            float gq[8], bq[8];
#pragma unroll
            for (int v4 = 0; v4 < 2; ++v4) {
                const float4 gv = *reinterpret_cast<const float4*>(sm + MlpTcSmem::P1 + 256 + ob * 16 + 8 * q + 4 * v4);
                const float4 bv = *reinterpret_cast<const float4*>(sm + MlpTcSmem::P1 + 512 + ob * 16 + 8 * q + 4 * v4);
                gq[4 * v4] = gv.x; gq[4 * v4 + 1] = gv.y; gq[4 * v4 + 2] = gv.z; gq[4 * v4 + 3] = gv.w;
                bq[4 * v4] = bv.x; bq[4 * v4 + 1] = bv.y; bq[4 * v4 + 2] = bv.z; bq[4 * v4 + 3] = bv.w;
            }
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
                float y[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = 8 * q + 2 * j2 + e;
                    y[e] = fmaxf((z[j] - mean) * rstd * gq[2 * j2 + e] + bq[2 * j2 + e], 0.0f);
                }
                __half h0, l0, h1, l1;
                split_f16(y[0], h0, l0);
                split_f16(y[1], h1, l1);
                hh[j2] = __halves2half2(h0, h1);
                ll[j2] = __halves2half2(l0, l1);
            }
            const int off = s * kTcLd + ob * 16 + 8 * q;
            *reinterpret_cast<uint4*>(Ah + off) = *reinterpret_cast<const uint4*>(hh);
            *reinterpret_cast<uint4*>(Al + off) = *reinterpret_cast<const uint4*>(ll);
        }
    }
    mlp_sync<kNamed>();
    PLUME_MLP_TL(2);
    // ---- layer 2 on the tensor cores: warp = 16 samples x 16 outputs, K = 256 in 16 steps of 3 x 2 MMAs ----
    float zacc[2][4];             // [n tile][fragment]: z2 without the bias
    {
        const __half* W2h = reinterpret_cast<const __half*>(sm + MlpTcSmem::W2h);
        const __half* W2l = reinterpret_cast<const __half*>(sm + MlpTcSmem::W2l);
        float cm[2][4], cs[2][4], ct[2][4];       // main terms; hi x lo and lo x hi cross terms in separate chains
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) cm[nt][e] = cs[nt][e] = ct[nt][e] = 0.0f;
        const int mj = lane >> 3, mi = lane & 7;        // ldmatrix: this lane addresses row mi of matrix mj
        // A fragment matrices: 0 = rows 0-7 / k 0-7, 1 = rows 8-15 / k 0-7, 2 = rows 0-7 / k 8-15, 3 = rows 8-15 / k 8-15
        const int a_off = ((mj & 1) * 8 + mi) * kTcLd + (mj >> 1) * 8;
        // B matrices: 0 = n 0-7 / k 0-7, 1 = n 0-7 / k 8-15, 2 = n 8-15 / k 0-7, 3 = n 8-15 / k 8-15
        const int b_off = (16 * warp + (mj >> 1) * 8 + mi) * kTcLd + (mj & 1) * 8;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
            const int k0 = 16 * ks;
            uint32_t ah[4], al[4], bh[4], bl[4];
            ldmatrix_x4(bh, W2h + b_off + k0);
            ldmatrix_x4(bl, W2l + b_off + k0);
            ldmatrix_x4(ah, Ah + a_off + k0);
            ldmatrix_x4(al, Al + a_off + k0);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                mma_f16(cm[nt], ah, bh[2 * nt], bh[2 * nt + 1]);
                mma_f16(cs[nt], ah, bl[2 * nt], bl[2 * nt + 1]);
                mma_f16(ct[nt], al, bh[2 * nt], bh[2 * nt + 1]);
            }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) zacc[nt][e] = fmaf(cs[nt][e] + ct[nt][e], kSplitInv, cm[nt][e]);
    }
    PLUME_MLP_TL(3);
    // ---- bias, LayerNorm-2, ReLU: fragment (nt, e) = row 8 (e >> 1) + g, column 16 warp + 8 nt + 2 t + (e & 1)
    const int g = lane >> 2, t = lane & 3;
    float y2[2][4];               // relu(LN2(z2)) in the accumulator layout
    {
        float part[2] = {0.0f, 0.0f};       // rows g and 8 + g
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                zacc[nt][e] += sm[MlpTcSmem::P2 + 16 * warp + 8 * nt + 2 * t + (e & 1)];
                part[e >> 1] += zacc[nt][e];
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            part[r] += __shfl_xor_sync(0xffffffffu, part[r], 1);
            part[r] += __shfl_xor_sync(0xffffffffu, part[r], 2);
        }
        // one exchange (see LayerNorm-1): the warp's 16 columns of a row carry their sum and their M2 about their mean
        float sq[2] = {0.0f, 0.0f};
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = zacc[nt][e] - part[e >> 1] * (1.0f / 16.0f);
                sq[e >> 1] = fmaf(d, d, sq[e >> 1]);
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], 1);
            sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], 2);
        }
        if (t == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                sm[MlpTcSmem::red + warp * 16 + 8 * r + g] = part[r];
                sm[MlpTcSmem::red + 128 + warp * 16 + 8 * r + g] = sq[r];
            }
        }
        mlp_sync<kNamed>();       // also: every warp has finished reading the A operands (the head partials alias them)
        float mean[2], rstd[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float m = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) m += sm[MlpTcSmem::red + w * 16 + 8 * r + g];
            mean[r] = m * (1.0f / 128.0f);
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const float dm = sm[MlpTcSmem::red + w * 16 + 8 * r + g] * (1.0f / 16.0f) - mean[r];
                v += fmaf(16.0f * dm, dm, sm[MlpTcSmem::red + 128 + w * 16 + 8 * r + g]);
            }
            rstd[r] = 1.0f / sqrtf(v * (1.0f / 128.0f) + kLnEps);
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = 16 * warp + 8 * nt + 2 * t + (e & 1);
                y2[nt][e] = fmaxf((zacc[nt][e] - mean[e >> 1]) * rstd[e >> 1] * sm[MlpTcSmem::P2 + 128 + col] +
                                      sm[MlpTcSmem::P2 + 256 + col], 0.0f);
            }
    }
    PLUME_MLP_TL(4);
    // ---- heads on the tensor cores, straight from the registers: the warp's 16 columns of h2 are the K slice
    // [16 warp, 16 warp + 16) of the 128 -> 6 head GEMM, and the accumulator layout of two n tiles IS the A fragment
    // layout of one m16n8k16 step (a0 = (g, 2t..) of n tile 0, a1 = (g+8, ..) of n tile 0, a2 / a3 = n tile 1).
    // One split MMA triple per warp, the eight K-slice partials are added through shared memory.
    {
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                __half h0, l0, h1, l1;
                split_f16(y2[nt][2 * r], h0, l0);
                split_f16(y2[nt][2 * r + 1], h1, l1);
                const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
                ahi[2 * nt + r] = *reinterpret_cast<const uint32_t*>(&hh);
                alo[2 * nt + r] = *reinterpret_cast<const uint32_t*>(&ll);
            }
        // B fragment: b0 = (k = 2t, 2t+1; n = g), b1 = (k = 2t+8, 2t+9; n = g) of Wh[16 warp + k][n] (n = 6, 7 are zero)
        uint32_t bhi[2], blo[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float w0 = sm[MlpTcSmem::Wh + (16 * warp + 8 * q + 2 * t) * 8 + g];
            const float w1 = sm[MlpTcSmem::Wh + (16 * warp + 8 * q + 2 * t + 1) * 8 + g];
            __half h0, l0, h1, l1;
            split_f16(w0, h0, l0);
            split_f16(w1, h1, l1);
            const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
            bhi[q] = *reinterpret_cast<const uint32_t*>(&hh);
            blo[q] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        float cmh[4] = {0.0f, 0.0f, 0.0f, 0.0f}, csh[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        mma_f16(cmh, ahi, bhi[0], bhi[1]);
        mma_f16(csh, ahi, blo[0], blo[1]);
        mma_f16(csh, alo, bhi[0], bhi[1]);
        float* pout = sm + MlpTcSmem::h2;          // [8 warps][16 rows][8]: partial head outputs of the K slices
#pragma unroll
        for (int r = 0; r < 2; ++r)
            *reinterpret_cast<float2*>(pout + (warp * 16 + 8 * r + g) * 8 + 2 * t) =
                make_float2(fmaf(csh[2 * r], kSplitInv, cmh[2 * r]), fmaf(csh[2 * r + 1], kSplitInv, cmh[2 * r + 1]));
        mlp_sync<kNamed>();
        if (tid < 128) {
            const int s = tid >> 3, n = tid & 7;
            float acc = sm[MlpTcSmem::bh + n];
#pragma unroll
            for (int w = 0; w < 8; ++w) acc += pout[(w * 16 + s) * 8 + n];
            sm[MlpTcSmem::out + (16 * half + s) * 8 + n] = acc;
        }
    }
#ifdef PLUME_ROLLOUT_TIMELINE
    if (blockIdx.x == 0 && tid == 0 && atomicAdd(&g_mlp_tl_calls, 1) % 3001 == 700)
        printf("mlp half timeline: layer 1 + LN1 stats %lld | normalise + split + barrier %lld | layer 2 %lld | LN2 %lld | heads %lld\n",
               mtl[1] - mtl[0], mtl[2] - mtl[1], mtl[3] - mtl[2], mtl[4] - mtl[3], clock64() - mtl[4]);
#endif
}

// Forward of the 32-sample tile in sm[x]: two halves.  Every thread of the 256-thread CTA must call it; ends with
// __syncthreads().
__device__ __forceinline__ void mlp_tc_forward_tile(float* sm) {
    __syncthreads();   // x tile visible; the previous call's h2 (aliasing A) has been consumed
    mlp_tc_forward_half<false>(sm, 0);
    mlp_tc_forward_half<false>(sm, 1);
    __syncthreads();
}

// The inference tile the rollout and the policy kernels use.  -DPLUME_MLP_FFMA builds them on the fp32 CUDA-core
// tile of mlp_tile.cuh instead (kept for A/B measurements; the training forward always uses that one).
#ifdef PLUME_MLP_FFMA
using PolicySmem = MlpSmem;
__device__ __forceinline__ void policy_load_weights(float* sm, const float* __restrict__ p) { mlp_load_weights(sm, p); }
__device__ __forceinline__ void policy_forward_tile(float* sm) { mlp_forward_tile<false>(sm); }
#else
using PolicySmem = MlpTcSmem;
__device__ __forceinline__ void policy_load_weights(float* sm, const float* __restrict__ p) { mlp_tc_load_weights(sm, p); }
__device__ __forceinline__ void policy_forward_tile(float* sm) { mlp_tc_forward_tile(sm); }
#endif

}  // namespace plume
