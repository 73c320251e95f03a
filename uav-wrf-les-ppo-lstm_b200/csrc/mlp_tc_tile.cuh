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
//   W1t [6][256]  P1 [3][256]  P2 [3][128]  Wh [128][8]  bh [8]  x [32][8]  red [2][8][32]  out [32][8]  stat [4][32]
//   W2h, W2l [128][264] half   feature.3.weight[o][k] split, row padded to 528 B (conflict-free ldmatrix rows)
//   Ah,  Al  [32][264]  half   layer-1 activations split; the region is reused for h2 [32][132] float
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

// Forward of the tile in sm[x] (rows beyond the valid ones must hold finite values, e.g. zeros).  On return
// sm[out][s][0..4] = logits, [5] = value.  Every thread of the 256-thread CTA must call it; ends with __syncthreads().
__device__ __forceinline__ void mlp_tc_forward_tile(float* sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __half* Ah = reinterpret_cast<__half*>(sm + MlpTcSmem::Ah);
    __half* Al = reinterpret_cast<__half*>(sm + MlpTcSmem::Al);
    __syncthreads();   // x tile visible; the previous call's h2 (aliasing A) has been consumed
    // ---- layer 1 on the CUDA cores: thread = (sample lane, 32 outputs of chunk `warp`) ---------------------
    {
        float xr[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) xr[k] = sm[MlpTcSmem::x + lane * 8 + k];
        // four outputs per step: float4 weight loads (warp-uniform addresses) and packed fp32 pairs (FFMA2: two
        // IEEE-rn FMAs per issue slot); every output still sums its six terms in the order k = 0..5, then the bias
        float z[32];
        float part = 0.0f;
        float2 xr2[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) xr2[k] = make_float2(xr[k], xr[k]);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const int o = warp * 32 + 4 * j4;
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
            part += a01.x;
            part += a01.y;
            part += a23.x;
            part += a23.y;
        }
        sm[MlpTcSmem::red + warp * 32 + lane] = part;
        __syncthreads();
        float mean = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) mean += sm[MlpTcSmem::red + w * 32 + lane];
        mean *= (1.0f / 256.0f);
        float sq = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float d = z[j] - mean;
            sq = fmaf(d, d, sq);
        }
        sm[MlpTcSmem::red + 256 + warp * 32 + lane] = sq;
        __syncthreads();
        float var = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) var += sm[MlpTcSmem::red + 256 + w * 32 + lane];
        const float rstd = 1.0f / sqrtf(var * (1.0f / 256.0f) + kLnEps);
        // relu(LN) -> fp16 hi / scaled lo, 8 values = one 16-byte store per array (conflict-free per quarter warp)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __align__(16) __half2 hh[4], ll[4];
            float gq[8], bq[8];           // LayerNorm-1 gamma / beta of the 8 outputs of this slot: four float4 loads
#pragma unroll
            for (int v4 = 0; v4 < 2; ++v4) {
                const float4 gv = *reinterpret_cast<const float4*>(sm + MlpTcSmem::P1 + 256 + warp * 32 + 8 * q + 4 * v4);
                const float4 bv = *reinterpret_cast<const float4*>(sm + MlpTcSmem::P1 + 512 + warp * 32 + 8 * q + 4 * v4);
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
            const int off = lane * kTcLd + warp * 32 + 8 * q;
            *reinterpret_cast<uint4*>(Ah + off) = *reinterpret_cast<const uint4*>(hh);
            *reinterpret_cast<uint4*>(Al + off) = *reinterpret_cast<const uint4*>(ll);
        }
    }
    __syncthreads();
    // ---- layer 2 on the tensor cores: warp = 32 samples x 16 outputs, K = 256 in 16 steps of 3 x 4 MMAs ----
    float zacc[2][2][4];          // [m tile][n tile][fragment]: z2 without the bias
    {
        const __half* W2h = reinterpret_cast<const __half*>(sm + MlpTcSmem::W2h);
        const __half* W2l = reinterpret_cast<const __half*>(sm + MlpTcSmem::W2l);
        float cm[2][2][4], cs[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) cm[mt][nt][e] = cs[mt][nt][e] = 0.0f;
        const int mj = lane >> 3, mi = lane & 7;        // ldmatrix: this lane addresses row mi of matrix mj
        // A fragment matrices: 0 = rows 0-7 / k 0-7, 1 = rows 8-15 / k 0-7, 2 = rows 0-7 / k 8-15, 3 = rows 8-15 / k 8-15
        const int a_off = ((mj & 1) * 8 + mi) * kTcLd + (mj >> 1) * 8;
        // B matrices: 0 = n 0-7 / k 0-7, 1 = n 0-7 / k 8-15, 2 = n 8-15 / k 0-7, 3 = n 8-15 / k 8-15
        const int b_off = (16 * warp + (mj >> 1) * 8 + mi) * kTcLd + (mj & 1) * 8;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
            const int k0 = 16 * ks;
            uint32_t ah[2][4], al[2][4], bh[4], bl[4];
            ldmatrix_x4(bh, W2h + b_off + k0);
            ldmatrix_x4(bl, W2l + b_off + k0);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                ldmatrix_x4(ah[mt], Ah + 16 * mt * kTcLd + a_off + k0);
                ldmatrix_x4(al[mt], Al + 16 * mt * kTcLd + a_off + k0);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    mma_f16(cm[mt][nt], ah[mt], bh[2 * nt], bh[2 * nt + 1]);
                    mma_f16(cs[mt][nt], ah[mt], bl[2 * nt], bl[2 * nt + 1]);
                    mma_f16(cs[mt][nt], al[mt], bh[2 * nt], bh[2 * nt + 1]);
                }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) zacc[mt][nt][e] = fmaf(cs[mt][nt][e], kSplitInv, cm[mt][nt][e]);
    }
    // ---- bias, LayerNorm-2, ReLU: fragment (mt, nt, e) = row 16 mt + 8 (e >> 1) + g, column 16 warp + 8 nt + 2 t + (e & 1)
    {
        const int g = lane >> 2, t = lane & 3;
        float part[4];            // rows: [mt][e >> 1]
#pragma unroll
        for (int r = 0; r < 4; ++r) part[r] = 0.0f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    zacc[mt][nt][e] += sm[MlpTcSmem::P2 + 16 * warp + 8 * nt + 2 * t + (e & 1)];
                    part[2 * mt + (e >> 1)] += zacc[mt][nt][e];
                }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            part[r] += __shfl_xor_sync(0xffffffffu, part[r], 1);
            part[r] += __shfl_xor_sync(0xffffffffu, part[r], 2);
        }
        if (t == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) sm[MlpTcSmem::red + warp * 32 + 16 * (r >> 1) + 8 * (r & 1) + g] = part[r];
        }
        __syncthreads();          // also: every warp has finished reading the A operands (h2 aliases them)
        float mean[4], sq[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = 16 * (r >> 1) + 8 * (r & 1) + g;
            float m = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) m += sm[MlpTcSmem::red + w * 32 + row];
            mean[r] = m * (1.0f / 128.0f);
            sq[r] = 0.0f;
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = zacc[mt][nt][e] - mean[2 * mt + (e >> 1)];
                    sq[2 * mt + (e >> 1)] = fmaf(d, d, sq[2 * mt + (e >> 1)]);
                }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], 1);
            sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], 2);
        }
        if (t == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) sm[MlpTcSmem::red + 256 + warp * 32 + 16 * (r >> 1) + 8 * (r & 1) + g] = sq[r];
        }
        __syncthreads();
        float rstd[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = 16 * (r >> 1) + 8 * (r & 1) + g;
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sm[MlpTcSmem::red + 256 + w * 32 + row];
            rstd[r] = 1.0f / sqrtf(v * (1.0f / 128.0f) + kLnEps);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int r = 2 * mt + hrow, row = 16 * mt + 8 * hrow + g, col = 16 * warp + 8 * nt + 2 * t;
                    float2 y;
                    y.x = fmaxf((zacc[mt][nt][2 * hrow] - mean[r]) * rstd[r] * sm[MlpTcSmem::P2 + 128 + col] +
                                    sm[MlpTcSmem::P2 + 256 + col], 0.0f);
                    y.y = fmaxf((zacc[mt][nt][2 * hrow + 1] - mean[r]) * rstd[r] * sm[MlpTcSmem::P2 + 128 + col + 1] +
                                    sm[MlpTcSmem::P2 + 256 + col + 1], 0.0f);
                    *reinterpret_cast<float2*>(sm + MlpTcSmem::h2 + row * kH2Stride + col) = y;
                }
    }
    __syncthreads();
    // ---- heads: thread = (sample lane, output warp < 6), K = 128 ---------------------------------------------
    if (warp < 6) {
        // four interleaved partial sums (k mod 4): a 32-deep instead of a 128-deep chain of dependent FMAs on the
        // lockstep loop's critical path
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        const float* hp = sm + MlpTcSmem::h2 + lane * kH2Stride;
        const float* wp = sm + MlpTcSmem::Wh + warp;
#pragma unroll 8
        for (int k = 0; k < 128; k += 4) {
            const float4 h = *reinterpret_cast<const float4*>(hp + k);
            a0 = fmaf(h.x, wp[(k + 0) * 8], a0);
            a1 = fmaf(h.y, wp[(k + 1) * 8], a1);
            a2 = fmaf(h.z, wp[(k + 2) * 8], a2);
            a3 = fmaf(h.w, wp[(k + 3) * 8], a3);
        }
        sm[MlpTcSmem::out + lane * 8 + warp] = ((a0 + a1) + (a2 + a3)) + sm[MlpTcSmem::bh + warp];
    }
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
