// lstm_tc_kernels.cu -- the deferred V2.1 stop head (PPOV2.1/evaluate_with_lstm.py:11-27,73-80) with the
// batched gate GEMM on the sm_100a tensor cores.
//
// One CTA = one tile of 128 windows (128 consecutive envs at one step t), 256 threads (8 warps: TMEM lane quarter x
// column half), THREE CTAs per SM so that one tile's activation math overlaps the other tiles' MMAs (measured, whole
// PPO iteration: 256 threads x 3 CTAs 13.51 ms, 512 x 2 13.61, 256 x 2 13.64, 512 x 1 14.0).  Per cell step
//
//     gates[128 windows][128] = [h_{t-1} (32) | x_t | 1 | 0..] (K = 48)  .  Wg[128][48]^T
//
// is ONE tcgen05 GEMM (M = 128, N = 128, kind::f16, 3 K-steps x 3 MMAs for the two-term fp16 split x = hi + lo of
// tc_gemm.cuh with unscaled lo in one accumulator: every operand is O(1), so lo keeps an absolute precision of 2^-25),
// accumulator in TMEM: the input weight and both biases ride along as two extra K columns, and the rows of Wg are
// pre-scaled by -log2(e) (i, f, o) / -2 log2(e) (g) so that the epilogue starts directly with ex2.  Wg (hi + lo, 24 KB)
// stays resident in shared memory for all tiles; h is written back as the next step's A operand by the threads that
// computed it: TMEM lane = window row and gate columns are interleaved (column = 4 * unit + gate), so i, f, g, o of a
// hidden unit sit in one thread, and a thread's 16 units are exactly two 16-byte operand slots (8 fp16 each) of its row.
// The cell state lives in registers.
//
// Bound: the 7 MUFU ops per (window, unit, step) (5 ex2 + 2 rcp; 16 MUFU lanes/clk/SM), not the GEMM:
// 128 x 32 x 7 / 16 = 1792 cycles per cell step and tile against 9 MMAs x 64 = 576 tensor cycles (the first version,
// 3xTF32 with K = 40 and 8 warps per CTA: 15 MMAs = 960 cycles, 24 % of the stall samples waiting for them,
// profiles/r1r_stop_head_ncu_summary.txt).  FLOP per window as in lstm_kernels.cu (2*4H*(1+H)*W = 168 960 for
// H = 32, W = 20).
#include "lstm_tile.cuh"
#include "tc_gemm.cuh"

namespace plume {

#ifndef PLUME_LT_PARTS
#define PLUME_LT_PARTS 2
#endif
#ifndef PLUME_LT_CTAS
#define PLUME_LT_CTAS 3
#endif
constexpr int kLtParts = PLUME_LT_PARTS;         // column parts per window row: 4 (512 threads) or 2 (256 threads)
constexpr int kLtThreads = 128 * kLtParts;
constexpr int kLtUPT = 32 / kLtParts;            // hidden units per thread
constexpr int kLtCtas = PLUME_LT_CTAS;           // resident CTAs per SM
constexpr int kLtK = 48;                         // padded K (fp16): 32 hidden + x + 1 + 14 zeros = 3 MMA K-steps
constexpr int kLtUnits = kLtK / 8;               // 16-byte units (8 fp16) per operand row
constexpr uint32_t kLtSBO = kLtUnits * 128;      // bytes between 8-row groups
constexpr int kLtOperand = 128 * kLtUnits;       // 16-byte slots per [128][48] fp16 operand (12 KB)

struct LtSmem {                                  // offsets in 16-byte slots, then floats
    static constexpr int b_hi = 0;
    static constexpr int b_lo = b_hi + kLtOperand;
    static constexpr int a_hi = b_lo + kLtOperand;
    static constexpr int a_lo = a_hi + kLtOperand;
    static constexpr int f_base = (a_lo + kLtOperand) * 4;      // float index of what follows
    static constexpr int xs = f_base;                           // [32 steps][128] window values
    static constexpr int hd = xs + kLstmMaxSteps * 128;         // [2][32] head weights, [2] biases
    static constexpr int exch = hd + 2 * 32 + 4;                // [4 quarters][128][2]
    static constexpr int total = exch + 4 * 128 * 2;
};
static_assert(LtSmem::total * 4 * kLtCtas <= 224 * 1024, "the resident CTAs of the stop-head kernel must fit one SM");

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 16-byte slot of (row, k/8) in a [128][48] fp16 K-major no-swizzle operand
__device__ __forceinline__ int lt_slot(int row, int unit) { return (row >> 3) * (kLtUnits * 8) + unit * 8 + (row & 7); }

__global__ void __launch_bounds__(kLtThreads, kLtCtas) stop_head_segment_tc_kernel(LtArgs a) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wq = warp & 3, quarter = warp >> 2;    // TMEM lane quarter / quarter of the 32 hidden units
    const int row = wq * 32 + lane;                  // TMEM lane = window of the tile
    const int N = a.n_envs, W = a.W;
    constexpr float kL2e = 1.4426950408889634f;
    uint4* const op = reinterpret_cast<uint4*>(sm);
    uint4* const ah = op + LtSmem::a_hi;
    uint4* const al = op + LtSmem::a_lo;

    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<128>(&tmem_slot);
    // ---- resident B operand: row n = 4*j + g of Wg, columns [w_hh row (32) | w_ih | b_ih + b_hh | 0 x 14] ----
    for (int i = tid; i < 128 * kLtUnits; i += kLtThreads) {
        const int n = i / kLtUnits, unit = i - n * kLtUnits;
        const int j = n >> 2, g = n & 3, src = g * 32 + j;
        const float scale = (g == 2) ? -2.0f * kL2e : -kL2e;
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
        if (unit < 4) {
            w0 = *reinterpret_cast<const float4*>(a.w_hh + src * 32 + 8 * unit);
            w1 = *reinterpret_cast<const float4*>(a.w_hh + src * 32 + 8 * unit + 4);
        } else if (unit == 4) {
            w0.x = a.w_ih[src];
            w0.y = a.b_ih[src] + a.b_hh[src];
        }
        w0.x *= scale; w0.y *= scale; w0.z *= scale; w0.w *= scale;
        w1.x *= scale; w1.y *= scale; w1.z *= scale; w1.w *= scale;
        uint4 hi, lo;
        tc::split_f16x8(w0, w1, 1.0f, hi, lo);
        op[LtSmem::b_hi + lt_slot(n, unit)] = hi;
        op[LtSmem::b_lo + lt_slot(n, unit)] = lo;
    }
    // K columns 40..47 of A stay zero for the whole kernel
    if (quarter == kLtParts - 1) {
        ah[lt_slot(row, 5)] = make_uint4(0u, 0u, 0u, 0u);
        al[lt_slot(row, 5)] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < 32) {
        sm[LtSmem::hd + tid] = a.w_peak[tid];
        sm[LtSmem::hd + 32 + tid] = a.w_stop[tid];
    }
    if (tid == 0) {
        sm[LtSmem::hd + 64] = a.b_peak[0];
        sm[LtSmem::hd + 65] = a.b_stop[0];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc_f16(128, 128);
    float* const xs = sm + LtSmem::xs;
    uint32_t phase = 0;

    const int env_tiles = (N + 127) / 128;
    const long long tiles = (long long)env_tiles * a.horizon;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t = (int)(tile / env_tiles), env0 = (int)(tile - (long long)t * env_tiles) * 128;
        __syncthreads();
        for (int i = tid; i < W * 128; i += kLtThreads) {      // xs[k][s], k = 0 oldest
            const int k = i >> 7, s = i & 127, env = env0 + s;
            const int tt = t - (W - 1) + k;
            float v = 0.0f;
            if (env < N) v = tt >= 0 ? a.conc_sample[(size_t)tt * N + env] : a.window_in[(size_t)env * W + (W + tt)];
            xs[k * 128 + s] = v;
        }
        int fill = 0;
        if (quarter == 0 && env0 + row < N) fill = a.fill_t[(size_t)t * N + env0 + row];
        const bool full = fill >= W;
        const bool any = __syncthreads_or(full);               // also publishes xs
        float hreg[kLtUPT];
#pragma unroll
        for (int u = 0; u < kLtUPT; ++u) hreg[u] = 0.0f;
        if (any) {
            float cst[kLtUPT];
#pragma unroll
            for (int u = 0; u < kLtUPT; ++u) cst[u] = 0.0f;
            // A operand of step 0: h = 0 (8 units = one slot), [x_0, 1] written below
#pragma unroll
            for (int q = 0; q < kLtUPT / 8; ++q) {
                ah[lt_slot(row, quarter * (kLtUPT / 8) + q)] = make_uint4(0u, 0u, 0u, 0u);
                al[lt_slot(row, quarter * (kLtUPT / 8) + q)] = make_uint4(0u, 0u, 0u, 0u);
            }
            for (int step = 0; step < W; ++step) {
                if (quarter == 0) {                            // [x_step, 1, 0 x 6] in K columns 32..39
                    uint32_t xh_, xl_;
                    tc::split_f16x2(xs[step * 128 + row], 1.0f, 1.0f, xh_, xl_);
                    ah[lt_slot(row, 4)] = make_uint4(xh_, 0u, 0u, 0u);
                    al[lt_slot(row, 4)] = make_uint4(xl_ & 0xFFFFu, 0u, 0u, 0u);        // lo of the constant 1 is 0
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc::tc_fence_after();
                    const uint32_t sah = tc::smem_u32(ah), sal = tc::smem_u32(al);
                    const uint32_t sbh = tc::smem_u32(op + LtSmem::b_hi), sbl = tc::smem_u32(op + LtSmem::b_lo);
#pragma unroll
                    for (int j = 0; j < kLtK / 16; ++j) {
                        const uint32_t off = j * 2 * tc::kLBO;          // 16 fp16 = two 16-byte core-matrix columns
                        const uint64_t dah = tc::make_smem_desc(sah + off, tc::kLBO, kLtSBO);
                        const uint64_t dal = tc::make_smem_desc(sal + off, tc::kLBO, kLtSBO);
                        const uint64_t dbh = tc::make_smem_desc(sbh + off, tc::kLBO, kLtSBO);
                        const uint64_t dbl = tc::make_smem_desc(sbl + off, tc::kLBO, kLtSBO);
                        tc::mma_f16(tmem, dal, dbh, idesc, j == 0 ? 0u : 1u);
                        tc::mma_f16(tmem, dah, dbl, idesc, 1u);
                        tc::mma_f16(tmem, dah, dbh, idesc, 1u);
                    }
                    tc::mma_commit(&bar);
                }
                tc::mbar_wait(&bar, phase & 1u);
                ++phase;
                tc::tc_fence_after();
                float v[4 * kLtUPT];
#pragma unroll
                for (int q = 0; q < kLtUPT / 8; ++q)
                    tc::tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(4 * kLtUPT * quarter + 32 * q), v + 32 * q);
                tc::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < kLtUPT; ++u) {
                    // pre-activations arrive as -log2e * (i, f, o) and -2 log2e * g; clamping the exponent at
                    // 40 changes a sigmoid by < 1e-12 and keeps every product below 2^123
                    const float ei = ex2_approx(fminf(v[4 * u + 0], 40.0f));
                    const float ef = ex2_approx(fminf(v[4 * u + 1], 40.0f));
                    const float eg = ex2_approx(fminf(v[4 * u + 2], 40.0f));
                    const float eo = ex2_approx(fminf(v[4 * u + 3], 40.0f));
                    const float pi = 1.0f + ei, pf = 1.0f + ef, pg = 1.0f + eg;
                    const float pig = pi * pg;
                    const float r = rcp_approx(pig * pf);
                    // c = sigmoid(f) c + sigmoid(i) tanh(g)
                    cst[u] = fmaf(cst[u], r * pig, (1.0f - eg) * (r * pf));
                    const float ec = ex2_approx(fminf(cst[u] * (-2.0f * kL2e), 40.0f));
                    hreg[u] = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));      // sigmoid(o) tanh(c)
                }
                if (step + 1 < W) {                            // h is the next step's A operand: one slot per thread
#pragma unroll
                    for (int q = 0; q < kLtUPT / 8; ++q) {
                        uint4 hi, lo;
                        tc::split_f16x8(make_float4(hreg[8 * q], hreg[8 * q + 1], hreg[8 * q + 2], hreg[8 * q + 3]),
                                        make_float4(hreg[8 * q + 4], hreg[8 * q + 5], hreg[8 * q + 6], hreg[8 * q + 7]),
                                        1.0f, hi, lo);
                        ah[lt_slot(row, quarter * (kLtUPT / 8) + q)] = hi;
                        al[lt_slot(row, quarter * (kLtUPT / 8) + q)] = lo;
                    }
                }
            }
        }
        // ---- heads: each thread holds 8 of the 32 hidden units of its window ----------------------------
        float pp = 0.0f, ps = 0.0f;
#pragma unroll
        for (int u = 0; u < kLtUPT; ++u) {
            pp = fmaf(hreg[u], sm[LtSmem::hd + kLtUPT * quarter + u], pp);
            ps = fmaf(hreg[u], sm[LtSmem::hd + 32 + kLtUPT * quarter + u], ps);
        }
        sm[LtSmem::exch + (quarter * 128 + row) * 2] = pp;
        sm[LtSmem::exch + (quarter * 128 + row) * 2 + 1] = ps;
        __syncthreads();
        if (quarter == 0 && env0 + row < N) {
            const size_t i = (size_t)t * N + env0 + row;
            float peak = 0.0f, stop_p = 0.0f;
            if (full) {
                float sp = 0.0f, ss = 0.0f;
#pragma unroll
                for (int q = 0; q < kLtParts; ++q) {
                    sp += sm[LtSmem::exch + (q * 128 + row) * 2];
                    ss += sm[LtSmem::exch + (q * 128 + row) * 2 + 1];
                }
                peak = sp + sm[LtSmem::hd + 64];
                stop_p = sigmoidf_acc(ss + sm[LtSmem::hd + 65]);
            }
            if (a.stop_prob) a.stop_prob[i] = stop_p;
            if (a.stop_flag) a.stop_flag[i] = (full && stop_p > a.threshold) ? 1 : 0;   // evaluate_with_lstm.py:77
            if (a.peak_pred) a.peak_pred[i] = peak;
            if (a.trend) {
                float tr[4] = {0, 0, 0, 0};
                if (full && W >= 4)
                    trend_from_last4(100.0 * (double)xs[(W - 4) * 128 + row], 100.0 * (double)xs[(W - 3) * 128 + row],
                                     100.0 * (double)xs[(W - 2) * 128 + row], 100.0 * (double)xs[(W - 1) * 128 + row],
                                     a.src_dist[i], a.conc_peak, tr);
                *reinterpret_cast<float4*>(a.trend + i * 4) = make_float4(tr[0], tr[1], tr[2], tr[3]);
            }
            if (t == a.horizon - 1 && a.window_out)
                for (int k = 0; k < W; ++k) a.window_out[(size_t)(env0 + row) * W + k] = xs[k * 128 + row];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

int launch_stop_head_segment_tc(const LtArgs& a, cudaStream_t s) {
    static bool configured = false;
    const int smem = LtSmem::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(stop_head_segment_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return fail("stop-head tensor-core kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const long long tiles = (long long)((a.n_envs + 127) / 128) * a.horizon;
    long long grid = (long long)kLtCtas * sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    stop_head_segment_tc_kernel<<<(int)grid, kLtThreads, smem, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return fail("stop-head tensor-core kernel launch failed");
    return 0;
}

}  // namespace plume
