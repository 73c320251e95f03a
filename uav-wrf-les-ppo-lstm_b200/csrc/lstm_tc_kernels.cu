// lstm_tc_kernels.cu -- the deferred stop head (PPOV2.1/evaluate_with_lstm.py:11-27,73-80; hidden sizes of
// BASELINE configs[4]) with the batched gate GEMM on the sm_100a tensor cores.  H = 32 and H = 64 here: the gate
// weights stay resident in shared memory; larger hidden sizes and stacked layers stream them (lstm_tc_stream.cu).
//
// A tile = 128 windows (128 consecutive envs at one step t), evaluated by a GROUP of 256 threads (8 warps: TMEM lane
// quarter x two halves of the hidden units).  A CTA holds kTPC groups that share one copy of the gate weights and run
// independently of one another (named barriers, one mbarrier and one TMEM accumulator per group), so that one tile's
// activation math overlaps another tile's MMAs:
//     H = 32: 1 tile x 3 CTAs per SM (measured best in round 1) or 2 tiles x 2 CTAs;  H = 64: 2 tiles x 1 CTA.
// Per cell step
//
//     gates[128 windows][4H] = [h_{t-1} (H) | x_t | 1 | 0 x 14] (K = H + 16)  .  Wg[4H][K]^T
//
// is ONE tcgen05 GEMM (M = 128, N = 4H <= 256, kind::f16, K/16 K-steps x 3 MMAs for the two-term fp16 split
// x = hi + lo of tc_gemm.cuh with unscaled lo in one accumulator: every operand is O(1), so lo keeps an absolute
// precision of 2^-25), accumulator in TMEM: the input weight and both biases ride along as two extra K columns, and the
// rows of Wg are pre-scaled by -log2(e) (i, f, o) / -2 log2(e) (g) so that the epilogue starts directly with ex2.
// h is written back as the next step's A operand by the threads that computed it: TMEM lane = window row and gate
// columns are interleaved (column = 4 * unit + gate), so i, f, g, o of a hidden unit sit in one thread, and 8 units of a
// thread are exactly one 16-byte operand slot (8 fp16) of its row.  The cell state lives in registers.
//
// Bound: the 7 MUFU ops per (window, unit, step) (5 ex2 + 2 rcp; 16 MUFU lanes/clk/SM), not the GEMM:
// 128 x H x 7 / 16 cycles per cell step and tile (H = 32: 1792, H = 64: 3584) against (K/16) x 3 MMAs x N/2 tensor
// cycles (H = 32: 9 x 64 = 576, H = 64: 15 x 128 = 1920).  FLOP per window as in lstm_kernels.cu (2*4H*(1+H)*W).
#include "lstm_tile.cuh"
#include "tc_gemm.cuh"

namespace plume {

#ifndef PLUME_LT_TPC32
#define PLUME_LT_TPC32 1          // tiles per CTA for H = 32 (1: three CTAs per SM, 2: two CTAs per SM)
#endif
constexpr int kLtGroupThreads = 256;

template <int H>
struct LtShape {
    static constexpr int K = H + 16;                     // padded K (fp16): H hidden + x + 1 + 14 zeros
    static constexpr int units = K / 8;                  // 16-byte units (8 fp16) per operand row
    static constexpr uint32_t sbo = units * 128;         // bytes between 8-row groups
    static constexpr int N = 4 * H;                      // gate columns = MMA N = TMEM columns per tile
    static constexpr int a_slots = 128 * units;          // 16-byte slots per [128][K] operand
    static constexpr int b_slots = N * units;            // per [4H][K] operand
    static constexpr int upt = H / 2;                    // hidden units per thread
    static constexpr int chunks = upt / 8;               // 32-column TMEM slabs per thread and step
};

template <int H, int kTPC>
struct LtSmem {                                          // offsets in 16-byte slots, then floats
    using S = LtShape<H>;
    static constexpr int b_hi = 0;
    static constexpr int b_lo = b_hi + S::b_slots;
    static constexpr int a0 = b_lo + S::b_slots;         // per tile group: a_hi, a_lo
    static constexpr int f_base = (a0 + kTPC * 2 * S::a_slots) * 4;        // float index of what follows
    static constexpr int hd = f_base;                                       // [2][H] head weights, [2] biases
    static constexpr int xs = hd + 2 * H + 4;                               // per group [32 steps][128] window values
    static constexpr int exch = xs + kTPC * kLstmMaxSteps * 128;            // per group [2 parts][128][2]
    static constexpr int total = exch + kTPC * 2 * 128 * 2;
};

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
// barriers of one 256-thread tile group (ids 1 .. kTPC; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool group_or(int id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 q, %2, 0;\n\t"
        "bar.red.or.pred p, %1, 256, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(id), "r"((uint32_t)pred)
        : "memory");
    return r != 0;
}

template <int H, int kTPC, int kCtas>
__global__ void __launch_bounds__(kLtGroupThreads * kTPC, kCtas) stop_head_segment_tc_kernel(LtArgs a) {
    using S = LtShape<H>;
    using L = LtSmem<H, kTPC>;
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar[kTPC];
    __shared__ uint32_t tmem_slot;
    // (grp, gwarp and the TMEM base go through a shuffle: the compiler then knows them to be warp-uniform and keeps the MMA
    // descriptors derived from them in uniform registers)
    const int tid = threadIdx.x, grp = __shfl_sync(0xffffffffu, tid / kLtGroupThreads, 0), gt = tid - grp * kLtGroupThreads;
    const int warp = gt >> 5, lane = gt & 31;
    const int gwarp = __shfl_sync(0xffffffffu, warp, 0);     // the same value, known to the compiler as warp-uniform
    const int wq = warp & 3, part = warp >> 2;        // TMEM lane quarter / half of the hidden units
    const int row = wq * 32 + lane;                   // TMEM lane = window of the tile
    const int N = a.n_envs, W = a.W;
    constexpr float kL2e = 1.4426950408889634f;
    uint4* const op = reinterpret_cast<uint4*>(sm);
    uint4* const ah = op + L::a0 + grp * 2 * S::a_slots;
    uint4* const al = ah + S::a_slots;
    // 16-byte slot of (row, k/8) in a K-major no-swizzle fp16 operand with S::units units per row
    auto slot = [](int r, int unit) { return (r >> 3) * (S::units * 8) + unit * 8 + (r & 7); };

    if (tid < kTPC) tc::mbar_init(&bar[tid], 1);
    if (tid == 0) tc::mbar_fence_init();
    if (tid < 32) tc::tmem_alloc<(uint32_t)(kTPC * S::N)>(&tmem_slot);
    // ---- resident B operand: row n = 4*j + g of Wg, columns [w_hh row (H) | w_ih | b_ih + b_hh | 0 x 14] ----
    for (int i = tid; i < S::b_slots; i += blockDim.x) {
        const int n = i / S::units, unit = i - n * S::units;
        const int j = n >> 2, g = n & 3, src = g * H + j;
        const float scale = (g == 2) ? -2.0f * kL2e : -kL2e;
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
        if (unit < H / 8) {
            w0 = *reinterpret_cast<const float4*>(a.w_hh + src * H + 8 * unit);
            w1 = *reinterpret_cast<const float4*>(a.w_hh + src * H + 8 * unit + 4);
        } else if (unit == H / 8) {
            w0.x = a.w_ih[src];
            w0.y = a.b_ih[src] + a.b_hh[src];
        }
        w0.x *= scale; w0.y *= scale; w0.z *= scale; w0.w *= scale;
        w1.x *= scale; w1.y *= scale; w1.z *= scale; w1.w *= scale;
        uint4 hi, lo;
        tc::split_f16x8(w0, w1, 1.0f, hi, lo);
        op[L::b_hi + slot(n, unit)] = hi;
        op[L::b_lo + slot(n, unit)] = lo;
    }
    // K columns H+8 .. H+15 of A stay zero for the whole kernel
    if (part == 1) {
        ah[slot(row, H / 8 + 1)] = make_uint4(0u, 0u, 0u, 0u);
        al[slot(row, H / 8 + 1)] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < H; i += blockDim.x) {
        sm[L::hd + i] = a.w_peak[i];
        sm[L::hd + H + i] = a.w_stop[i];
    }
    if (tid == 0) {
        sm[L::hd + 2 * H] = a.b_peak[0];
        sm[L::hd + 2 * H + 1] = a.b_stop[0];
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0) + (uint32_t)(grp * S::N);
    const uint32_t idesc = tc::make_idesc_f16(128, S::N);
    float* const xs = sm + L::xs + grp * kLstmMaxSteps * 128;
    float* const exch = sm + L::exch + grp * 2 * 128 * 2;
    const int bid = 1 + grp;
    uint32_t phase = 0;

    const int env_tiles = (N + 127) / 128;
    const long long tiles = (long long)env_tiles * a.horizon;
    for (long long tile = (long long)blockIdx.x * kTPC + grp; tile < tiles; tile += (long long)gridDim.x * kTPC) {
        const int t = (int)(tile / env_tiles), env0 = (int)(tile - (long long)t * env_tiles) * 128;
        group_sync(bid);
        for (int i = gt; i < W * 128; i += kLtGroupThreads) {      // xs[k][s], k = 0 oldest
            const int k = i >> 7, s = i & 127, env = env0 + s;
            const int tt = t - (W - 1) + k;
            float v = 0.0f;
            if (env < N) v = tt >= 0 ? a.conc_sample[(size_t)tt * N + env] : a.window_in[(size_t)env * W + (W + tt)];
            xs[k * 128 + s] = v;
        }
        int fill = 0;
        if (part == 0 && env0 + row < N) fill = a.fill_t[(size_t)t * N + env0 + row];
        const bool full = fill >= W;
        const bool any = group_or(bid, full);                  // also publishes xs
        float pp = 0.0f, ps = 0.0f;                            // this thread's part of the two head dot products
        if (any) {
            float cst[S::upt];
#pragma unroll
            for (int u = 0; u < S::upt; ++u) cst[u] = 0.0f;
            // A operand of step 0: h = 0, [x_0, 1] written below
#pragma unroll
            for (int q = 0; q < S::chunks; ++q) {
                ah[slot(row, part * S::chunks + q)] = make_uint4(0u, 0u, 0u, 0u);
                al[slot(row, part * S::chunks + q)] = make_uint4(0u, 0u, 0u, 0u);
            }
            for (int step = 0; step < W; ++step) {
                if (part == 0) {                               // [x_step, 1, 0 x 6] in K columns H .. H+7
                    uint32_t xh_, xl_;
                    tc::split_f16x2(xs[step * 128 + row], 1.0f, 1.0f, xh_, xl_);
                    ah[slot(row, H / 8)] = make_uint4(xh_, 0u, 0u, 0u);
                    al[slot(row, H / 8)] = make_uint4(xl_ & 0xFFFFu, 0u, 0u, 0u);        // lo of the constant 1 is 0
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                group_sync(bid);
                // The group's first warp issues, with WARP-UNIFORM control flow: inside a one-thread branch the compiler keeps
                // the descriptors in vector registers and wraps every tcgen05.mma in ELECT + 7 x R2UR + BRA.U.ANY (~100 cycles
                // of issue latency per instruction, on the critical path of every cell step); here they live in uniform
                // registers and only the instructions themselves are predicated on lane 0.
                if (gwarp == 0) {
                    tc::tc_fence_after();
                    const uint32_t sah = tc::smem_u32(ah), sal = tc::smem_u32(al);
                    const uint32_t sbh = tc::smem_u32(op + L::b_hi), sbl = tc::smem_u32(op + L::b_lo);
#pragma unroll
                    for (int j = 0; j < S::K / 16; ++j) {
                        const uint32_t off = j * 2 * tc::kLBO;          // 16 fp16 = two 16-byte core-matrix columns
                        const uint64_t dah = tc::make_smem_desc(sah + off, tc::kLBO, S::sbo);
                        const uint64_t dal = tc::make_smem_desc(sal + off, tc::kLBO, S::sbo);
                        const uint64_t dbh = tc::make_smem_desc(sbh + off, tc::kLBO, S::sbo);
                        const uint64_t dbl = tc::make_smem_desc(sbl + off, tc::kLBO, S::sbo);
                        if (lane == 0) {
                            tc::mma_f16(tmem, dal, dbh, idesc, j == 0 ? 0u : 1u);
                            tc::mma_f16(tmem, dah, dbl, idesc, 1u);
                            tc::mma_f16(tmem, dah, dbh, idesc, 1u);
                        }
                    }
                    if (lane == 0) tc::mma_commit(&bar[grp]);
                }
                tc::mbar_wait(&bar[grp], phase & 1u);
                ++phase;
                tc::tc_fence_after();
                const bool last = step + 1 == W;
#pragma unroll
                for (int q = 0; q < S::chunks; ++q) {
                    float v[32], hv[8];
                    tc::tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(4 * S::upt * part + 32 * q), v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
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
                        float& c = cst[8 * q + u];
                        c = fmaf(c, r * pig, (1.0f - eg) * (r * pf));
                        const float ec = ex2_approx(fminf(c * (-2.0f * kL2e), 40.0f));
                        hv[u] = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));      // sigmoid(o) tanh(c)
                    }
                    if (!last) {                               // h is the next step's A operand: one slot per chunk
                        uint4 hi, lo;
                        tc::split_f16x8(make_float4(hv[0], hv[1], hv[2], hv[3]), make_float4(hv[4], hv[5], hv[6], hv[7]),
                                        1.0f, hi, lo);
                        ah[slot(row, part * S::chunks + q)] = hi;
                        al[slot(row, part * S::chunks + q)] = lo;
                    } else {                                   // heads: fc_peak / fc_stop on h_W
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            pp = fmaf(hv[u], sm[L::hd + S::upt * part + 8 * q + u], pp);
                            ps = fmaf(hv[u], sm[L::hd + H + S::upt * part + 8 * q + u], ps);
                        }
                    }
                }
            }
        }
        exch[(part * 128 + row) * 2] = pp;
        exch[(part * 128 + row) * 2 + 1] = ps;
        group_sync(bid);
        if (part == 0 && env0 + row < N) {
            const size_t i = (size_t)t * N + env0 + row;
            float peak = 0.0f, stop_p = 0.0f;
            if (full) {
                const float sp = exch[row * 2] + exch[(128 + row) * 2];
                const float ss = exch[row * 2 + 1] + exch[(128 + row) * 2 + 1];
                peak = sp + sm[L::hd + 2 * H];
                stop_p = sigmoidf_acc(ss + sm[L::hd + 2 * H + 1]);
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
    if (tid < 32) tc::tmem_dealloc<(uint32_t)(kTPC * S::N)>(tmem_slot);
}

template <int H, int kTPC, int kCtas>
static int launch_lt(const LtArgs& a, cudaStream_t s) {
    static bool configured = false;
    const int smem = LtSmem<H, kTPC>::total * (int)sizeof(float);
    static_assert(LtSmem<H, kTPC>::total * 4 * kCtas + 1024 * kCtas <= 228 * 1024,
                  "the resident CTAs of the stop-head kernel must fit one SM");
    static_assert(kTPC * LtShape<H>::N * kCtas <= 512, "TMEM: 512 columns per SM");
    auto kernel = stop_head_segment_tc_kernel<H, kTPC, kCtas>;
    if (!configured) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("stop-head tensor-core kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const long long tiles = (long long)((a.n_envs + 127) / 128) * a.horizon;
    long long grid = (long long)kCtas * sm_count();
    if (grid <= 0) return fail("no CUDA device");
    const long long want = (tiles + kTPC - 1) / kTPC;
    if (want < grid) grid = want;
    kernel<<<(int)grid, kLtGroupThreads * kTPC, smem, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return fail("stop-head tensor-core kernel launch failed");
    return 0;
}

bool stop_head_segment_tc_supports(int hidden) { return hidden == 32 || hidden == 64; }

int launch_stop_head_segment_tc(const LtArgs& a, int hidden, cudaStream_t s) {
    if (hidden == 32) {
#if PLUME_LT_TPC32 == 1
        return launch_lt<32, 1, 3>(a, s);
#else
        return launch_lt<32, 2, 2>(a, s);
#endif
    }
    if (hidden == 64) return launch_lt<64, 2, 1>(a, s);
    return fail("stop-head tensor-core kernel: hidden %d has no resident-weight kernel", hidden);
}

}  // namespace plume
