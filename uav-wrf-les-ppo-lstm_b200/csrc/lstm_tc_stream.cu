// lstm_tc_stream.cu -- the deferred stop head for the LARGE hidden sizes of BASELINE configs[4] (H = 128, 256):
// tcgen05 gate GEMM with the gate weights STREAMED from L2, because they no longer fit in shared memory
// (H = 256: 4H x (H + 16) x (hi + lo) fp16 = 1.1 MB).
//
// One persistent CTA per SM, tile = 128 windows (TMEM lane = window), 16 compute warps (TMEM lane quarter x four column
// groups) + one warp whose lane 0 issues the MMAs + one warp whose lane 0 is the TMA producer.  Per cell step
//
//     gates[128][4H] = [h_{t-1} (H) | x_t | 1 | 0 x 14] (K = H + 16)  .  Wg[4H][K]^T          (kind::f16, hi/lo split)
//
// is computed in column chunks of 128 gate columns = 32 hidden units (columns interleaved 4 * unit + gate as in
// lstm_tc_kernels.cu, rows of Wg pre-scaled by -log2 e / -2 log2 e):
//   * A = the tile's h operand (hi + lo, 128 x K) stays RESIDENT in shared memory for the whole step (H = 256: 139 KB);
//   * B = the chunk's weights, pre-split and pre-arranged in operand layout by lstm_stream_prep_kernel, arrive as
//     cp.async.bulk (TMA) copies of 32 KB per (column chunk, K chunk of 64) through a two-stage ring
//     (mbarrier complete_tx / tcgen05.commit) filled by a producer thread of its own;
//   * two TMEM accumulators of 128 columns ping-pong: while the tensor core works on chunk c + 1, the compute warps run
//     the activations of chunk c (7 MUFU per unit: 1792 cycles per chunk, hidden behind the chunk's
//     3 x (K / 16) x 64 = 3264 MMA cycles at H = 256);
//   * the cell state lives in TMEM too (columns [256, 256 + H), tcgen05.ld / tcgen05.st);
//   * the new h of a chunk cannot overwrite A while later chunks of the same step still read the old h, and 128 KB of
//     pending h fit neither registers nor a second A buffer: it goes to a per-CTA scratch in global memory (L2 resident,
//     written in operand layout) and is copied back into A once the step's last MMA has completed.
// Bound: the L2 -> shared-memory weight stream.  Every SM pulls the whole 4H x K x 4 B of weights through its ring in
// every cell step: at the MMA-bound rate (8 chunks x 3264 cycles per step and tile at H = 256) that is 42 B per cycle
// and SM = 6.2 KB per cycle for the chip, the L2's whole delivery rate, so the tensor pipe runs at about a third of
// its rate (measured, 4096 x 256 windows: H = 256 37.4 ms = 2.8e7 env-steps/s with the rollout, H = 128 14.0 ms;
// round 1 on the CUDA cores: 2400 ms / 4.3e5).  What would lift it: a 2- or 4-CTA cluster that multicasts each
// chunk (one L2 read per cluster), or cta_group::2 pairs that hold half of B each.
// FLOP per window 2 * 4H * (1 + H) * W as in lstm_kernels.cu.
#include "lstm_tile.cuh"
#include "tc_gemm.cuh"

namespace plume {

constexpr int kLsComputeThreads = 512;
constexpr int kLsThreads = kLsComputeThreads + 64;     // + the MMA issuer warp + the TMA producer warp
constexpr int kLsFullSlots = 1024;          // 16-byte slots of a [128][64] fp16 chunk of the A operand (16 KB)
constexpr int kLsTailSlots = 256;           // of the [128][16] tail chunk (x, 1, zeros): 4 KB
// B travels in [128][kLsBK] chunks (hi + lo per ring stage, 64 KB of ring in all), filled by a producer thread of its
// own: the thread that waits for a stage to drain must not be the one that issues the MMAs (with one thread in both
// roles every chunk waited for the completion of the previous chunk's MMAs).  Measured at H = 256, 4096 x 256 windows
// incl. the rollout loop: one thread, 2 x 32 KB 42.9 ms; one thread, 4 x 16 KB 56.1 ms; own producer, 4 x 16 KB
// 45.2 ms; own producer, 2 x 32 KB 37.4 ms (fewer, larger chunks: fewer mbarrier round trips per byte)
#ifndef PLUME_LS_BK
#define PLUME_LS_BK 64
#endif
constexpr int kLsBK = PLUME_LS_BK;          // K width of a B chunk: 64 (two ring stages of 32 KB) or 32 (four of 16 KB)
constexpr int kLsBSlots = 16 * kLsBK;       // 16-byte slots of a [128][kLsBK] fp16 chunk of the B operand
constexpr int kLsStages = 128 / kLsBK;
constexpr uint32_t kLsStageBytes = 2 * kLsBSlots * 16;         // hi + lo of one B chunk

template <int H>
struct LsShape {
    static constexpr int KCF = H / 64;                   // full K chunks of A (h part)
    static constexpr int KB = H / kLsBK;                 // K chunks of B per column chunk (+ the tail)
    static constexpr int NC = 4 * H / 128;               // column chunks of 128 gate columns = 32 units
    static constexpr int a_slots = KCF * 2 * kLsFullSlots + 2 * kLsTailSlots;
    static constexpr int w_chunk_slots = KB * 2 * kLsBSlots + 2 * kLsTailSlots;       // weights of one column chunk
    static constexpr int scratch_slots = KCF * 2 * kLsFullSlots;                       // per CTA
};

template <int H>
struct LsSmem {                                          // offsets in 16-byte slots, then floats
    using S = LsShape<H>;
    static constexpr int a = 0;                                  // [KCF][hi | lo][1024], then tail [hi | lo][256]
    static constexpr int ring = a + S::a_slots;                  // [kLsStages][hi | lo][kLsBSlots]
    static constexpr int f_base = (ring + kLsStages * 2 * kLsBSlots) * 4;
    static constexpr int xs = f_base;                            // [20 steps][128] window values
    static constexpr int hd = xs + 20 * 128;                     // [2][H] head weights, [4]
    static constexpr int exch = hd + 2 * H + 4;                  // [4 column groups][128][2]
    static constexpr int total = exch + 4 * 128 * 2;
};
static_assert(LsSmem<256>::total * 4 + 256 <= 227 * 1024, "stream kernel: shared memory plan exceeds 227 KB");

__device__ __forceinline__ float ls_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ls_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void ls_compute_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void ls_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ls_tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ls_tmem_st8(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- weights -> operand chunks (once per launch) ---------------------------------------------------------------
// w_ops[cc][ K chunk kc < H / kLsBK: hi, lo of kLsBSlots slots each ; tail: hi 256, lo 256 ]; row n of chunk cc = gate
// g = n & 3 of unit 32 cc + (n >> 2); K = [w_hh row (H) | w_ih | b_ih + b_hh | 0 x 14]
template <int H>
__global__ void __launch_bounds__(256) lstm_stream_prep_kernel(LtArgs a, uint4* __restrict__ w_ops) {
    using S = LsShape<H>;
    constexpr float kL2e = 1.4426950408889634f;
    constexpr int units_per_row = H / 8 + 2;                     // 16-byte units along K
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S::NC * 128 * units_per_row) return;
    const int unit = i % units_per_row, rown = i / units_per_row;
    const int cc = rown >> 7, n = rown & 127;
    const int g = n & 3, src = g * H + 32 * cc + (n >> 2);
    const float scale = (g == 2) ? -2.0f * kL2e : -kL2e;
    float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
    if (unit < H / 8) {
        w0 = *reinterpret_cast<const float4*>(a.w_hh + (size_t)src * H + 8 * unit);
        w1 = *reinterpret_cast<const float4*>(a.w_hh + (size_t)src * H + 8 * unit + 4);
    } else if (unit == H / 8) {
        w0.x = a.w_ih[src];
        w0.y = a.b_ih[src] + a.b_hh[src];
    }
    w0.x *= scale; w0.y *= scale; w0.z *= scale; w0.w *= scale;
    w1.x *= scale; w1.y *= scale; w1.z *= scale; w1.w *= scale;
    uint4 hi, lo;
    tc::split_f16x8(w0, w1, 1.0f, hi, lo);
    uint4* chunk = w_ops + (size_t)cc * S::w_chunk_slots;
    if (unit < H / 8) {
        constexpr int upc = kLsBK / 8;                           // 16-byte units per chunk row
        const int kc = unit / upc, u = unit % upc;
        const int f = (n >> 3) * (upc * 8) + u * 8 + (n & 7);
        chunk[kc * 2 * kLsBSlots + f] = hi;
        chunk[kc * 2 * kLsBSlots + kLsBSlots + f] = lo;
    } else {
        const int u = unit - H / 8;
        const int f = (n >> 3) * 16 + u * 8 + (n & 7);
        chunk[S::KB * 2 * kLsBSlots + f] = hi;
        chunk[S::KB * 2 * kLsBSlots + kLsTailSlots + f] = lo;
    }
}

template <int H>
__global__ void __launch_bounds__(kLsThreads, 1) stop_head_stream_kernel(LtArgs a, const uint4* __restrict__ w_ops,
                                                                         uint4* __restrict__ scratch_all) {
    using S = LsShape<H>;
    using L = LsSmem<H>;
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t full[kLsStages], empty[kLsStages], acc_full[2], acc_free[2], a_ready;
    __shared__ uint32_t tmem_slot;
    // (the warp index and the TMEM base go through a shuffle: the compiler then knows them to be warp-uniform, and the MMA
    // warp -- whose control flow is run by all 32 lanes -- keeps its descriptors in uniform registers)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int wq = warp & 3, cg = (warp >> 2) & 3;    // TMEM lane quarter / column group (8 of a chunk's 32 units)
    const int row = wq * 32 + lane;
    const int N = a.n_envs, W = a.W;
    constexpr float kL2e = 1.4426950408889634f;
    uint4* const op = reinterpret_cast<uint4*>(sm);
    uint4* const A = op + L::a;
    uint4* const a_tail_hi = A + S::KCF * 2 * kLsFullSlots;
    uint4* const a_tail_lo = a_tail_hi + kLsTailSlots;
    uint4* const ring = op + L::ring;
    uint4* const scratch = scratch_all + (size_t)blockIdx.x * S::scratch_slots;

    if (tid == 0) {
        for (int q = 0; q < kLsStages; ++q) {
            tc::mbar_init(&full[q], 1);
            tc::mbar_init(&empty[q], 1);
        }
        tc::mbar_init(&acc_full[0], 1);
        tc::mbar_init(&acc_full[1], 1);
        tc::mbar_init(&acc_free[0], 16);
        tc::mbar_init(&acc_free[1], 16);
        tc::mbar_init(&a_ready, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    for (int i = tid; i < H; i += kLsThreads) {
        sm[L::hd + i] = a.w_peak[i];
        sm[L::hd + H + i] = a.w_stop[i];
    }
    if (tid == 0) {
        sm[L::hd + 2 * H] = a.b_peak[0];
        sm[L::hd + 2 * H + 1] = a.b_stop[0];
    }
    // second unit of the tail chunk (K columns H+8 .. H+15) stays zero for the whole kernel
    if (tid < 128) {
        a_tail_hi[(tid >> 3) * 16 + 8 + (tid & 7)] = make_uint4(0u, 0u, 0u, 0u);
        a_tail_lo[(tid >> 3) * 16 + 8 + (tid & 7)] = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t idesc = tc::make_idesc_f16(128, 128);
    float* const xs = sm + L::xs;
    float* const exch = sm + L::exch;

    const int env_tiles = (N + 127) / 128;
    const long long tiles = (long long)env_tiles * a.horizon;
    uint32_t item = 0;            // producer / issuer: B chunks streamed / consumed so far (each role counts its own)
    uint32_t chunk_q = 0;         // column chunks processed so far (both roles count them identically)
    uint32_t steps_done = 0;      // cell steps so far (phase of a_ready)
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t = (int)(tile / env_tiles), env0 = (int)(tile - (long long)t * env_tiles) * 128;
        __syncthreads();
        for (int i = tid; i < W * 128; i += kLsThreads) {          // xs[k][s], k = 0 oldest
            const int k = i >> 7, s = i & 127, env = env0 + s;
            const int tt = t - (W - 1) + k;
            float v = 0.0f;
            if (env < N) v = tt >= 0 ? a.conc_sample[(size_t)tt * N + env] : a.window_in[(size_t)env * W + (W + tt)];
            xs[k * 128 + s] = v;
        }
        int fill = 0;
        if (tid < 128 && env0 + tid < N) fill = a.fill_t[(size_t)t * N + env0 + tid];
        const bool full_w = fill >= W;                             // meaningful for tid < 128 (row = tid)
        const bool any = __syncthreads_or(full_w);                 // also publishes xs
        float pp = 0.0f, ps = 0.0f;
        if (any && warp == kLsComputeThreads / 32 + 1) {
            // ---- TMA producer (one thread): the weight chunks of the W steps, in the order the issuer consumes them -----
            if (lane == 0) {
                constexpr uint32_t per_step = (uint32_t)(S::NC * (S::KB + 1));
                for (uint32_t n = 0; n < (uint32_t)W * per_step; ++n, ++item) {
                    const uint32_t st = item % kLsStages, use = item / kLsStages;
                    if (use >= 1) tc::mbar_wait(&empty[st], (use - 1) & 1u);
                    const uint32_t within = n % per_step;
                    const uint32_t cc = within / (S::KB + 1), kc = within % (S::KB + 1);
                    const bool tail = kc == (uint32_t)S::KB;
                    const uint4* src = w_ops + (size_t)cc * S::w_chunk_slots + (size_t)kc * 2 * kLsBSlots;
                    const uint32_t bytes = tail ? 2u * kLsTailSlots * 16u : kLsStageBytes;
                    const uint32_t mb = tc::smem_u32(&full[st]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(tc::smem_u32(ring + st * 2 * kLsBSlots)), "l"(src), "r"(bytes), "r"(mb) : "memory");
                }
            }
        } else if (any && warp == kLsComputeThreads / 32) {
            // ---- MMA issuer: one warp with warp-uniform control flow, lane 0 issues ------------------------------------------
            // (inside a one-lane branch every tcgen05.mma was wrapped in ELECT + 7 x R2UR + BRA.U.ANY: ~100 cycles each)
            {
                constexpr uint32_t per_step = (uint32_t)(S::NC * (S::KB + 1));
                for (int step = 0; step < W; ++step) {
                    tc::mbar_wait_warp(&a_ready, steps_done & 1u, 20);      // A of this step is in shared memory
                    ++steps_done;
                    tc::tc_fence_after();
                    for (uint32_t w = 0; w < per_step; ++w, ++item) {
                        const uint32_t st = item % kLsStages;
                        const uint32_t cc = w / (S::KB + 1), kc = w % (S::KB + 1);
                        (void)cc;
                        const bool tail = kc == (uint32_t)S::KB;
                        const uint32_t b = chunk_q & 1u, use = chunk_q >> 1;
                        if (kc == 0 && use >= 1) {                 // the accumulator's previous contents have been read
                            tc::mbar_wait_warp(&acc_free[b], (use - 1) & 1u, 20);
                            tc::tc_fence_after();
                        }
                        tc::mbar_wait_warp(&full[st], (item / kLsStages) & 1u, 20);
                        tc::tc_fence_after();
                        // A: the K chunk of 64 that holds column kc * kLsBK, plus 256 bytes per K step of 16 inside it
                        const uint32_t a_sbo = tail ? 256u : 1024u, b_sbo = tail ? 256u : (uint32_t)(kLsBK / 8) * 128u;
                        const uint32_t k0 = kc * (uint32_t)kLsBK;
                        const uint32_t ah = tail ? tc::smem_u32(a_tail_hi)
                                                 : tc::smem_u32(A + (k0 >> 6) * 2 * kLsFullSlots) + ((k0 & 63u) >> 4) * 256u;
                        const uint32_t al = tail ? tc::smem_u32(a_tail_lo) : ah + kLsFullSlots * 16u;
                        const uint32_t bh = tc::smem_u32(ring + st * 2 * kLsBSlots);
                        const uint32_t bl = bh + (tail ? kLsTailSlots : kLsBSlots) * 16u;
                        const int ksteps = tail ? 1 : kLsBK / 16;
                        for (int j = 0; j < ksteps; ++j) {
                            const uint32_t off = j * 2 * tc::kLBO;
                            const uint64_t dah = tc::make_smem_desc(ah + off, tc::kLBO, a_sbo);
                            const uint64_t dal = tc::make_smem_desc(al + off, tc::kLBO, a_sbo);
                            const uint64_t dbh = tc::make_smem_desc(bh + off, tc::kLBO, b_sbo);
                            const uint64_t dbl = tc::make_smem_desc(bl + off, tc::kLBO, b_sbo);
                            const uint32_t d = tmem + 128u * b;
                            if (lane == 0) {
                                tc::mma_f16(d, dal, dbh, idesc, (kc == 0 && j == 0) ? 0u : 1u);
                                tc::mma_f16(d, dah, dbl, idesc, 1u);
                                tc::mma_f16(d, dah, dbh, idesc, 1u);
                            }
                        }
                        if (lane == 0) tc::mma_commit(&empty[st]);
                        if (tail) {
                            if (lane == 0) tc::mma_commit(&acc_full[b]);
                            ++chunk_q;
                        }
                    }
                }
            }
        } else if (any) {
            // ---- compute warps --------------------------------------------------------------------------------------
            // A of step 0: h = 0 (this thread's share of the resident operand), [x_0, 1] in the tail chunk
            for (int i = tid; i < S::KCF * 2 * kLsFullSlots; i += kLsComputeThreads) A[i] = make_uint4(0u, 0u, 0u, 0u);
            for (int step = 0; step < W; ++step) {
                if (tid < 128) {
                    uint32_t xh_, xl_;
                    tc::split_f16x2(xs[step * 128 + tid], 1.0f, 1.0f, xh_, xl_);
                    a_tail_hi[(tid >> 3) * 16 + (tid & 7)] = make_uint4(xh_, 0u, 0u, 0u);
                    a_tail_lo[(tid >> 3) * 16 + (tid & 7)] = make_uint4(xl_ & 0xFFFFu, 0u, 0u, 0u);
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                ls_compute_sync();
                if (tid == 0) ls_mbar_arrive(&a_ready);
                const bool last = step + 1 == W;
                for (int cc = 0; cc < S::NC; ++cc, ++chunk_q) {
                    const uint32_t b = chunk_q & 1u, use = chunk_q >> 1;
                    tc::mbar_wait(&acc_full[b], use & 1u);
                    tc::tc_fence_after();
                    float v[32], c8[8], hv[8];
                    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
                    tc::tmem_ld32(tmem + lane_addr + 128u * b + 32u * cg, v);
                    if (step > 0) ls_tmem_ld8(tmem + lane_addr + 256u + 32u * cc + 8u * cg, c8);
                    tc::tmem_ld_wait();
                    if (step == 0) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) c8[u] = 0.0f;
                    }
                    // the accumulator is in registers: hand the buffer back to the tensor core
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ls_mbar_arrive(&acc_free[b]);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float ei = ls_ex2(fminf(v[4 * u + 0], 40.0f));
                        const float ef = ls_ex2(fminf(v[4 * u + 1], 40.0f));
                        const float eg = ls_ex2(fminf(v[4 * u + 2], 40.0f));
                        const float eo = ls_ex2(fminf(v[4 * u + 3], 40.0f));
                        const float pi = 1.0f + ei, pf = 1.0f + ef, pg = 1.0f + eg;
                        const float pig = pi * pg;
                        const float r = ls_rcp(pig * pf);
                        c8[u] = fmaf(c8[u], r * pig, (1.0f - eg) * (r * pf));        // c = sigmoid(f) c + sigmoid(i) tanh(g)
                        const float ec = ls_ex2(fminf(c8[u] * (-2.0f * kL2e), 40.0f));
                        hv[u] = (1.0f - ec) * ls_rcp((1.0f + eo) * (1.0f + ec));      // sigmoid(o) tanh(c)
                    }
                    if (!last) {
                        ls_tmem_st8(tmem + lane_addr + 256u + 32u * cc + 8u * cg, c8);
                        uint4 hi, lo;
                        tc::split_f16x8(make_float4(hv[0], hv[1], hv[2], hv[3]), make_float4(hv[4], hv[5], hv[6], hv[7]),
                                        1.0f, hi, lo);
                        // units 32 cc + 8 cg ..: K chunk cc / 2, unit 4 (cc & 1) + cg of that chunk
                        const int f = (cc >> 1) * 2 * kLsFullSlots + (row >> 3) * 64 + (4 * (cc & 1) + cg) * 8 + (row & 7);
                        scratch[f] = hi;
                        scratch[f + kLsFullSlots] = lo;
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            pp = fmaf(hv[u], sm[L::hd + 32 * cc + 8 * cg + u], pp);
                            ps = fmaf(hv[u], sm[L::hd + H + 32 * cc + 8 * cg + u], ps);
                        }
                    }
                }
                if (!last) {
                    // every MMA of this step has completed (the last chunk's accumulator was committed after them):
                    // the new h moves from the scratch into the resident operand
                    ls_compute_sync();
                    for (int i = tid; i < S::scratch_slots; i += kLsComputeThreads) A[i] = scratch[i];
                }
            }
        }
        if (tid < kLsComputeThreads) {
            exch[(cg * 128 + row) * 2] = pp;
            exch[(cg * 128 + row) * 2 + 1] = ps;
        }
        __syncthreads();
        if (tid < 128 && env0 + tid < N) {
            const size_t i = (size_t)t * N + env0 + tid;
            float peak = 0.0f, stop_p = 0.0f;
            if (full_w) {
                float sp = 0.0f, ss = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    sp += exch[(q * 128 + tid) * 2];
                    ss += exch[(q * 128 + tid) * 2 + 1];
                }
                peak = sp + sm[L::hd + 2 * H];
                stop_p = sigmoidf_acc(ss + sm[L::hd + 2 * H + 1]);
            }
            if (a.stop_prob) a.stop_prob[i] = stop_p;
            if (a.stop_flag) a.stop_flag[i] = (full_w && stop_p > a.threshold) ? 1 : 0;   // evaluate_with_lstm.py:77
            if (a.peak_pred) a.peak_pred[i] = peak;
            if (a.trend) {
                float tr[4] = {0, 0, 0, 0};
                if (full_w && W >= 4)
                    trend_from_last4(100.0 * (double)xs[(W - 4) * 128 + tid], 100.0 * (double)xs[(W - 3) * 128 + tid],
                                     100.0 * (double)xs[(W - 2) * 128 + tid], 100.0 * (double)xs[(W - 1) * 128 + tid],
                                     a.src_dist[i], a.conc_peak, tr);
                *reinterpret_cast<float4*>(a.trend + i * 4) = make_float4(tr[0], tr[1], tr[2], tr[3]);
            }
            if (t == a.horizon - 1 && a.window_out)
                for (int k = 0; k < W; ++k) a.window_out[(size_t)(env0 + tid) * W + k] = xs[k * 128 + tid];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

static uint4* stream_buffer(size_t bytes, int which) {        // grow-only device buffers, stream-ordered reuse
    static void* buf[2] = {nullptr, nullptr};
    static size_t cap[2] = {0, 0};
    if (bytes > cap[which]) {
        if (buf[which]) cudaFree(buf[which]);
        if (cudaMalloc(&buf[which], bytes) != cudaSuccess) {
            buf[which] = nullptr;
            cap[which] = 0;
            return nullptr;
        }
        cap[which] = bytes;
    }
    return reinterpret_cast<uint4*>(buf[which]);
}

template <int H>
static int launch_ls(const LtArgs& a, cudaStream_t s) {
    using S = LsShape<H>;
    static bool configured = false;
    const int smem = LsSmem<H>::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(stop_head_stream_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return fail("stop-head stream kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const long long tiles = (long long)((a.n_envs + 127) / 128) * a.horizon;
    long long grid = sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    uint4* w_ops = stream_buffer((size_t)S::NC * S::w_chunk_slots * 16, 0);
    uint4* scratch = stream_buffer((size_t)grid * S::scratch_slots * 16, 1);
    if (!w_ops || !scratch) return fail("stop-head stream kernel: cannot allocate the operand buffers");
    const int prep_threads = S::NC * 128 * (H / 8 + 2);
    lstm_stream_prep_kernel<H><<<(prep_threads + 255) / 256, 256, 0, s>>>(a, w_ops);
    stop_head_stream_kernel<H><<<(int)grid, kLsThreads, smem, s>>>(a, w_ops, scratch);
    if (cudaGetLastError() != cudaSuccess) return fail("stop-head stream kernel launch failed");
    return 0;
}

bool stop_head_segment_stream_supports(int hidden) { return hidden == 128 || hidden == 256; }

int launch_stop_head_segment_stream(const LtArgs& a, int hidden, cudaStream_t s) {
    if (hidden == 128) return launch_ls<128>(a, s);
    if (hidden == 256) return launch_ls<256>(a, s);
    return fail("stop-head stream kernel: hidden must be 128 or 256 (got %d)", hidden);
}

}  // namespace plume
