// ppo_tc_kernels.cu -- K6 on the sm_100a tensor cores: the PPO minibatch gradient (train_ppo2.0.py:42-85).  Every
// GEMM-shaped part of the actor-critic (model.py:16-46) runs on tcgen05.mma (kind::f16 with the two-term fp16 split
// x = hi + lo of tc_gemm.cuh: three MMAs per K-step, fp32-grade products; accumulators in TMEM); LayerNorms, ReLUs, heads,
// loss and the per-parameter reductions run on the CUDA cores of the same persistent CTA (packed fp32 pairs FFMA2 / FMUL2 /
// FADD2 wherever register pairs form).  One CTA per SM, 128-sample tiles, 16 compute warps + 2 MMA warps.  Per-launch
// operands (16 W2 pre-split for G1 and for G2, the layer-1 operands, the LayerNorm-1 quadratic form) come from
// ppo_tc_prep_kernel.  DESIGN.md section 5 has the measurements behind every choice below.
//
//   L1  z1c[s][i]  = sum_k  x[s][k] W1c[i][k] + b1c[i]   K = 16 packed as [x_hi | x_lo] . [w_hi | w_hi] + [x_hi | x_lo] . [w_lo | 0];
//                                                         per chunk of 64 inputs into one of two TMEM buffers
//   G1  z2[s][o]   = sum_i  h1[s][i] W2[o][i]             A = h1 chunk = TMEM load of L1 -> LayerNorm scale -> ReLU -> split,
//                                                         B = 16 W2 chunk, bulk copy from L2
//   G2  dh1[s][i]  = sum_o  dz2[s][o] W2[o][i]            A = the resident dz2 operand read MN-major, B = 16 W2^T chunks (bulk)
//   G3  dW2[o][i] += sum_s  dz2[s][o] h1[s][i]            A = the same dz2 bytes read K-major, B = the STASHED h1: every G1 chunk
//                                                         is copied by one TMA tensor store into an L2 scratch where a pair of
//                                                         chunks is interleaved, and comes back as one MN-major operand, N = 128;
//                                                         accumulates in TMEM across ALL tiles of the CTA
//   P   P[i][c]    = sum_s  dy1[s][i] X[s][c]             the LayerNorm-1 backward column sums, M = 64 inputs x N = 8 per chunk,
//                                                         A = dy1 written MN-major by Ph6, X = {rstd x_0..x_5, rstd, 1}
//
// Nothing is computed twice on the CUDA cores except layer 1 in the LayerNorm-1 backward (ReLU mask and xhat1; TMEM is full
// by then), and every operand that already exists as bytes is moved by the copy engines.  MMA warp 1 (lane 0 issues) owns
// all copies and L1 / G1 / G2 / G3, MMA warp 2 the column-sum GEMMs; both run WARP-UNIFORM control flow (all 32 lanes wait
// and compute descriptors, which therefore live in uniform registers): inside a one-lane branch the compiler wrapped every
// tcgen05.mma in ELECT + 7 x R2UR + BRA.U.ANY, 105-140 cycles per instruction against 64 back to back.  mbarriers carry
// every hand-over; the compute warps meet at CTA barriers only inside Ph3 / Ph4 and at the end of a tile.
//
// TMEM map (512 columns): [0,128) G1 accumulator, then dh1 of inputs 0..127; [128,256) the two layer-1 buffers during the
// G1 pass, then dh1 of inputs 128..255; the column-sum results overwrite the first 8 (consumed) dh1 columns of their chunk;
// [256,512) the persistent dW2 accumulator (row = output o, column = input i).  Main and cross terms share one accumulator
// everywhere: operands are brought to O(1) (16 W2, dz_scale dz2) instead of scaling lo, undone where results are read.
//
// Layer 1 never materialises dz1: with dy1 = relu'(y1) o dh1 and the per-sample LayerNorm scalars
// m1 = mean_i(g1 dy1), m2 = mean_i(g1 dy1 xhat1), every layer-1 gradient is a linear function of the column sums P
// and of 35 per-sample scalar sums (S0, S[6], R[7], Q[6][6] symmetric); see the epilogue at the end
// of the kernel.  z1 - mean(z1) is evaluated directly from centred weights (W1 - column mean), so
// no mean pass is needed.  The algebra was checked against autograd in float64 and its float32
// error matches autograd's own.
//
// Algorithmic bytes per sample: one 48-byte record (obs 24, adv 4, ret 4, old value 4, old logp 4, action 4, pad 4;
// plume_ppo_pack) + an 8-byte permutation index; without the records 44 B from six arrays.  The operand chunks and the
// activation stash (2 KB per sample) stay in L2.  FLOP per sample: 3 x 65 536 + layer 1 on the tensor cores (x3 MMAs for
// the split), ~10 000 on the CUDA cores.
//
// Build-time knobs: PLUME_U4 unroll factor of Ph4's column loop, PLUME_TC_MAXNREG (register-sensitivity experiment),
// PLUME_TC_POLL_NS / PLUME_TC_WAIT_NS (back-off of the MMA warps / compute warps between barrier polls: no effect between
// 20 and 200 ns), PLUME_TC_PM / PLUME_TC_PN (shape of the column-sum MMA), PLUME_TC_TIMELINE (clock64 stamps per phase, per
// G1 chunk and per MMA-warp turn, printed by CTA 0), PLUME_TC_STORE_PROBE (duration of a stash store).
#include <cuda.h>            // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)

#include "ppo_loss.cuh"
#include "tc_gemm.cuh"

#define PLUME_STR(x) #x
#define PLUME_UNROLL(n) _Pragma(PLUME_STR(unroll n))
#ifndef PLUME_U4
#define PLUME_U4 16
#endif
// shape of the column-sum GEMM's instruction: M = 64 / N = 8 reads only the 64 real rows of its A operand; M = 128 / N = 16
// (rows 64..127 and columns 8..15 garbage that nobody reads) is the form whose accumulator layout needs no explanation
#ifndef PLUME_TC_PM
#define PLUME_TC_PM 64
#endif
#ifndef PLUME_TC_PN
#define PLUME_TC_PN 8
#endif
// back-off of the compute warps between polls of an mbarrier they wait on for long (measured: no effect between 20 and 200 ns)
#ifndef PLUME_TC_WAIT_NS
#define PLUME_TC_WAIT_NS 64
#endif
namespace plume {

constexpr int kTcTile = 128;
constexpr int kTcGroups = 4;                       // 128-thread groups of compute threads
constexpr int kTcThreads = 128 * kTcGroups;        // compute threads (producers + epilogues): 16 warps
constexpr int kTcLaunchThreads = kTcThreads + 64;  // + two warps whose lane 0 only issues MMAs (G1/G2/G3; column-sum GEMM)
constexpr int kXhStride = 132;             // [128][132]: conflict-free float4 rows and columns
constexpr int kStageStride = 132;          // [128][132]: transposed dy1 staging (conflict-free scalar stores along the
                                           // samples, float4 loads along the samples of one input)
constexpr int kChunkFloats = 128 * tc::kChunkK;

struct TcSmem {
    static constexpr int ring = 0;                                   // [2 stages][A_hi, A_lo, B_hi, B_lo][4096]
    static constexpr int xh = ring + 2 * 4 * kChunkFloats;           // [128][132] xhat2 -> dz2 -> dy1 staging
    static constexpr int W1c = xh + kTcTile * kXhStride;             // [6][256] centred feature.0.weight, k-major
    static constexpr int P1 = W1c + 6 * 256;                         // [3][256] centred bias, LN1 gamma, beta
    static constexpr int P2 = P1 + 3 * 256;                          // [3][128] feature.3.bias, LN2 gamma, beta
    static constexpr int Wh = P2 + 3 * 128;                          // [128][8] heads (6), LN2 gamma, beta
    static constexpr int bh = Wh + 128 * 8;                          // [8]
    static constexpr int x = bh + 8;                                 // [128][8] x0..x5, rstd1, 0
    static constexpr int dout = x + kTcTile * 8;                     // [128][8] d loss / d (logits, value), pol, val
    static constexpr int sc = dout + kTcTile * 8;                    // [128][4] rstd2, m1, m2, entropy term
    static constexpr int pf = sc + kTcTile * 4;                      // [128][12] next tile's gathered sample
    static constexpr int total = pf + kTcTile * 12;
    // [groups][128][8] exchange of partial sums between the column groups of one sample row: aliases the last operand
    // buffer of stage 0, which is idle whenever it is used (Ph3: G1 is complete and Ph4 has not yet written dz2 there;
    // end of Ph6: every G2 / G3 MMA of the tile has completed).  Stage 1 is not available: the MMA thread prefetches the
    // first W2^T chunks into it while Ph3 runs.
    static constexpr int exch = ring + 3 * kChunkFloats;
};
static_assert(kTcGroups * kTcTile * 8 <= kChunkFloats, "exchange area must fit one operand buffer");
static_assert(TcSmem::total * 4 + 64 <= 227 * 1024, "ppo_tc_kernel: shared memory plan exceeds 227 KB");
static_assert(kTcTile * kStageStride <= kTcTile * kXhStride, "dy1 staging must fit in the xhat region");

// layout of the pre-split weight workspace (float units; fp16 hi / lo operand chunks of 128 rows x 64 K = 16 KB each)
constexpr int kW2SplitG1 = 0;                         // [4 chunks][hi, lo][128 out][64 in] (K = in), 16 W2, lo unscaled; hi and
                                                      // lo of a chunk are contiguous: one 32 KB bulk copy
constexpr int kW2SplitG2 = 32768;                     // [2 halves][2 chunks][hi, lo][128 in][64 out] (K = out), 16 W2^T, lo
                                                      // unscaled; hi and lo of a chunk are contiguous: one 32 KB bulk copy
constexpr int kW2SplitFloats = 65536;
// per-launch invariants of layer 1, evaluated once by the last block of the prep kernel instead of by every CTA:
constexpr int kWsW1c = kW2SplitFloats;                // [6][256] feature.0.weight minus its column means, k-major
constexpr int kWsB1c = kWsW1c + 6 * 256;              // [256]    feature.0.bias minus its mean
constexpr int kWsLn1q = kWsB1c + 256;                 // [28] doubles: the LayerNorm-1 quadratic form (see ln1q below)
// B operands of the layer-1 GEMM (below): per chunk of 64 inputs [B1 2 KB][B2 2 KB], row n = input, two 16-byte K slots
// per row at (n >> 3) * 256 + slot * 128 + (n & 7) * 16 bytes: B1 = [w_hi | w_hi], B2 = [w_lo | 0] with w = (centred
// feature.0.weight row, centred bias, 0)
constexpr int kWsL1Ops = (kWsLn1q + 2 * 28 + 63) & ~63;
constexpr int kL1OpsFloats = 4 * 1024;                                  // 16 KB
constexpr int kWsStash = kWsL1Ops + kL1OpsFloats;          // per-CTA activation stash (below), 256-byte aligned
// The forward activations h1 of a tile are produced ONCE, as the A operand of G1 (four 32 KB chunks: fp16 hi + lo of
// 128 samples x 64 inputs).  Each chunk is copied to this L2-resident scratch by a bulk copy (cp.async.bulk, issued by
// the MMA thread) as soon as its MMAs are issued, and comes back by a bulk copy as the B operand of G3 (read MN-major:
// the same bytes) -- the CUDA cores no longer recompute h1^T for the weight-gradient GEMM.
constexpr int kStashChunkBytes = 2 * kChunkFloats * 4;                  // hi + lo of one chunk (32 KB)
constexpr int kStashFloatsPerCta = 4 * kStashChunkBytes / 4;            // 128 KB per CTA
constexpr int kTcMaxCtas = 192;                                         // >= the SM count (B200: 148)
constexpr long long kWsFloats = (long long)kWsStash + (long long)kTcMaxCtas * kStashFloatsPerCta;
static_assert((kWsLn1q % 2) == 0, "the float64 coefficients must be 8-byte aligned");
// The backward GEMMs keep main and cross terms in ONE accumulator (TMEM is full), so their lo parts are not scaled;
// instead the operands are brought to O(1): W2^T is multiplied by 16 (|w| ~ 0.1) and dz2 by a power of two near the
// global batch size (dz2 ~ 1/batch), passed to the kernel as dz_scale.  Both are undone where the results are read.
constexpr float kW2BwdScale = 16.0f;

// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE-rn operations per issue slot; the kernel is bound by
// instruction issue, not by the FMA pipe).  Each lane of a pair rounds exactly like the scalar instruction.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// barrier of the 256 compute threads (the issuer warp never takes part)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTcThreads) : "memory"); }
// barrier of the four warps that share a sample row (the TMEM lane quarter wq x the four column groups): the exchanges
// of per-sample partial sums in Ph3 / Ph6 only involve them, so the quarters need not wait for one another there
#ifndef PLUME_TC_NO_QBAR
__device__ __forceinline__ void quarter_sync(int wq) { asm volatile("bar.sync %0, 128;" ::"r"(2 + wq) : "memory"); }
#else
__device__ __forceinline__ void quarter_sync(int) { compute_sync(); }
#endif
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// ---- W2 -> fp16 hi/lo operand chunks in the canonical K-major layout (once per minibatch) -------------
// One thread per 16-byte operand slot (8 consecutive K values of one row).
constexpr int kPrepSplitBlocks = 32;                  // 32 x 256 threads = 8192 operand slots; block 32 = layer-1 block
__global__ void __launch_bounds__(256) ppo_tc_prep_kernel(const float* __restrict__ params, float* __restrict__ w2s,
                                                          float* __restrict__ zero) {
    // the gradient accumulator of the launch that follows (plume_ppo_update: no separate memset on the stream)
    if (zero != nullptr && blockIdx.x < kPrepSplitBlocks)
        for (int i = blockIdx.x * 256 + threadIdx.x; i < PLUME_MLP_PARAMS; i += kPrepSplitBlocks * 256) zero[i] = 0.0f;
    if (blockIdx.x == kPrepSplitBlocks) {
        // ---- layer-1 invariants: thread = output o --------------------------------------------------------------
        // z_o - mean_o(z) = (b_o - mean b) + sum_k x_k (w_ok - mean_o w_ok): centred weights make the LayerNorm-1 mean
        // pass unnecessary, and sum_o z_o^2 = q0 + sum_k q1[k] x_k + sum_{k<=l} q2[kl] x_k x_l (28 float64 coefficients)
        // replaces the variance pass.  Fixed reduction order (shuffle tree, then warps 0..7): deterministic.
        __shared__ float s_part[8][7];
        __shared__ float s_mean[7];
        __shared__ double s_q[8][28];
        const int o = threadIdx.x, warp = o >> 5, lane = o & 31;
        float w[7];
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = params[PLUME_OFF_W1 + o * 6 + k];
        w[6] = params[PLUME_OFF_B1 + o];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            float t = w[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) s_part[warp][k] = t;
        }
        __syncthreads();
        if (o < 7) {
            float t = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += s_part[q][o];
            s_mean[o] = t * (1.0f / 256.0f);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 7; ++k) w[k] -= s_mean[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) w2s[kWsW1c + k * 256 + o] = w[k];
        w2s[kWsB1c + o] = w[6];
        {   // B operands of the layer-1 GEMM: z1c[s][o] = sum_k x_k w_ok + b_o as [x_hi | x_lo] . [w_hi | w_hi] + [x_hi | x_lo] . [w_lo | 0]
            uint4 hi, lo;
            tc::split_f16x8(make_float4(w[0], w[1], w[2], w[3]), make_float4(w[4], w[5], w[6], 0.0f), 1.0f, hi, lo);
            uint4* ops = reinterpret_cast<uint4*>(w2s + kWsL1Ops) + (o >> 6) * 256;      // 4 KB per chunk of 64 inputs
            const int n = o & 63, f = (n >> 3) * 16 + (n & 7);
            ops[f] = hi;
            ops[f + 8] = hi;
            ops[128 + f] = lo;
            ops[128 + f + 8] = make_uint4(0u, 0u, 0u, 0u);
        }
        double q[28];
        const double b = (double)w[6];
        q[0] = b * b;
        {
            int qi = 7;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                q[1 + k] = 2.0 * b * (double)w[k];
#pragma unroll
                for (int l = k; l < 6; ++l) {
                    q[qi] = (k == l ? 1.0 : 2.0) * (double)w[k] * (double)w[l];
                    ++qi;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 28; ++c) {
            double t = q[c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) s_q[warp][c] = t;
        }
        __syncthreads();
        if (o < 28) {
            double t = 0.0;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) t += s_q[qq][o];
            reinterpret_cast<double*>(w2s + kWsLn1q)[o] = t;
        }
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 8192) return;
    float4 v0, v1;
    int chunk_base, f;
    float scale;
    if (i < 4096) {                                   // G1: row = out o, K = in
        const int o = i >> 5, in0 = (i & 31) * 8;
        v0 = *reinterpret_cast<const float4*>(params + PLUME_OFF_W2 + o * 256 + in0);
        v1 = *reinterpret_cast<const float4*>(params + PLUME_OFF_W2 + o * 256 + in0 + 4);
        v0 = make_float4(kW2BwdScale * v0.x, kW2BwdScale * v0.y, kW2BwdScale * v0.z, kW2BwdScale * v0.w);
        v1 = make_float4(kW2BwdScale * v1.x, kW2BwdScale * v1.y, kW2BwdScale * v1.z, kW2BwdScale * v1.w);
        chunk_base = kW2SplitG1 + (in0 >> 6) * 2 * kChunkFloats;
        f = (o >> 3) * 64 + ((in0 & 63) >> 3) * 8 + (o & 7);
        scale = 1.0f;
    } else {                                          // G2: row = in (two halves of 128), K = out, values 16 W2[out][in]
        const int j = i - 4096, in = j >> 4, out0 = (j & 15) * 8;
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = kW2BwdScale * params[PLUME_OFF_W2 + (out0 + e) * 256 + in];
        v0 = make_float4(w[0], w[1], w[2], w[3]);
        v1 = make_float4(w[4], w[5], w[6], w[7]);
        const int row = in & 127;
        chunk_base = kW2SplitG2 + ((in >> 7) * 2 + (out0 >> 6)) * 2 * kChunkFloats;
        f = (row >> 3) * 64 + ((out0 & 63) >> 3) * 8 + (row & 7);
        scale = 1.0f;
    }
    uint4 hi, lo;
    tc::split_f16x8(v0, v1, scale, hi, lo);
    reinterpret_cast<uint4*>(w2s + chunk_base)[f] = hi;
    reinterpret_cast<uint4*>(w2s + chunk_base + kChunkFloats)[f] = lo;
}

#ifdef PLUME_TC_MAXNREG
__global__ void __maxnreg__(PLUME_TC_MAXNREG)
#else
__global__ void __launch_bounds__(kTcLaunchThreads, 1)
#endif
ppo_tc_kernel(const float* __restrict__ params, PpoArgs a, const float* __restrict__ w2s, float dz_scale,
              const __grid_constant__ CUtensorMap stash_map) {
    extern __shared__ __align__(128) float sm[];     // no-swizzle operand layouts need 16 B alignment only
    __shared__ uint64_t bar[2];           // "stage free": arrived by tcgen05.commit
    __shared__ uint64_t full[2];          // "operands of the stage written": one arrival per compute thread
    __shared__ uint64_t sdone[2];         // "the stage's activation chunk has been read by its bulk store to the stash"
    __shared__ uint64_t bfull[2];         // "a bulk copy has landed in this half of stage 1" (W2^T chunk or stashed activations)
    __shared__ uint64_t bfree[2];         // "the MMAs reading this half have completed"
    __shared__ uint64_t dzfull;           // "Ph4 has written the tile's dz2 operand": one arrival per compute thread
    __shared__ uint64_t g2half[2];        // "dh1 of inputs [128 h, 128 h + 128) is complete in TMEM"
    __shared__ uint64_t g3done;           // "every G3 MMA of the tile has completed"
    __shared__ uint64_t x0full;           // "Ph0 has written the tile's x operand of the layer-1 GEMM": one arrival per compute thread
    __shared__ uint64_t l1full;           // "the layer-1 B operands have landed" (bulk-copy bytes)
    __shared__ uint64_t zfull[2];         // "the layer-1 pre-activations of a chunk are complete in this TMEM buffer"
    __shared__ uint64_t zread[2];         // "every compute thread has loaded this TMEM buffer": it may take the chunk after next
    __shared__ uint64_t wfull[2];         // "the W2 chunk of a G1 step has landed in the stage's B buffers" (bulk-copy bytes)
    __shared__ uint64_t dyfull[2];        // "the dy1 operand chunk in this buffer is written": one arrival per compute thread
    __shared__ uint64_t pdone[2];         // "the column-sum MMAs reading this buffer have completed"
    __shared__ uint32_t tmem_slot;
    __shared__ float cta_acc[48];         // per-CTA sums of the per-sample scalars (see the flush)
    __shared__ double cta_loss[4];
    // LayerNorm-1 variance as a quadratic form of the 6 inputs: sum_o z_o^2 = q0 + sum_k q1[k] x_k + sum_{k<=l} q2[kl] x_k x_l
    // with z_o = b_o + sum_k x_k w_ko from the CENTRED weights (so mean_o z_o = 0); 28 coefficients per launch,
    // evaluated by the prep kernel
    __shared__ double ln1q[28];

    constexpr int G = kTcGroups;
    constexpr int CW = 128 / G;           // columns per thread in the TMEM epilogues (32)
    constexpr int UPT = 8 / G;            // 16-byte operand units per thread and chunk (2)
    constexpr int SPT = 128 / G;          // samples per thread in the column-oriented phases (32)
    static_assert(CW == 32 && UPT >= 1, "the epilogues read one 32-column TMEM slab per thread");

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // (a value the compiler knows to be warp-uniform)
#ifdef PLUME_TC_TIMELINE
    long long kt_[4] = {0, 0, 0, 0}, kp_[3] = {0, 0, 0};
    const bool kt_on = (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && tid == 0;
    if (kt_on) kt_[0] = clock64();
#endif
    const int wq = warp & 3, cg = (warp >> 2) & (G - 1);   // TMEM lane quarter / column group of this warp
    const int srow = wq * 32 + lane;                       // TMEM lane = tile row owned in the epilogues
    const int r128 = tid & 127, ug = (tid >> 7) & (G - 1); // (row, group) mapping of the producer phases

    // ---- one-time setup ----------------------------------------------------------------------------
    if (tid == 0) {
        tc::mbar_init(&bar[0], 1);
        tc::mbar_init(&bar[1], 1);
        tc::mbar_init(&full[0], kTcThreads);
        tc::mbar_init(&full[1], kTcThreads);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&sdone[i], 1);
            tc::mbar_init(&bfull[i], 1);
            tc::mbar_init(&bfree[i], 1);
            tc::mbar_init(&g2half[i], 1);
            tc::mbar_init(&dyfull[i], kTcThreads);
            tc::mbar_init(&zfull[i], 1);
            tc::mbar_init(&zread[i], kTcThreads);
            tc::mbar_init(&wfull[i], 1);
            tc::mbar_init(&pdone[i], 1);
        }
        tc::mbar_init(&dzfull, kTcThreads);
        tc::mbar_init(&x0full, kTcThreads);
        tc::mbar_init(&l1full, 1);
        tc::mbar_init(&g3done, 1);
        tc::mbar_fence_init();
    }
    if (tid < 48) cta_acc[tid] = 0.0f;
    if (tid < 4) cta_loss[tid] = 0.0;
#ifdef PLUME_TC_TIMELINE
    if (kt_on) kp_[0] = clock64();
#endif
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
#ifdef PLUME_TC_TIMELINE
    if (kt_on) kp_[1] = clock64();
#endif
    float* const exch = sm + TcSmem::exch;
    // exchange slot k of column group g, row r: slot-major, so that the 32 rows of a warp hit 32 banks
    auto EX = [&](int k, int g, int r) -> float& { return exch[(k * G + g) * kTcTile + r]; };
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
#ifdef PLUME_TC_TIMELINE
    if (kt_on) kp_[2] = clock64();
#endif
    if (tid < kTcThreads) {
        // centred layer-1 weights / bias and the quadratic form come from the prep kernel's layer-1 block.  All global loads
        // are issued before the first shared-memory store: loop by loop (load, store, next loop) the staging was six
        // dependent round trips to L2, 4-6 k cycles per launch.
        float t_w1[3], t_p1[3] = {0.0f, 0.0f, 0.0f}, t_p2[3] = {0.0f, 0.0f, 0.0f}, t_wh[2], t_bh = 0.0f;
        double t_q = 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r) t_w1[r] = w2s[kWsW1c + tid + r * kTcThreads];
        if (tid < 256) {
            t_p1[0] = w2s[kWsB1c + tid];
            t_p1[1] = params[PLUME_OFF_G1 + tid];
            t_p1[2] = params[PLUME_OFF_BE1 + tid];
        }
        if (tid < 128) {
            t_p2[0] = params[PLUME_OFF_B2 + tid];
            t_p2[1] = params[PLUME_OFF_G2 + tid];
            t_p2[2] = params[PLUME_OFF_BE2 + tid];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * kTcThreads, k = i >> 3, o = i & 7;
            const int off = o < 5 ? PLUME_OFF_WA + o * 128 + k
                                  : (o == 5 ? PLUME_OFF_WC + k
                                            : (o == 6 ? PLUME_OFF_G2 + k          // LayerNorm-2 gamma / beta ride in the two spare
                                                      : PLUME_OFF_BE2 + k));      // slots of the row: one LDS.128 pair per column
            t_wh[r] = params[off];
        }
        if (tid < 6) t_bh = tid < 5 ? params[PLUME_OFF_BA + tid] : params[PLUME_OFF_BC];
        if (tid < 28) t_q = reinterpret_cast<const double*>(w2s + kWsLn1q)[tid];
#pragma unroll
        for (int r = 0; r < 3; ++r) sm[TcSmem::W1c + tid + r * kTcThreads] = t_w1[r];
        if (tid < 256) {
            sm[TcSmem::P1 + tid] = t_p1[0];
            sm[TcSmem::P1 + 256 + tid] = t_p1[1];
            sm[TcSmem::P1 + 512 + tid] = t_p1[2];
        }
        if (tid < 128) {
            sm[TcSmem::P2 + tid] = t_p2[0];
            sm[TcSmem::P2 + 128 + tid] = t_p2[1];
            sm[TcSmem::P2 + 256 + tid] = t_p2[2];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) sm[TcSmem::Wh + tid + r * kTcThreads] = t_wh[r];
        if (tid < 8) sm[TcSmem::bh + tid] = t_bh;
        if (tid < 28) ln1q[tid] = t_q;
    }
    __syncthreads();

#ifdef PLUME_TC_TIMELINE
    if (kt_on) kt_[1] = clock64();
#endif
    const uint32_t idesc = tc::make_idesc_f16(128, 128);
    float* const xh = sm + TcSmem::xh;
    const float* const W1c = sm + TcSmem::W1c;
    const float* const P1 = sm + TcSmem::P1;
    const float* const P2 = sm + TcSmem::P2;
    const float* const Wh = sm + TcSmem::Wh;
    float* const xt = sm + TcSmem::x;

    uint32_t step = 0;        // ring steps so far (every compute thread counts them identically)
    // Forward (G1) use of the ring.  First 64 KB = the two A stages: stage q = [A_hi 16 KB][A_lo 16 KB], the activation chunk
    // c = 2 p + q in the canonical K-major layout (slot (sg, u, s & 7) at sg * 1024 + u * 128 + (s & 7) * 16 bytes = h1[s][64 c
    // + 8 u .. + 7]).  ONE tensor copy per chunk (cp.async.bulk.tensor, 5-D map built by launch_ppo_tc) scatters the stage
    // into the stash, where the two chunks of a pair (128 inputs) are interleaved per group of 8 samples -- slot (sg, q, u,
    // s & 7) at sg * 2048 + q * 1024 + u * 128 + (s & 7) * 16 bytes, hi in the pair's first 32 KB, lo in the second -- so that
    // a pair comes back as ONE MN-major B operand with N = 128 for G3 (next 8 inputs 128 bytes on, next 8 samples 2048 bytes
    // on).  Second 64 KB = the B stages: stage q holds a W2 chunk as [B_hi 16 KB][B_lo 16 KB].
    // THREE A stages: chunks 0, 1, 3 use the ring's two, chunk 2 a third one in the staging region (idle during the G1 pass;
    // bytes [18 KB, 50 KB): behind the scalar transposition rows of the previous tile, in front of the layer-1 B operands) --
    // a stage is only reused (chunk 3 after chunk 0) long after its tensor store has read it.  With two stages chunks 2 and 3
    // each started 1.7-2 k cycles late, waiting for the store of their stage's previous chunk.
    auto a_stage = [&](int c) -> float* {
        return c == 2 ? sm + TcSmem::xh + 4608 : sm + TcSmem::ring + (c == 1 ? 2 * kChunkFloats : 0);   // chunk 3 -> chunk 0's
    };
    auto b_stage = [&](uint32_t st) -> float* { return sm + TcSmem::ring + (4 + (st & 1u) * 2) * kChunkFloats; };
    // Backward GEMMs (G2, G3): stage 0 of the ring (64 KB) holds the RESIDENT dz2 operand of the tile -- 16-byte slot
    // (og, sg, o & 7) at og * 2048 + sg * 128 + (o & 7) * 16 bytes = dz_scale dz2[8 sg .. 8 sg + 7][o], og = o >> 3,
    // sg = s >> 3; hi in the first 32 KB, lo in the second -- written once by Ph4.  G3 reads it K-major along the samples
    // (A = dz2^T: LBO 128, SBO 2048), G2 reads the SAME bytes MN-major along the outputs (A = dz2: LBO 2048, SBO 128,
    // tc_gemm.cuh make_idesc_f16_a_mn): neither GEMM produces an A operand any more.  Their B operands go through the
    // two halves of stage 1.
    float* const dz_hi = sm + TcSmem::ring;
    float* const dz_lo = dz_hi + 2 * kChunkFloats;
    auto bstage_buf = [&](uint32_t st, int which) -> float* {     // which = 0 B_hi, 1 B_lo
        return sm + TcSmem::ring + (4 + (st & 1u) * 2 + which) * kChunkFloats;
    };
    // producers: the operands of step st are complete in shared memory -> visible to the async proxy, arrive
    auto publish = [&](uint32_t st) {
        tc::fence_proxy_async();
        tc::tc_fence_before();
        mbar_arrive(&full[st & 1u]);
    };
    // (the MMA threads share their schedulers with compute warps: they back off between polls instead of spinning)
#ifndef PLUME_TC_POLL_NS
#define PLUME_TC_POLL_NS 40
#endif

    // G2 step: A = resident dz2, MN-major, K = outputs [64 c, 64 c + 64); B = the W2^T chunk in half `half` of the B region
    const uint32_t idesc_amn = tc::make_idesc_f16_a_mn(128, 128);
    auto step_g2 = [&](uint32_t half, int c) {
        const uint32_t ah = tc::smem_u32(dz_hi) + (uint32_t)(8 * c) * 2048u, al = tc::smem_u32(dz_lo) + (uint32_t)(8 * c) * 2048u;
        const uint32_t bh = tc::smem_u32(bstage_buf(half, 0));
        return tc::make_operands(ah, al, 2 * 2048u, 2048, 128, bh, bh + 16384u, 2 * tc::kLBO, tc::kLBO, tc::kSBO);
    };
    // G3 step for the 128 inputs of an activation pair and the 64 samples of half kh (4 K-steps): A = resident dz2^T, K-major
    // along the samples; B = that half of the stashed pair in half `half` of the B region, MN-major, N = 128
    const uint32_t idesc_g3 = tc::make_idesc_f16_b_mn(128, 128);
    auto step_g3 = [&](uint32_t half, int kh) {
        const uint32_t ah = tc::smem_u32(dz_hi) + (uint32_t)(8 * kh) * 128u, al = tc::smem_u32(dz_lo) + (uint32_t)(8 * kh) * 128u;
        const uint32_t bh = tc::smem_u32(bstage_buf(half, 0));
        return tc::make_operands(ah, al, 2 * 128u, 128, 2048, bh, bh + 16384u, 2 * 2048u, 2048, 128);
    };
    // every MMA of the steps counted so far has completed
    auto wait_all_mma = [&]() {
        const uint32_t last = step - 1;
        tc::mbar_wait_sleep(&bar[last & 1u], (last >> 1) & 1u, PLUME_TC_WAIT_NS);
        tc::tc_fence_after();
    };

    const long long tiles = (a.mb_size + kTcTile - 1) / kTcTile;

    // ---- the MMA issuer: lane 0 of the last warp replays the step sequence of every tile -----------------
    // (a dedicated warp keeps the blocking tcgen05.mma issue out of the producers' instruction streams and
    // lets the ring run on mbarriers only: producers never meet at a CTA barrier inside a GEMM)
    const bool ld = lane == 0;        // the lane of an MMA warp that issues MMAs, copies and commits
    if (warp == kTcThreads / 32) {
        {   // all 32 lanes run the control flow (descriptors stay in uniform registers); `ld` issues
            uint32_t st = 0;                   // producer steps so far: 4 per tile (G1), so (st & 1) == (c & 1)
            uint32_t lt = 0;                   // tiles of this CTA so far
            char* const stash = reinterpret_cast<char*>(const_cast<float*>(w2s) + kWsStash) +
                                (size_t)blockIdx.x * (size_t)kStashFloatsPerCta * 4;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++lt) {
                const bool first_tile = (tile == (long long)blockIdx.x);
                // ---- forward.  Layer 1 runs on the tensor cores too: z1c[s][i] = sum_k x_k w_ik + b_i (centred weights: the
                // LayerNorm-1 mean is 0) as a K = 16 GEMM per chunk of 64 inputs, A = [x_hi | x_lo] (written by Ph0), B1 = [w_hi |
                // w_hi], B2 = [w_lo | 0]: two MMAs give x_hi w_hi + x_lo w_hi + x_hi w_lo.  The result goes to one of two 64-column
                // TMEM buffers in [128,256) (free until G2), from where the compute warps turn it into the chunk's G1 operand.
                if (lt > 0) {                               // the previous tile's G3 and column-sum MMAs have read the ring / xh
                    tc::mbar_wait_warp(&g3done, (lt - 1u) & 1u, PLUME_TC_POLL_NS);
                    tc::mbar_wait_warp(&pdone[0], 1u, PLUME_TC_POLL_NS);
                    tc::mbar_wait_warp(&pdone[1], 1u, PLUME_TC_POLL_NS);
                }
                // (the B operands go to the last 16 KB of the region: the compute threads transpose the previous tile's
                // per-sample scalars through its first 18 KB while this copy is in flight)
                const uint32_t l1a = tc::smem_u32(sm + TcSmem::xh), l1b = l1a + 51200u;
                if (ld) tc::bulk_load(sm + TcSmem::xh + 12800, w2s + kWsL1Ops, (uint32_t)(kL1OpsFloats * 4), &l1full);
                const char* const w2g1 = reinterpret_cast<const char*>(w2s + kW2SplitG1);
                const uint32_t idesc_l1 = tc::make_idesc_f16(128, 64);
                const uint64_t l1_da = tc::make_smem_desc(l1a, 128, 256);
                uint64_t l1_db[4][2];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    l1_db[c][0] = tc::make_smem_desc(l1b + (uint32_t)c * 4096u, 128, 256);
                    l1_db[c][1] = tc::make_smem_desc(l1b + (uint32_t)c * 4096u + 2048u, 128, 256);
                }
                auto issue_l1 = [&](int c) {                // chunk c -> TMEM buffer c & 1
                    if (ld) tc::mma_f16(tmem + 128u + 64u * (uint32_t)(c & 1), l1_da, l1_db[c][0], idesc_l1, 0u);
                    if (ld) tc::mma_f16(tmem + 128u + 64u * (uint32_t)(c & 1), l1_da, l1_db[c][1], idesc_l1, 1u);
                    if (ld) tc::mma_commit(&zfull[c & 1]);
                };
                if (ld) tc::bulk_load(b_stage(0u), w2g1, (uint32_t)kStashChunkBytes, &wfull[0]);
                if (ld) tc::bulk_load(b_stage(1u), w2g1 + kStashChunkBytes, (uint32_t)kStashChunkBytes, &wfull[1]);
                tc::mbar_wait_warp(&x0full, lt & 1u, PLUME_TC_POLL_NS);
#ifdef PLUME_TC_TIMELINE
                long long g1_[12];
                const bool g1_on = blockIdx.x == 0 && lt == 3;
                if (g1_on) g1_[0] = clock64();
#endif
                tc::mbar_wait_warp(&l1full, lt & 1u, PLUME_TC_POLL_NS);
                tc::tc_fence_after();
                issue_l1(0);
                issue_l1(1);
#pragma unroll 1
                for (int c = 0; c < 4; ++c, ++st) {                                                      // G1
                    const uint32_t g1a = tc::smem_u32(a_stage(c)), g1b = tc::smem_u32(b_stage(st));
                    const tc::MmaOperands g1 = tc::make_operands(g1a, g1a + 16384u, 2 * tc::kLBO, tc::kLBO, tc::kSBO, g1b,
                                                                 g1b + 16384u, 2 * tc::kLBO, tc::kLBO, tc::kSBO);
                    if (c < 2) {                            // the chunk's TMEM buffer has been loaded by everyone (long before the
                        tc::mbar_wait_warp(&zread[c], lt & 1u, PLUME_TC_POLL_NS);     // chunk is published): refill it with chunk c + 2
                        tc::tc_fence_after();
                        issue_l1(c + 2);
                    }
                    tc::mbar_wait_warp(&full[st & 1u], (st >> 1) & 1u, PLUME_TC_POLL_NS);
#ifdef PLUME_TC_TIMELINE
                    if (g1_on) g1_[1 + 2 * c] = clock64();
#endif
                    tc::mbar_wait_warp(&wfull[c & 1], (uint32_t)(c >> 1), PLUME_TC_POLL_NS);
#ifdef PLUME_TC_TIMELINE
                    if (g1_on) g1_[2 + 2 * c] = clock64();
#endif
                    tc::tc_fence_after();
                    if (c == 1) {                           // chunk 3 will reuse chunk 0's stage: its store (a turn ago, the only one
                        if (ld) tc::bulk_wait_group_read_all();      // pending) has read it -- told a turn before chunk 3 needs it
                        if (ld) mbar_arrive(&sdone[0]);
                    }
                    // stash the chunk (one tensor copy: the stage's 32 KB -> the chunk's interleaved half of pair c >> 1; the
                    // producers' fence.proxy.async + the `full` barrier made their writes visible to the async proxy), then G1
                    if (ld) tc::tensor_store_5d(&stash_map, a_stage(c), 0, c & 1, 0, 0, (int)blockIdx.x * 2 + (c >> 1));
                    if (ld) tc::bulk_commit_group();
                    tc::issue_split_steps(ld, tmem, g1, 4, idesc, c == 0 ? 0u : 1u);
                    if (ld) tc::mma_commit(&bar[st & 1u]);
#ifdef PLUME_TC_STORE_PROBE
                    if (blockIdx.x == 0 && lt == 3 && c == 0) {     // how long does the tensor store read its 32 KB source?
                        const long long p0 = clock64();
                        if (ld) tc::bulk_wait_group_read_all();
                        const long long p1 = clock64();
                        if (ld) tc::bulk_wait_group_all();
                        if (ld) printf("tensor store of chunk 0: source read %lld cycles after the MMAs were issued, complete %lld\n",
                                       p1 - p0, clock64() - p0);
                    }
#endif
                    // The rest of the turn: the chunk's MMAs complete (~0.9 k cycles, while the compute warps produce chunk
                    // c + 1) -> the stage's B buffers take W2 chunk c + 2, a whole production ahead of its use
                    tc::mbar_wait_warp(&bar[st & 1u], (st >> 1) & 1u, PLUME_TC_POLL_NS);
                    if (c + 2 < 4) {
                        if (ld) tc::bulk_load(b_stage(st), w2g1 + (size_t)(c + 2) * kStashChunkBytes, (uint32_t)kStashChunkBytes,
                                              &wfull[c & 1]);
                    }
                }
#ifdef PLUME_TC_TIMELINE
                if (g1_on && ld)
                    printf("MMA warp, G1 pass (cycles after x0full): chunk full / its W2 landed: %lld/%lld %lld/%lld %lld/%lld %lld/%lld | last "
                           "MMAs complete %lld\n", g1_[1] - g1_[0], g1_[2] - g1_[0], g1_[3] - g1_[0], g1_[4] - g1_[0], g1_[5] - g1_[0],
                           g1_[6] - g1_[0], g1_[7] - g1_[0], g1_[8] - g1_[0], clock64() - g1_[0]);
#endif
                if (ld) {                                   // every store has read its stage: chunks 1, 2, 3
                    tc::bulk_wait_group_read_all();
                    mbar_arrive(&sdone[1]);
                    mbar_arrive(&sdone[0]);
                    mbar_arrive(&sdone[1]);
                }
                // ---- backward: eight 32 KB bulk copies per tile go through the two halves of stage 1, in this order per half h:
                // W2^T chunk h (G2, inputs 0..127), W2^T chunk 2 + h (G2, inputs 128..255), stashed activation chunks h and
                // 2 + h (G3).  bfull / bfree see four phases per tile and half: the parity of phase k is k & 1.
#ifdef PLUME_TC_TIMELINE
                long long is_[24];
                const bool is_on = blockIdx.x == 0 && lt == 3;
#define PLUME_IS(n) if (is_on) is_[n] = clock64()
#else
#define PLUME_IS(n)
#endif
                // (G1's last step has completed -- waited for at the end of its turn: stage 1 is free)
                PLUME_IS(0);
                const char* const w2t = reinterpret_cast<const char*>(w2s + kW2SplitG2);
                if (ld) tc::bulk_load(bstage_buf(0u, 0), w2t, (uint32_t)kStashChunkBytes, &bfull[0]);
                if (ld) tc::bulk_load(bstage_buf(1u, 0), w2t + kStashChunkBytes, (uint32_t)kStashChunkBytes, &bfull[1]);
                tc::mbar_wait_warp(&dzfull, lt & 1u, PLUME_TC_POLL_NS);     // Ph4 has written dz2
                tc::tc_fence_after();
                PLUME_IS(1);
                for (int hN = 0; hN < 2; ++hN) {                                                         // G2
                    for (int c = 0; c < 2; ++c) {
                        const tc::MmaOperands g2 = step_g2((uint32_t)c, c);
                        tc::mbar_wait_warp(&bfull[c], (uint32_t)hN, PLUME_TC_POLL_NS);
                        if (hN == 0) { PLUME_IS(14 + 2 * c); }
                        tc::tc_fence_after();
                        tc::issue_split_steps(ld, tmem + (uint32_t)(128 * hN), g2, 4, idesc_amn, c == 0 ? 0u : 1u);
                        if (ld) tc::mma_commit(&bfree[c]);
                        if (hN == 0) { PLUME_IS(15 + 2 * c); }
                    }
                    if (ld) tc::mma_commit(&g2half[hN]);
                    PLUME_IS(2 + 3 * hN);
                    if (hN == 0) {
                        for (int c = 0; c < 2; ++c) {
                            tc::mbar_wait_warp(&bfree[c], 0u, PLUME_TC_POLL_NS);
                            PLUME_IS(3 + c);
                            if (ld) tc::bulk_load(bstage_buf((uint32_t)c, 0), w2t + (size_t)(2 + c) * kStashChunkBytes,
                                          (uint32_t)kStashChunkBytes, &bfull[c]);
                        }
                    }
                }
                // G3: the four stashed chunks; nothing for the compute warps to do
                if (ld) tc::bulk_wait_group_all();                  // the stash holds the tile's four chunks
                for (int r = 0; r < 2; ++r) {              // pair r = inputs [128 r, 128 r + 128); c = half of the samples
                    for (int c = 0; c < 2; ++c) {
                        tc::mbar_wait_warp(&bfree[c], (uint32_t)((1 + r) & 1), PLUME_TC_POLL_NS);
                        PLUME_IS(6 + 4 * r + c);
                        const char* const src = stash + (size_t)r * (2 * kStashChunkBytes) + (size_t)c * 16384;
                        if (ld) tc::bulk_load2(bstage_buf((uint32_t)c, 0), src, bstage_buf((uint32_t)c, 1), src + kStashChunkBytes, 16384u,
                                       &bfull[c]);
                    }
                    for (int c = 0; c < 2; ++c) {
                        const tc::MmaOperands g3 = step_g3((uint32_t)c, c);
                        tc::mbar_wait_warp(&bfull[c], (uint32_t)((2 + r) & 1), PLUME_TC_POLL_NS);
                        PLUME_IS(8 + 4 * r + c);
                        tc::tc_fence_after();
                        tc::issue_split_steps(ld, tmem + (uint32_t)(256 + 128 * r), g3, 4, idesc_g3, (first_tile && c == 0) ? 0u : 1u);
                        if (ld) tc::mma_commit(&bfree[c]);
                    }
                }
                if (ld) tc::mma_commit(&g3done);
#ifdef PLUME_TC_TIMELINE
                if (is_on) {
                    tc::mbar_wait_warp(&g3done, lt & 1u, PLUME_TC_POLL_NS);
                    const long long e = clock64();
                    if (ld) printf("issuer (cycles after G1 done): dz ready %lld | G2h0 issued %lld | bfree0 %lld | bfree1 %lld | G2h1 issued %lld | "
                           "G3: bfree0 %lld bfree1 %lld | H0 landed %lld H1 landed %lld | bfree0 %lld bfree1 %lld | H2 landed %lld H3 landed "
                           "%lld | all done %lld || G2h0: chunk 0 landed %lld, issued %lld, chunk 1 landed %lld, issued %lld\n", is_[1] - is_[0], is_[2] - is_[0], is_[3] - is_[0], is_[4] - is_[0], is_[5] - is_[0],
                           is_[6] - is_[0], is_[7] - is_[0], is_[8] - is_[0], is_[9] - is_[0], is_[10] - is_[0], is_[11] - is_[0],
                           is_[12] - is_[0], is_[13] - is_[0], e - is_[0], is_[14] - is_[0], is_[15] - is_[0], is_[16] - is_[0],
                           is_[17] - is_[0]);
                }
#endif
            }
        }
    } else if (warp == kTcThreads / 32 + 1) {
        // ---- the column-sum GEMM of the LayerNorm-1 backward (its own issuing thread: its operands become ready while the
        // other MMA thread sits in the bulk-copy waits of G3) -----------------------------------------------------------------
        //   P[i][c] = sum_s dy1[s][i] X[s][c],   X[s][.] = {rstd1 x_0..x_5, rstd1, 1}
        // per chunk of 64 inputs: A = the dy1 chunk written by Ph6 (fp16 hi / lo in the layout of an activation chunk, read
        // MN-major with M = inputs: next 8 inputs 128 bytes on, next 8 samples 1024 bytes on; the instruction's M is 128, rows
        // 64..127 read whatever follows and land in TMEM lanes that nobody reads), B = X (one 16-byte slot per sample, read
        // MN-major with N = 16: columns 8..15 are the next sample's slot, ignored), K = the tile's 128 samples; the result
        // overwrites the first 16 of the chunk's own dh1 columns, which every thread has consumed by then.
        {   // all 32 lanes run the control flow (descriptors stay in uniform registers); `ld` issues
            const uint32_t idesc_p = tc::make_idesc_f16(PLUME_TC_PM, PLUME_TC_PN) | (1u << 15) | (1u << 16);
            const uint32_t xph = tc::smem_u32(sm + TcSmem::x), xpl = xph + 2048u;
            for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const uint32_t ah = tc::smem_u32(sm + TcSmem::xh) + (uint32_t)(c & 1) * 32768u, al = ah + 16384u;
                    const tc::MmaOperands po = tc::make_operands(ah, al, 2 * 1024u, 1024, 128, xph, xpl, 2 * 128u, 128, 16);
                    tc::mbar_wait_warp(&dyfull[c & 1], (uint32_t)(c >> 1), PLUME_TC_POLL_NS);
                    tc::tc_fence_after();
                    tc::issue_split_steps(ld, tmem + (uint32_t)(64 * c), po, 8, idesc_p, 0u);
                    if (ld) tc::mma_commit(&pdone[c & 1]);
                }
            }
        }
    } else {
    // ---- persistent accumulators (everything else is reduced into shared memory tile by tile) -----------
    float g_b2 = 0.0f, g_g2 = 0.0f, g_be2 = 0.0f;                                  // (output r128, sample group ug)
    float2 g_wh2[3] = {f2(0.0f, 0.0f), f2(0.0f, 0.0f), f2(0.0f, 0.0f)};           // head-weight gradients, 3 pairs
    float Pacc[8];                                   // the 8 column sums of input 64 cg + 16 wq + lane (lanes 0..15),
#pragma unroll                                        // in units of dz_scale * kW2BwdScale
    for (int c = 0; c < 8; ++c) Pacc[c] = 0.0f;

    float* const pf = sm + TcSmem::pf;
    // gather of one tile's samples into pf with 4-byte cp.async (one thread per sample; rows beyond the minibatch = 0);
    // the next tile's gather (its Feistel index is ~300 dependent instructions) is issued by the four warps that have
    // nothing to do during the loss section of Ph3, and waited for at the next tile's Ph0
    auto prefetch_tile = [&](long long tl, int row) {
        if (tl >= tiles) return;
        const long long b0 = tl * kTcTile;
        float* dst = pf + row * 12;
        if (b0 + row < a.mb_size) {
            const long long pos = a.mb_start + b0 + row;
            const long long idx = a.perm ? a.perm[pos]
                                         : (long long)feistel_permute((uint64_t)pos, (uint64_t)a.batch.total,
                                                                      a.perm_seed, (uint32_t)a.epoch);
            if (a.batch.packed) {           // one 48-byte record per sample (plume_ppo_pack): three 16-byte copies
                const float* src = a.batch.packed + idx * 12;
                cp_async16(dst, src);
                cp_async16(dst + 4, src + 4);
                cp_async16(dst + 8, src + 8);
                return;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) cp_async4(dst + k, a.batch.obs + idx * 6 + k);
            cp_async4(dst + 6, a.batch.advantages + idx);
            cp_async4(dst + 7, a.batch.returns + idx);
            cp_async4(dst + 8, a.batch.old_values + idx);
            cp_async4(dst + 9, a.batch.old_log_probs + idx);
            cp_async4(dst + 10, a.batch.actions + idx);
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) dst[k] = 0.0f;
        }
    };

    uint32_t lt = 0;          // tiles of this CTA so far
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++lt) {
#ifdef PLUME_TC_TIMELINE
        long long tl_[24];
        const bool tl_on = blockIdx.x == 0 && tid == 0 && tile == (long long)blockIdx.x + 3 * gridDim.x;
#define PLUME_TL(n) if (tl_on) tl_[n] = clock64()
#else
#define PLUME_TL(n)
#endif
        const long long base = tile * kTcTile;
        const int n_valid = (int)((a.mb_size - base) < kTcTile ? (a.mb_size - base) : kTcTile);

        PLUME_TL(0);
        // ---- Ph0: this tile's samples: gathered by cp.async during the previous tile (first tile: now) --------
        // pf[s] = {obs 0..5, adv, ret, old value, old logp, action (int bits), 0}
        if (tile == (long long)blockIdx.x && tid < kTcTile) prefetch_tile(tile, tid);
        cp_async_wait_all();
        compute_sync();
        float s_adv = 0.0f, s_ret = 0.0f, s_vold = 0.0f, s_lpold = 0.0f;
        int s_act = 0;
        if (tid < kTcTile) {
            const float4 q0 = *reinterpret_cast<const float4*>(pf + tid * 12);
            const float4 q1 = *reinterpret_cast<const float4*>(pf + tid * 12 + 4);
            const float4 q2 = *reinterpret_cast<const float4*>(pf + tid * 12 + 8);
            s_adv = q1.z;
            s_ret = q1.w;
            s_vold = q2.x;
            s_lpold = q2.y;
            s_act = __float_as_int(q2.z);
            // LayerNorm-1 rstd of this sample from the quadratic form (float64: 48 operations instead of the
            // 1800 FMAs of evaluating all 256 pre-activations once more just for their sum of squares)
            const double xd[6] = {(double)q0.x, (double)q0.y, (double)q0.z, (double)q0.w, (double)q1.x, (double)q1.y};
            // (six independent chains -- one per row of the quadratic form -- instead of one of 27 dependent DFMAs: only four
            // of the sixteen warps are here and the others wait for them)
            double part[6];
            int qi = 7;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double row = ln1q[1 + k];
#pragma unroll
                for (int l = k; l < 6; ++l) {
                    row = fma(ln1q[qi], xd[l], row);
                    ++qi;
                }
                part[k] = row * xd[k];
            }
            const double ssq = ln1q[0] + ((part[0] + part[1]) + (part[2] + part[3])) + (part[4] + part[5]);
            // float32 from here: rsqrt of a float64-accurate mean square (one Newton step brings it to <= 1 ulp)
            const float msq = (float)(ssq * (1.0 / 256.0) + (double)kLnEps);
            float rstd1_s = rsqrtf(msq);
            rstd1_s = rstd1_s * fmaf(-0.5f * msq * rstd1_s, rstd1_s, 1.5f);
            *reinterpret_cast<float4*>(xt + tid * 8) = q0;
            *reinterpret_cast<float4*>(xt + tid * 8 + 4) = make_float4(q1.x, q1.y, rstd1_s, 0.0f);
            // A operand of the layer-1 GEMM (in the xhat2 / staging region, idle until Ph3): row = sample, K slot 0 = fp16 hi of
            // (x_0..x_5, 1, 0), K slot 1 = the fp16 remainders; slot (s, k) at (s >> 3) * 256 + k * 128 + (s & 7) * 16 bytes
            {
                uint4 hi, lo;
                tc::split_f16x8(q0, make_float4(q1.x, q1.y, 1.0f, 0.0f), 1.0f, hi, lo);
                uint4* xa = reinterpret_cast<uint4*>(xh);
                const int f = (tid >> 3) * 16 + (tid & 7);
                xa[f] = hi;
                xa[f + 8] = lo;
            }
            // d loss / d (logits, value) of this tile: the two loss halves of Ph3 add into it
            *reinterpret_cast<float4*>(sm + TcSmem::dout + tid * 8) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            *reinterpret_cast<float4*>(sm + TcSmem::dout + tid * 8 + 4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        tc::fence_proxy_async();
        mbar_arrive(&x0full);     // the MMA thread starts the layer-1 GEMM of the first two chunks
        compute_sync();

        PLUME_TL(1);
        // ---- Ph1: rstd1 of this thread's sample (r128), computed in Ph0 ------------------------------------
        const float2 rs2 = splat2(xt[r128 * 8 + 6]);

        // ---- Ph2: G1 forward, K = 256 inputs in 4 chunks of 64; thread = (sample r128, 8/G of the 8 slots) ----
        // (the previous tile's G3 reads dz2 in stage 0 and the stashed chunks in stage 1 until here: its tail overlaps that
        // tile's scalar sums and this tile's Ph0)
        if (lt > 0) tc::mbar_wait_sleep(&g3done, (lt - 1u) & 1u, PLUME_TC_WAIT_NS);
        for (int c = 0; c < 4; ++c) {
            const uint32_t st = step;
            // The only reused stage is chunk 3's (= chunk 0's): chunk 0's store signals `sdone` at the MMA warp's turn 2 (the
            // barrier's next phase needs this thread's own chunk-3 arrival: the parity wait cannot fall a phase behind);
            // chunk 0's MMAs are waited for before chunk 2 is published, below, for the same reason.
            if (c == 3) {
                tc::mbar_wait_sleep(&sdone[0], 0u, PLUME_TC_WAIT_NS);
                PLUME_TL(23);
            }
            uint4* ah = reinterpret_cast<uint4*>(a_stage(c));
            uint4* al = ah + 1024;
            // the centred pre-activations of this thread's sample and 16 inputs (64 c + 16 ug ..) from the chunk's TMEM buffer
            // (loading them one chunk ahead, before the previous chunk is published, was tried: no gain)
            float z[16];
            tc::mbar_wait_sleep(&zfull[c & 1], (uint32_t)(c >> 1), PLUME_TC_WAIT_NS);
            tc::tc_fence_after();
            PLUME_TL(18 + c);
            tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(128 + 64 * (c & 1) + 16 * ug), z);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            if (c < 2) mbar_arrive(&zread[c]);      // (chunks 0 and 1: their buffers take chunks 2 and 3)
#pragma unroll
            for (int uu = 0; uu < UPT; ++uu) {
                const int u = UPT * ug + uu;
                float4 hq[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int in0 = 64 * c + 8 * u + 4 * q;
                    const float4 g = *reinterpret_cast<const float4*>(P1 + 256 + in0);
                    const float4 be = *reinterpret_cast<const float4*>(P1 + 512 + in0);
                    const float2 z01 = f2(z[8 * uu + 4 * q], z[8 * uu + 4 * q + 1]), z23 = f2(z[8 * uu + 4 * q + 2], z[8 * uu + 4 * q + 3]);
                    const float2 y01 = __ffma2_rn(__fmul2_rn(z01, rs2), f2(g.x, g.y), f2(be.x, be.y));
                    const float2 y23 = __ffma2_rn(__fmul2_rn(z23, rs2), f2(g.z, g.w), f2(be.z, be.w));
                    hq[q] = make_float4(fmaxf(y01.x, 0.0f), fmaxf(y01.y, 0.0f), fmaxf(y23.x, 0.0f), fmaxf(y23.y, 0.0f));
                }
                uint4 hi, lo;
                tc::split_f16x8(hq[0], hq[1], 1.0f, hi, lo);
                const int f = (r128 >> 3) * 64 + u * 8 + (r128 & 7);
                ah[f] = hi;
                al[f] = lo;
            }
            // (chunk 0's MMAs have long completed; waiting here -- before this thread's arrival lets the MMA warp issue chunk 2's,
            // the next phase of the same barrier -- keeps the parity wait safe)
            if (c == 2) tc::mbar_wait_sleep(&bar[0], 0u, PLUME_TC_WAIT_NS);
            publish(st);          // issuer: G1 -> columns [0,128), then the chunk's bulk store to the stash
            if (c == 2) { PLUME_TL(22); }
            ++step;
        }
        PLUME_TL(2);
        wait_all_mma();
        tc::mbar_wait(&sdone[0], 1u);     // chunks 2 and 3 have left for the stash: Ph4 / G2 may overwrite their stages
        tc::mbar_wait(&sdone[1], 1u);

        PLUME_TL(3);
        // ---- Ph3: LN2, heads, loss, LN2-backward means: thread = (sample srow, 32 of the 128 outputs) ----
        float v[CW];
        {
            const int c0 = CW * cg;
            const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0;
            tc::tmem_ld32(taddr, v);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            float sum = 0.0f;
            constexpr float kUnW = 1.0f / kW2BwdScale;      // the accumulator holds h1 . (16 W2)^T; exact power of two
#pragma unroll
            for (int j4 = 0; j4 < CW / 4; ++j4) {
                const float4 b2 = *reinterpret_cast<const float4*>(P2 + c0 + 4 * j4);
                v[4 * j4 + 0] = fmaf(v[4 * j4 + 0], kUnW, b2.x);
                v[4 * j4 + 1] = fmaf(v[4 * j4 + 1], kUnW, b2.y);
                v[4 * j4 + 2] = fmaf(v[4 * j4 + 2], kUnW, b2.z);
                v[4 * j4 + 3] = fmaf(v[4 * j4 + 3], kUnW, b2.w);
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) sum += v[j];
            // LayerNorm-2 statistics with ONE exchange: each column group contributes its sum and the sum of squares about
            // its OWN mean; the groups are merged with the pairwise update M2 = sum_g [M2_g + 32 (mean_g - mean)^2]
            // (as accurate as the two-pass form, one barrier less per tile)
            const float mean_g = sum * (1.0f / (float)CW);
            float sq = 0.0f;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const float d = v[j] - mean_g;
                sq = fmaf(d, d, sq);
            }
            EX(6, cg, srow) = sum;
            EX(7, cg, srow) = sq;
            quarter_sync(wq);
            float tot = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) tot += EX(6, g, srow);
            const float mean = tot * (1.0f / 128.0f);
            float m2 = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float dm = EX(6, g, srow) * (1.0f / (float)CW) - mean;
                m2 += fmaf((float)CW * dm, dm, EX(7, g, srow));
            }
            const float rstd2 = 1.0f / sqrtf(m2 * (1.0f / 128.0f) + kLnEps);
            PLUME_TL(10);
            float2 head2[3] = {f2(0.0f, 0.0f), f2(0.0f, 0.0f), f2(0.0f, 0.0f)};     // the 6 head outputs as 3 pairs
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int o = c0 + j;
                const float x_hat = (v[j] - mean) * rstd2;
                v[j] = x_hat;
                const float4 w0 = *reinterpret_cast<const float4*>(Wh + o * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(Wh + o * 8 + 4);          // .z = gamma2, .w = beta2
                const float h2 = fmaxf(fmaf(x_hat, w1.z, w1.w), 0.0f);
                const float2 h22 = splat2(h2);
                head2[0] = __ffma2_rn(h22, f2(w0.x, w0.y), head2[0]);
                head2[1] = __ffma2_rn(h22, f2(w0.z, w0.w), head2[1]);
                head2[2] = __ffma2_rn(h22, f2(w1.x, w1.y), head2[2]);
            }
#pragma unroll
            for (int q = 0; q < CW / 4; ++q)
                *reinterpret_cast<float4*>(xh + srow * kXhStride + c0 + 4 * q) =
                    make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                EX(2 * k, cg, srow) = head2[k].x;
                EX(2 * k + 1, cg, srow) = head2[k].y;
            }
            quarter_sync(wq);
            PLUME_TL(11);
            // The per-sample loss runs as two halves in two warps of the SAME scheduler (cg = 0: surrogate + value,
            // cg = 1: entropy; warps wq and 4 + wq), each from its own softmax; both add their part of d loss / d logits
            // into dout[.][0..4] (zeroed in Ph0; two addends, so the order does not matter).  dout[.][5] = d loss / d value,
            // dout[.][6], [7] and sc[.][3] carry the sample's policy / value / entropy terms: their sums over the tile
            // (and those of dout = the head-bias gradients) are taken by all 16 warps in the scalar-sum phase at the end of
            // the tile, not by these warps while the others wait.
            if (cg == 2) prefetch_tile(tile + gridDim.x, srow);      // pf was consumed in Ph0; see prefetch_tile
            if (cg < 2) {
                float* dop = sm + TcSmem::dout + srow * 8;
                float l_ent = 0.0f, dv = 0.0f, l_pol = 0.0f, l_val = 0.0f;
                if (srow < n_valid) {
                    float o6[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        float t = sm[TcSmem::bh + k];
#pragma unroll
                        for (int g = 0; g < G; ++g) t += EX(k, g, srow);
                        o6[k] = t;
                    }
                    float p[5];
                    softmax5(o6, p);
                    if (cg == 0) {
                        bool bad = false;
#pragma unroll
                        for (int k = 0; k < 5; ++k) bad |= isnan(o6[k]);
                        if (bad) atomicExch(a.nan_flag, 1);                         // train_ppo2.0.py:57-61
                        const PolicyValuePart pv = ppo_policy_value_part(p, o6[5], s_act, s_adv, s_ret, s_vold, s_lpold,
                                                                         a.clip_eps);
#pragma unroll
                        for (int k = 0; k < 5; ++k) atomicAdd(dop + k, pv.dpol[k] * a.inv_global);
                        dv = pv.dv * a.inv_global;
                        l_pol = pv.pol;
                        l_val = pv.val;
                    } else {
                        const EntropyPart en = ppo_entropy_part(p, a.entropy_beta);
#pragma unroll
                        for (int k = 0; k < 5; ++k) atomicAdd(dop + k, en.dent[k] * a.inv_global);
                        l_ent = en.ent;
                    }
                }
                if (cg == 0) {
                    dop[5] = dv;
                    dop[6] = l_pol;
                    dop[7] = l_val;
                } else {
                    sm[TcSmem::sc + srow * 4 + 3] = l_ent;
                }
                PLUME_TL(12);
            }
            quarter_sync(wq);
            PLUME_TL(13);
            const float4 d0 = *reinterpret_cast<const float4*>(sm + TcSmem::dout + srow * 8);
            const float4 d1 = *reinterpret_cast<const float4*>(sm + TcSmem::dout + srow * 8 + 4);
            const float2 d01 = f2(d0.x, d0.y), d23 = f2(d0.z, d0.w), d45 = f2(d1.x, d1.y);
            float m1p = 0.0f, m2p = 0.0f;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int o = c0 + j;
                const float4 w0 = *reinterpret_cast<const float4*>(Wh + o * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(Wh + o * 8 + 4);
                const float g2 = w1.z;
                const float y = fmaf(v[j], g2, w1.w);
                float2 dh2 = __fmul2_rn(d01, f2(w0.x, w0.y));
                dh2 = __ffma2_rn(d23, f2(w0.z, w0.w), dh2);
                dh2 = __ffma2_rn(d45, f2(w1.x, w1.y), dh2);
                const float dh = dh2.x + dh2.y;
                const float dxh = (y > 0.0f) ? dh * g2 : 0.0f;
                m1p += dxh;
                m2p = fmaf(dxh, v[j], m2p);
            }
            EX(6, cg, srow) = m1p;
            EX(7, cg, srow) = m2p;
            quarter_sync(wq);
            if (cg == 0) {
                float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    t1 += EX(6, g, srow);
                    t2 += EX(7, g, srow);
                }
                float* scp = sm + TcSmem::sc + srow * 4;          // [3] = the sample's entropy term (written above)
                scp[0] = rstd2;
                scp[1] = t1 * (1.0f / 128.0f);
                scp[2] = t2 * (1.0f / 128.0f);
            }
            compute_sync();
        }

        PLUME_TL(4);
        // ---- Ph4: heads + LN2 backward per column: thread = (output r128, 128/G of the 128 samples) ------
        {
            const int o = r128;
            const float g2 = P2[128 + o], be2 = P2[256 + o];
            const float2 wrow2[3] = {f2(Wh[o * 8], Wh[o * 8 + 1]), f2(Wh[o * 8 + 2], Wh[o * 8 + 3]),
                                     f2(Wh[o * 8 + 4], Wh[o * 8 + 5])};
            uint4* const dzh = reinterpret_cast<uint4*>(dz_hi);
            uint4* const dzl = reinterpret_cast<uint4*>(dz_lo);
#pragma unroll
            for (int g8 = 0; g8 < SPT / 8; ++g8) {
                float dzv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int s = SPT * ug + 8 * g8 + i;
                    const float x_hat = xh[s * kXhStride + o];
                    const float4 d0 = *reinterpret_cast<const float4*>(sm + TcSmem::dout + s * 8);
                    const float4 d1 = *reinterpret_cast<const float4*>(sm + TcSmem::dout + s * 8 + 4);
                    const float4 sc = *reinterpret_cast<const float4*>(sm + TcSmem::sc + s * 4);
                    const float2 d01 = f2(d0.x, d0.y), d23 = f2(d0.z, d0.w), d45 = f2(d1.x, d1.y);
                    const float y = fmaf(x_hat, g2, be2);
                    const float2 h2v = splat2(fmaxf(y, 0.0f));
                    float2 dh2 = __fmul2_rn(d01, wrow2[0]);
                    dh2 = __ffma2_rn(d23, wrow2[1], dh2);
                    dh2 = __ffma2_rn(d45, wrow2[2], dh2);
                    g_wh2[0] = __ffma2_rn(d01, h2v, g_wh2[0]);
                    g_wh2[1] = __ffma2_rn(d23, h2v, g_wh2[1]);
                    g_wh2[2] = __ffma2_rn(d45, h2v, g_wh2[2]);
                    const float dy = (y > 0.0f) ? dh2.x + dh2.y : 0.0f;
                    g_g2 = fmaf(dy, x_hat, g_g2);
                    g_be2 += dy;
                    const float dz = sc.x * (dy * g2 - sc.y - x_hat * sc.z);
                    g_b2 += dz;
                    dzv[i] = dz * dz_scale;                         // exact: a power of two
                }
                // eight consecutive samples of output o = one 16-byte slot of the resident dz2 operand (hi and lo)
                uint4 hi, lo;
                tc::split_f16x8(make_float4(dzv[0], dzv[1], dzv[2], dzv[3]), make_float4(dzv[4], dzv[5], dzv[6], dzv[7]), 1.0f,
                                hi, lo);
                const int f = (o >> 3) * 128 + ((SPT / 8) * ug + g8) * 8 + (o & 7);
                dzh[f] = hi;
                dzl[f] = lo;
            }
        }
        // the tile's dz2 operand is complete: visible to the async proxy, hand over to the MMA thread, which runs G2 (B = the
        // W2^T chunks it bulk-copies into the halves of stage 1; dh1 -> TMEM columns [0,256)) and then G3 (B = the stashed
        // activation chunks; the accumulator holds dz_scale dW2 until the flush) on its own
        tc::fence_proxy_async();
        tc::tc_fence_before();
        mbar_arrive(&dzfull);

        PLUME_TL(5);
        PLUME_TL(6);
        PLUME_TL(7);
        // ---- Ph6: LN1 backward: dy1 from TMEM, per-sample means, column sums P through shared memory -----
        {
            float2 m1p2 = f2(0.0f, 0.0f), m2p2 = f2(0.0f, 0.0f);         // even / odd inputs of this thread's slab
            const float2 un2 = splat2(1.0f / (dz_scale * kW2BwdScale));    // exact: both are powers of two
            // this thread's sample: inputs + rstd1
            const float4 sx0 = *reinterpret_cast<const float4*>(xt + srow * 8);
            const float4 sx1 = *reinterpret_cast<const float4*>(xt + srow * 8 + 4);
            const float xs[6] = {sx0.x, sx0.y, sx0.z, sx0.w, sx1.x, sx1.y};
            const float2 xs2[6] = {splat2(sx0.x), splat2(sx0.y), splat2(sx0.z), splat2(sx0.w), splat2(sx1.x), splat2(sx1.y)};
            const float2 srs2 = splat2(sx1.z);
            // The B operand of the column-sum GEMM takes the place of the x tile (nobody reads it after this point): per sample
            // one 16-byte slot of fp16 hi and one of lo holding X[s][.] = {rstd1 x_0..x_5, rstd1, 1}.
            compute_sync();
            if (tid < kTcTile) {           // tid < 128: srow == tid, cg == 0
                const float r = sx1.z;
                uint4 hi, lo;
                tc::split_f16x8(make_float4(r * sx0.x, r * sx0.y, r * sx0.z, r * sx0.w), make_float4(r * sx1.x, r * sx1.y, r, 1.0f),
                                1.0f, hi, lo);
                reinterpret_cast<uint4*>(xt)[tid] = hi;
                reinterpret_cast<uint4*>(xt)[kTcTile + tid] = lo;
            }
            // Per chunk c of 64 inputs: thread = (sample srow, the 16 inputs 64 c + 16 cg ..).  dh1 comes from the TMEM columns
            // of those inputs, the ReLU mask and xhat1 from a re-evaluation of layer 1; the masked dh1 (still in units of
            // dz_scale * kW2BwdScale, O(1)) goes, split into fp16 hi / lo, into the chunk's operand buffer (the xhat2 / staging
            // region: two buffers of 32 KB) as A of the column-sum GEMM, which the second MMA thread issues as soon as all
            // compute threads have arrived.  No CTA barrier inside the loop.
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((c & 1) == 0) {
                    tc::mbar_wait_sleep(&g2half[c >> 1], lt & 1u, PLUME_TC_WAIT_NS);   // dh1 of this half is complete (the other half's MMAs may still run)
                    tc::tc_fence_after();
                }
                float dv[16];
                tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(64 * c + 16 * cg), dv);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int in0 = 64 * c + 16 * cg + 4 * j4;
                    const float4 b = *reinterpret_cast<const float4*>(P1 + in0);
                    float2 zz[2] = {f2(b.x, b.y), f2(b.z, b.w)};
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const float4 w = *reinterpret_cast<const float4*>(W1c + k * 256 + in0);
                        zz[0] = __ffma2_rn(xs2[k], f2(w.x, w.y), zz[0]);
                        zz[1] = __ffma2_rn(xs2[k], f2(w.z, w.w), zz[1]);
                    }
                    const float4 g = *reinterpret_cast<const float4*>(P1 + 256 + in0);
                    const float4 be = *reinterpret_cast<const float4*>(P1 + 512 + in0);
                    const float2 gg[2] = {f2(g.x, g.y), f2(g.z, g.w)}, bb[2] = {f2(be.x, be.y), f2(be.z, be.w)};
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const float2 x_hat = __fmul2_rn(zz[jj], srs2);
                        const float2 y = __ffma2_rn(x_hat, gg[jj], bb[jj]);
                        const float d0 = y.x > 0.0f ? dv[4 * j4 + 2 * jj] : 0.0f;
                        const float d1 = y.y > 0.0f ? dv[4 * j4 + 2 * jj + 1] : 0.0f;
                        dv[4 * j4 + 2 * jj] = d0;
                        dv[4 * j4 + 2 * jj + 1] = d1;
                        const float2 t = __fmul2_rn(__fmul2_rn(f2(d0, d1), un2), gg[jj]);
                        m1p2 = __fadd2_rn(m1p2, t);
                        m2p2 = __ffma2_rn(t, x_hat, m2p2);
                    }
                }
                if (c >= 2) tc::mbar_wait(&pdone[c & 1], 0u);       // the column-sum MMAs of chunk c - 2 have read the buffer
                uint4* const dyh = reinterpret_cast<uint4*>(xh + (c & 1) * 8192);
                uint4* const dyl = dyh + 1024;
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    uint4 hi, lo;
                    tc::split_f16x8(make_float4(dv[8 * uu], dv[8 * uu + 1], dv[8 * uu + 2], dv[8 * uu + 3]),
                                    make_float4(dv[8 * uu + 4], dv[8 * uu + 5], dv[8 * uu + 6], dv[8 * uu + 7]), 1.0f, hi, lo);
                    const int f = (srow >> 3) * 64 + (2 * cg + uu) * 8 + (srow & 7);
                    dyh[f] = hi;
                    dyl[f] = lo;
                }
                tc::fence_proxy_async();
                tc::tc_fence_before();
                mbar_arrive(&dyfull[c & 1]);
                PLUME_TL(14 + c);
            }
        PLUME_TL(8);
            // The per-sample scalars first: they need neither G3 nor the last column-sum GEMM (queued behind G3's MMAs in the
            // tensor pipe), only the first dy1 buffer back -- chunk 2's GEMM has read it.  Their exchange area (bytes [20 KB,
            // 24 KB)) and transposition rows ([0, 18 KB)) lie inside that buffer.
            tc::mbar_wait(&pdone[0], 1u);
            float* const ex6 = xh + 5120;
            ex6[(0 * G + cg) * kTcTile + srow] = m1p2.x + m1p2.y;
            ex6[(1 * G + cg) * kTcTile + srow] = m2p2.x + m2p2.y;
            quarter_sync(wq);
            if (cg == 0) {          // per-sample scalar sums of the layer-1 backward -> per-CTA accumulators
                float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    t1 += ex6[(0 * G + g) * kTcTile + srow];
                    t2 += ex6[(1 * G + g) * kTcTile + srow];
                }
                const float m1 = t1 * (1.0f / 256.0f), m2 = t2 * (1.0f / 256.0f);
                const float rs = sx1.z;
                const float a1 = rs * m1, a2 = rs * rs * m2;
                float red[35];
                red[0] = a1;
                red[13] = a2;
                {
                    int qi = 0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        red[1 + k] = a1 * xs[k];
                        const float ax = a2 * xs[k];
                        red[7 + k] = ax;
#pragma unroll
                        for (int k2 = k; k2 < 6; ++k2) {
                            red[14 + qi] = ax * xs[k2];
                            ++qi;
                        }
                    }
                }
                // sums over the tile's 128 samples: transpose through the (now idle) staging region and let all 16
                // warps add two or three of the 35 rows each, instead of 175 shuffles in each of these 4 warps
#pragma unroll
                for (int k = 0; k < 35; ++k) xh[k * kTcTile + srow] = red[k];
            }
            compute_sync();
            // rows 0..34: the layer-1 scalars staged above; 35..40: d loss / d head outputs (= head-bias gradients);
            // 41, 42, 43: policy / value / entropy terms of the loss (tile sums in float32, accumulated in float64).
            // Warp w owns rows w, w + 16, w + 32: three independent load + shuffle chains in flight.
            {
                float sred[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int k = warp + 16 * r;
                    sred[r] = 0.0f;
                    if (k < 44) {
                        const float* rowp = k < 35 ? xh + k * kTcTile + lane
                                                   : (k < 43 ? sm + TcSmem::dout + lane * 8 + (k - 35)
                                                             : sm + TcSmem::sc + lane * 4 + 3);
                        const int rs = k < 35 ? 32 : (k < 43 ? 32 * 8 : 32 * 4);
                        sred[r] = (rowp[0] + rowp[rs]) + (rowp[2 * rs] + rowp[3 * rs]);
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int r = 0; r < 3; ++r) sred[r] += __shfl_xor_sync(0xffffffffu, sred[r], off);
                if (lane == 0) {                            // every row belongs to exactly one warp
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const int k = warp + 16 * r;
                        if (k < 41) cta_acc[k] += sred[r];
                        else if (k < 44) cta_loss[k - 40] += (double)sred[r];
                    }
                }
            }
            // the column sums last: chunk 3's GEMM (and with it every earlier one) is complete
            tc::mbar_wait(&pdone[1], 1u);
            tc::tc_fence_after();
            {                                  // an M = 64 accumulator keeps row r in TMEM lane 32 (r / 16) + r % 16: lanes 0..15
                float pv[8];                   // of warp (wq, cg) hold inputs 64 cg + 16 wq .. of chunk cg
                tc::tmem_ld8(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(64 * cg), pv);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
#pragma unroll
                for (int k = 0; k < 8; ++k) Pacc[k] += pv[k];      // (lanes 16..31 accumulate columns nobody reads)
            }
        PLUME_TL(9);
            compute_sync();       // exch / x tile / staging are rewritten by the next tile
        }
#ifdef PLUME_TC_TIMELINE
        if (tl_on) {
            const long long end = clock64();
            printf("ppo_tc timeline (cycles): Ph0 gather+rstd %lld | G1 production %lld | G1 mma drain %lld | Ph3 LN2/loss %lld | "
                   "Ph4 %lld | G2 production %lld | G3 production %lld | Ph6 columns %lld | drain %lld | Ph6 scalars %lld | tail %lld | total %lld"
                   " || Ph3: stats %lld | heads %lld | loss %lld | loss barrier %lld | LN2-bwd means %lld"
                   " || Ph6: A0 %lld | B0 %lld | A1 %lld | B1 %lld || G1 (after Ph0): start c0 %lld, start c1 %lld, start c2 %lld, published c2 "
                   "%lld, c3: stage free %lld, start %lld\n",
                   tl_[1] - tl_[0], tl_[2] - tl_[1], tl_[3] - tl_[2], tl_[4] - tl_[3], tl_[5] - tl_[4], tl_[6] - tl_[5],
                   tl_[7] - tl_[6], tl_[8] - tl_[7], 0LL, tl_[9] - tl_[8], end - tl_[9], end - tl_[0],
                   tl_[10] - tl_[3], tl_[11] - tl_[10], tl_[12] - tl_[11], tl_[13] - tl_[12], tl_[4] - tl_[13],
                   tl_[14] - tl_[7], tl_[15] - tl_[14], tl_[16] - tl_[15], tl_[17] - tl_[16], tl_[18] - tl_[1], tl_[19] - tl_[1], tl_[20] - tl_[1],
                   tl_[22] - tl_[1], tl_[23] - tl_[1], tl_[21] - tl_[1]);
        }
#endif
    }

    // ---- flush ------------------------------------------------------------------------------------------
#ifdef PLUME_TC_TIMELINE
    if (kt_on) kt_[2] = clock64();
#endif
    float* g = a.grads;
    if (step > 0) {
        tc::mbar_wait(&g3done, (lt - 1u) & 1u);        // the last tile's G3
        tc::tc_fence_after();
        // dW2 accumulator: TMEM lane = output srow, this warp's column group = inputs [64 cg, 64 cg + 64)
#pragma unroll 1
        for (int q = 0; q < 256 / G / 32; ++q) {
            float vv[32];
            const int col = (256 / G) * cg + 32 * q;
            tc::tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(256 + col), vv);
            tc::tmem_ld_wait();
            float4* dst = reinterpret_cast<float4*>(g + PLUME_OFF_W2 + srow * 256 + col);
            const float un = 1.0f / dz_scale;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                atomicAdd(dst + i, make_float4(vv[4 * i] * un, vv[4 * i + 1] * un, vv[4 * i + 2] * un, vv[4 * i + 3] * un));
        }
        tc::tc_fence_before();
    }
    {
        const int o = r128;
        atomicAdd(g + PLUME_OFF_B2 + o, g_b2);
        atomicAdd(g + PLUME_OFF_G2 + o, g_g2);
        atomicAdd(g + PLUME_OFF_BE2 + o, g_be2);
        const float g_wh[6] = {g_wh2[0].x, g_wh2[0].y, g_wh2[1].x, g_wh2[1].y, g_wh2[2].x, g_wh2[2].y};
#pragma unroll
        for (int j = 0; j < 5; ++j) atomicAdd(g + PLUME_OFF_WA + j * 128 + o, g_wh[j]);
        atomicAdd(g + PLUME_OFF_WC + o, g_wh[5]);
    }
    // layer 1: combine the sample groups of P and the CTA's per-sample scalar sums
    compute_sync();
    float* pbuf = xh;                      // [256 inputs][8]
#if PLUME_TC_PM == 64
    if (lane < 16) {
        const int in = 64 * cg + 16 * wq + lane;
#else
    if (wq < 2) {                          // an M = 128 accumulator keeps row r in lane r
        const int in = 64 * cg + 32 * wq + lane;
#endif
        const float unp = 1.0f / (dz_scale * kW2BwdScale);
        *reinterpret_cast<float4*>(pbuf + in * 8) = make_float4(Pacc[0] * unp, Pacc[1] * unp, Pacc[2] * unp, Pacc[3] * unp);
        *reinterpret_cast<float4*>(pbuf + in * 8 + 4) = make_float4(Pacc[4] * unp, Pacc[5] * unp, Pacc[6] * unp, Pacc[7] * unp);
    }
    compute_sync();
    if (tid < 4) {                       // total = policy + value - beta * entropy (train_ppo2.0.py:82)
        const double lsum = tid == 0 ? cta_loss[1] + cta_loss[2] - (double)a.entropy_beta * cta_loss[3] : cta_loss[tid];
        atomicAdd(a.loss_out + tid, lsum * (double)a.inv_global);
    }
    if (tid < 256) {
        float S[41];
#pragma unroll
        for (int k = 0; k < 41; ++k) S[k] = cta_acc[k];
        if (tid < 5) atomicAdd(g + PLUME_OFF_BA + tid, S[35 + tid]);
        if (tid == 5) atomicAdd(g + PLUME_OFF_BC, S[40]);
        const int in = tid;                  // 256 threads = 256 layer-1 outputs
        float P[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) P[c] = pbuf[in * 8 + c];
        float w[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) w[k] = W1c[k * 256 + in];
        const float b1c = P1[in], g1 = P1[256 + in];
        // Q[k][k2] (symmetric) from the packed upper triangle
        float Q[6][6];
        {
            int qi = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k)
#pragma unroll
                for (int k2 = k; k2 < 6; ++k2) {
                    Q[k][k2] = S[14 + qi];
                    Q[k2][k] = S[14 + qi];
                    ++qi;
                }
        }
        float gg1 = b1c * P[6], wr = b1c * S[13];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            gg1 = fmaf(w[k], P[k], gg1);
            wr = fmaf(w[k], S[7 + k], wr);
        }
        atomicAdd(g + PLUME_OFF_BE1 + in, P[7]);
        atomicAdd(g + PLUME_OFF_G1 + in, gg1);
        atomicAdd(g + PLUME_OFF_B1 + in, g1 * P[6] - S[0] - wr);
#pragma unroll
        for (int k2 = 0; k2 < 6; ++k2) {
            float wq2 = b1c * S[7 + k2];
#pragma unroll
            for (int k = 0; k < 6; ++k) wq2 = fmaf(w[k], Q[k2][k], wq2);
            atomicAdd(g + PLUME_OFF_W1 + in * 6 + k2, g1 * P[k2] - S[1 + k2] - wq2);
        }
    }
#ifdef PLUME_TC_TIMELINE
    if (kt_on)
        printf("ppo_tc kernel timeline CTA %d (cycles): prologue %lld (barrier init %lld, TMEM alloc %lld, first barrier %lld, "
               "parameter staging %lld) | tiles %lld | flush %lld\n", (int)blockIdx.x, kt_[1] - kt_[0], kp_[0] - kt_[0],
               kp_[1] - kp_[0], kp_[2] - kp_[1], kt_[1] - kp_[2], kt_[2] - kt_[1], clock64() - kt_[2]);
#endif
    }   // compute warps
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// ---- sample records ---------------------------------------------------------------------------------------
// [M][12] = {obs 0..5, advantage, return, old value, old log-prob, action (int bits), 0}: the layout of the pf
// staging rows, so a gathered record needs no rearrangement.  One thread per record, three float4 stores.
__global__ void ppo_pack_kernel(plume_ppo_batch b, float* __restrict__ packed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.total) return;
    const float2* o2 = reinterpret_cast<const float2*>(b.obs + i * 6);         // 24-byte rows: 8-byte aligned
    const float2 o01 = o2[0], o23 = o2[1], o45 = o2[2];
    float4* dst = reinterpret_cast<float4*>(packed + i * 12);
    dst[0] = make_float4(o01.x, o01.y, o23.x, o23.y);
    dst[1] = make_float4(o45.x, o45.y, b.advantages[i], b.returns[i]);
    dst[2] = make_float4(b.old_values[i], b.old_log_probs[i], __int_as_float(b.actions[i]), 0.0f);
}

int launch_ppo_pack(const plume_ppo_batch& b, float* packed, cudaStream_t s) {
    if (b.total <= 0) return 0;
    ppo_pack_kernel<<<(unsigned)((b.total + 255) / 256), 256, 0, s>>>(b, packed);
    if (cudaGetLastError() != cudaSuccess) return fail("ppo_pack_kernel launch failed");
    return 0;
}

// ---- launch ---------------------------------------------------------------------------------------------
int64_t ppo_tc_workspace_bytes() { return (int64_t)kWsFloats * (int64_t)sizeof(float) + 256; }

// The stash as a 5-D tensor of 32-bit words for the per-chunk tensor store: (256 words = one 1 KB piece: 8 samples x 64 inputs
// of hi or lo) x (q: chunk of the pair, 1 KB on) x (16 groups of 8 samples, 2 KB on) x (hi / lo, 32 KB on) x (pair, 64 KB on).
// The box {256, 1, 16, 2, 1} is the 32 KB of an A stage in shared-memory order.
static int make_stash_map(CUtensorMap* map, void* stash, int pairs) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess)
            return fail("ppo_tc: cuTensorMapEncodeTiled is not available from this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[5] = {256, 2, 16, 2, (cuuint64_t)pairs};
    const cuuint64_t strides[4] = {1024, 2048, 32768, 65536};            // bytes, dimensions 1..4
    const cuuint32_t box[5] = {256, 1, 16, 2, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, stash, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("ppo_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int launch_ppo_tc(const float* params, const PpoArgs& a, void* workspace, cudaStream_t s, bool zero_grads) {
    static bool configured = false;
    const int smem = TcSmem::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(ppo_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("ppo_tc_kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    float* w2s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    ppo_tc_prep_kernel<<<kPrepSplitBlocks + 1, 256, 0, s>>>(params, w2s, zero_grads ? a.grads : nullptr);
    if (cudaGetLastError() != cudaSuccess) return fail("ppo_tc_prep_kernel launch failed");
    const long long tiles = (a.mb_size + kTcTile - 1) / kTcTile;
    int grid = sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (grid > kTcMaxCtas) grid = kTcMaxCtas;       // the workspace holds one activation stash per CTA
    if (tiles < grid) grid = (int)tiles;
    // dz2 ~ O(1..100) / global batch: a power of two 16x below the batch size brings it to O(0.1..10), four orders of
    // magnitude under fp16's largest value
    const float dz_scale = exp2f(floorf(log2f(1.0f / a.inv_global)) - 4.0f);
    // (one map per workspace address; rebuilt when the caller comes with another buffer)
    static CUtensorMap stash_map;
    static void* mapped = nullptr;
    if (mapped != (void*)(w2s + kWsStash)) {
        if (make_stash_map(&stash_map, w2s + kWsStash, 2 * kTcMaxCtas)) return -1;
        mapped = (void*)(w2s + kWsStash);
    }
    ppo_tc_kernel<<<grid, kTcLaunchThreads, smem, s>>>(params, a, w2s, dz_scale, stash_map);
    if (cudaGetLastError() != cudaSuccess) return fail("ppo_tc_kernel launch failed");
    return 0;
}

}  // namespace plume
