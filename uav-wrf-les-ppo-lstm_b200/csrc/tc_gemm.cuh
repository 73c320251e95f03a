// tc_gemm.cuh -- sm_100a tensor-core building blocks: tcgen05.mma (kind::tf32) with operands in
// shared memory and fp32 accumulators in TMEM, used with the 3xTF32 split so that the GEMM-shaped
// parts of the PPO update keep fp32-grade accuracy (parity bar: fp32 rel 1e-5):
//
//      a = a_hi + a_lo,  a_hi = a rounded to nearest TF32, a_lo = the remainder rounded to nearest TF32
//      A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi          (3 MMAs per K-step, error ~2^-21 relative)
//
// Operand tiles are written to shared memory by the CTA's own threads (they have to be split
// anyway) in the canonical K-major no-swizzle ("interleaved") UMMA layout: 8-row x 16-byte core
// matrices, stored [row/8][k/4][row%8][k%4]; LBO = 128 B (next core matrix along K), SBO =
// (KC/4)*128 B (next 8 rows).  For a [ROWS][32] fp32 chunk the 16-byte slot index is simply
// f = (row/8)*64 + (k/4)*8 + row%8, so consecutive threads store consecutive slots (conflict free)
// while reading 64 contiguous bytes per row from global memory.
//
// Descriptor formats follow the PTX ISA / CUTLASS cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor); the guide is /opt/skills/guides/blackwell_cuda_programming.md.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace plume {
namespace tc {

constexpr int kChunkK = 32;                      // fp32 elements per K chunk (128 B per row)
constexpr uint32_t kLBO = 128;                   // bytes between K-adjacent core matrices
constexpr uint32_t kSBO = (kChunkK / 4) * 128;   // bytes between 8-row groups (1024)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

// the same with a back-off between polls: sixteen warps spinning on try_wait take the issue slots of the one thread that
// issues the MMAs they are waiting for
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// Wait of a WHOLE warp whose control flow has to stay warp-uniform (the MMA-issuing warps: only then does the compiler keep
// descriptors and addresses in uniform registers): every lane polls, the vote makes the loop condition uniform.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, uint32_t ns) {
    while (!__any_sync(0xffffffffu, mbar_try_wait(bar, parity))) __nanosleep(ns);
}

// ---- proxy / tcgen05 fences -----------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread i of the warp = lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 8 consecutive columns -> 8 registers per thread
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------
// shared-memory matrix descriptor: K-major, no swizzle, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T, one K-step of 8 tf32; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread -> one arrive on `bar` when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- bulk copies (TMA, no tensor map): one thread moves a contiguous block; sizes and addresses multiples of 16 bytes ----
// shared -> global, tracked by the thread's bulk groups
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the sources of all committed groups have been read (the shared-memory buffers may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... of all committed groups but the most recent one
__device__ __forceinline__ void bulk_wait_group_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all committed groups are complete (their global-memory writes included)
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// global -> shared, completion (byte count) signalled on an mbarrier of count 1
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t mb = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(mb) : "memory");
}

// shared -> global through a 5-D tensor map (box = the contiguous shared-memory source), tracked by the thread's bulk groups
__device__ __forceinline__ void tensor_store_5d(const void* tensor_map, const void* smem_src, int c0, int c1, int c2, int c3,
                                                int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                 ::"l"(reinterpret_cast<uint64_t>(tensor_map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(smem_src))
                 : "memory");
}
// two copies of `bytes` each on one phase of the mbarrier
__device__ __forceinline__ void bulk_load2(void* dst0, const void* src0, void* dst1, const void* src1, uint32_t bytes,
                                           uint64_t* bar) {
    const uint32_t mb = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(2u * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst0)), "l"(src0), "r"(bytes), "r"(mb) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst1)), "l"(src1), "r"(bytes), "r"(mb) : "memory");
}

// ---- operand staging -----------------------------------------------------------------------
// hi = RN_tf32(x) (ties away from zero, like cvt.rna.tf32.f32, which ptxas expands into 4 instructions
// per value because of its Inf/NaN handling -- ncu r1b: 13 % of the update kernel's instructions).  Done
// on the bit pattern: +0x1000 rounds the magnitude, the mask clears the 13 low bits.  lo = x - hi is
// exact; the tensor core ignores the 13 low bits of an operand (it truncates), so adding 0x1000 to lo's
// bit pattern is all that is needed to make that truncation a round-to-nearest.
// |x - hi - lo_as_seen_by_the_tensor_core| <= 2^-23 |x|.  (Inf/NaN inputs stay non-finite.)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}

// Cooperative load of a [ROWS][32] fp32 chunk (row-major, leading dimension ld floats, rows
// beyond `valid_rows` read as zero) into the hi/lo operand buffers.  kThreads threads.
template <int ROWS, int kThreads>
__device__ __forceinline__ void load_split_chunk(float* __restrict__ s_hi, float* __restrict__ s_lo,
                                                 const float* __restrict__ src, size_t ld, int valid_rows, int tid) {
    constexpr int kSlots = ROWS * 8;
#pragma unroll
    for (int f = tid; f < kSlots; f += kThreads) {
        const int row = (f >> 6) * 8 + (f & 7), k4 = (f & 63) >> 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < valid_rows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * ld + k4 * 4));
        float4 h, l;
        split_tf32(v.x, h.x, l.x);
        split_tf32(v.y, h.y, l.y);
        split_tf32(v.z, h.z, l.z);
        split_tf32(v.w, h.w, l.w);
        reinterpret_cast<float4*>(s_hi)[f] = h;
        reinterpret_cast<float4*>(s_lo)[f] = l;
    }
}

// element (row, k) of a chunk buffer (for producers that write operands from registers)
__device__ __forceinline__ int chunk_offset(int row, int k) {
    return ((row >> 3) * 64 + (k >> 2) * 8 + (row & 7)) * 4 + (k & 3);
}

// The 12 MMAs of one 32-wide K chunk: M=128 rows of A, N rows of B; issued by one thread.
__device__ __forceinline__ void mma_chunk_3xtf32(uint32_t tmem_d, const float* a_hi, const float* a_lo,
                                                 const float* b_hi, const float* b_lo, uint32_t idesc,
                                                 bool first_chunk) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
    for (int j = 0; j < kChunkK / 8; ++j) {
        const uint32_t off = j * 2 * kLBO;       // 8 tf32 = two 16-byte core-matrix columns
        const uint64_t dah = make_smem_desc(ah + off, kLBO, kSBO), dal = make_smem_desc(al + off, kLBO, kSBO);
        const uint64_t dbh = make_smem_desc(bh + off, kLBO, kSBO), dbl = make_smem_desc(bl + off, kLBO, kSBO);
        mma_tf32(tmem_d, dal, dbh, idesc, (first_chunk && j == 0) ? 0u : 1u);   // small terms first
        mma_tf32(tmem_d, dah, dbl, idesc, 1u);
        mma_tf32(tmem_d, dah, dbh, idesc, 1u);
    }
}

// Same 12 MMAs, but the two small cross terms go to their own accumulator `tmem_small`.  The tensor core
// truncates when it adds into an accumulator; keeping the 2^-11-times smaller terms out of the main chain
// leaves one truncation per K-step on `tmem_big` instead of three (the caller adds the two at the end).
__device__ __forceinline__ void mma_chunk_3xtf32_split(uint32_t tmem_big, uint32_t tmem_small, const float* a_hi,
                                                       const float* a_lo, const float* b_hi, const float* b_lo,
                                                       uint32_t idesc, bool first_chunk) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
    for (int j = 0; j < kChunkK / 8; ++j) {
        const uint32_t off = j * 2 * kLBO;
        const uint64_t dah = make_smem_desc(ah + off, kLBO, kSBO), dal = make_smem_desc(al + off, kLBO, kSBO);
        const uint64_t dbh = make_smem_desc(bh + off, kLBO, kSBO), dbl = make_smem_desc(bl + off, kLBO, kSBO);
        const uint32_t acc = (first_chunk && j == 0) ? 0u : 1u;
        mma_tf32(tmem_small, dal, dbh, idesc, acc);
        mma_tf32(tmem_small, dah, dbl, idesc, 1u);
        mma_tf32(tmem_big, dah, dbh, idesc, acc);
    }
}

// ---- kind::f16 with the two-term fp16 split ----------------------------------------------------------------
// x = hi + lo / s,  hi = fp16(x),  lo = fp16((x - hi) s)   (s = 2^11 keeps lo in the normal fp16 range: |x - hi| <=
// 2^-11 |x|; s = 1 where the cross terms must share the accumulator of the main term).  A 64-element fp16 chunk has
// the same 128 bytes per row as the 32-element tf32 chunk: the slot layout f = (row/8)*64 + (k/8)*8 + row%8, the
// shared-memory descriptors (LBO 128, SBO 1024) and the 4 x 3 MMAs per chunk are unchanged -- each MMA now covers
// K = 16 at twice the MAC rate, so a chunk step carries twice the K in the same tensor-pipe time and half the
// operand bytes per element.
constexpr int kChunkKH = 64;
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {      // D fp32, A/B fp16, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// The same with A read MN-major (instruction-descriptor bit 15): a 16-byte slot of A then holds 8 consecutive M for ONE
// k, the eight k of a core matrix are consecutive slots; in the shared-memory descriptor the leading byte offset is the
// distance between core matrices along K, the stride byte offset the distance along M (CUTLASS
// cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>, LayoutType::INTERLEAVE).  This is what lets ONE copy of
// dz2 serve as A of both backward GEMMs of the update: K-major along the samples for G3 (dW2 = dz2^T h1), MN-major
// along the outputs for G2 (dh1 = dz2 W2).
__host__ __device__ constexpr uint32_t make_idesc_f16_a_mn(int M, int N) { return make_idesc_f16(M, N) | (1u << 15); }

// B read MN-major (bit 16): a 16-byte slot of B holds 8 consecutive N for ONE k -- the form in which the update kernel's G3
// reads a stashed forward-activation chunk (written K-major as A of G1: the same bytes).
__host__ __device__ constexpr uint32_t make_idesc_f16_b_mn(int M, int N) { return make_idesc_f16(M, N) | (1u << 16); }

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ONE rolled issue loop for every hi/lo GEMM step of the update kernel: n K-steps of (lo.hi, hi.lo, hi.hi) into one
// accumulator, the four descriptors advancing by a fixed number of bytes per K-step (profiles/debug/mma/umma_rate.cu measures
// what an instruction costs back to back: 64 cycles for M128 x N128 x K16 with this hi / lo pattern).
struct MmaOperands {
    uint64_t ah, al, bh, bl;          // descriptors of the first K-step
    uint32_t a_adv, b_adv;            // advance per K-step in 16-byte units (added to the descriptors' start-address field)
};
__device__ __forceinline__ MmaOperands make_operands(uint32_t a_hi, uint32_t a_lo, uint32_t a_adv_bytes, uint32_t a_lbo,
                                                     uint32_t a_sbo, uint32_t b_hi, uint32_t b_lo, uint32_t b_adv_bytes,
                                                     uint32_t b_lbo, uint32_t b_sbo) {
    MmaOperands o;
    o.ah = make_smem_desc(a_hi, a_lbo, a_sbo);
    o.al = make_smem_desc(a_lo, a_lbo, a_sbo);
    o.bh = make_smem_desc(b_hi, b_lbo, b_sbo);
    o.bl = make_smem_desc(b_lo, b_lbo, b_sbo);
    o.a_adv = a_adv_bytes >> 4;
    o.b_adv = b_adv_bytes >> 4;
    return o;
}
// Called by ALL lanes of the issuing warp with warp-uniform arguments; lane `leader` issues.  (Inside a one-lane branch the
// compiler keeps the descriptors in vector registers and wraps every tcgen05.mma in an ELECT + 7 x R2UR + BRA.U.ANY loop:
// ~100 cycles of issue latency per instruction -- measured 105-140 cycles per MMA in the kernel against 64 back to back.)
__device__ __forceinline__ void issue_split_steps(bool leader, uint32_t tmem_d, MmaOperands o, int n, uint32_t idesc,
                                                      uint32_t first_acc) {
    uint32_t acc = first_acc;
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        if (leader) {
            mma_f16(tmem_d, o.al, o.bh, idesc, acc);
            mma_f16(tmem_d, o.ah, o.bl, idesc, 1u);
            mma_f16(tmem_d, o.ah, o.bh, idesc, 1u);
        }
        acc = 1u;
        o.ah += o.a_adv;              // (shared-memory addresses >> 4 stay inside the 14-bit field: no carry into the LBO bits)
        o.al += o.a_adv;
        o.bh += o.b_adv;
        o.bl += o.b_adv;
    }
}

// two values -> packed fp16 hi and fp16 lo (scaled by s)
__device__ __forceinline__ void split_f16x2(float a, float b, float s, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    // remainder and scaling as packed fp32 pairs (FADD2 / FMUL2: same IEEE results, half the issue slots)
    const float2 r = __fmul2_rn(__fadd2_rn(make_float2(a, b), make_float2(-hf.x, -hf.y)), make_float2(s, s));
    const __half2 l = __floats2half2_rn(r.x, r.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// eight consecutive K values -> one 16-byte operand slot of each buffer
__device__ __forceinline__ void split_f16x8(const float4 v0, const float4 v1, float s, uint4& hi, uint4& lo) {
    split_f16x2(v0.x, v0.y, s, hi.x, lo.x);
    split_f16x2(v0.z, v0.w, s, hi.y, lo.y);
    split_f16x2(v1.x, v1.y, s, hi.z, lo.z);
    split_f16x2(v1.z, v1.w, s, hi.w, lo.w);
}

// Cooperative load of a [ROWS][64] fp32 chunk (row-major, leading dimension ld) into fp16 hi / lo operand buffers.
template <int ROWS, int kThreads>
__device__ __forceinline__ void load_split_chunk_f16(float* __restrict__ s_hi, float* __restrict__ s_lo,
                                                     const float* __restrict__ src, size_t ld, int valid_rows, int tid,
                                                     float lo_scale) {
    constexpr int kSlots = ROWS * 8;
#pragma unroll
    for (int f = tid; f < kSlots; f += kThreads) {
        const int row = (f >> 6) * 8 + (f & 7), k8 = (f & 63) >> 3;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (row < valid_rows) {
            v0 = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * ld + k8 * 8));
            v1 = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * ld + k8 * 8 + 4));
        }
        uint4 h, l;
        split_f16x8(v0, v1, lo_scale, h, l);
        reinterpret_cast<uint4*>(s_hi)[f] = h;
        reinterpret_cast<uint4*>(s_lo)[f] = l;
    }
}

// The 12 MMAs of one 64-wide fp16 chunk, all into one accumulator (unscaled lo).
__device__ __forceinline__ void mma_chunk_f16(uint32_t tmem_d, const float* a_hi, const float* a_lo, const float* b_hi,
                                              const float* b_lo, uint32_t idesc, bool first_chunk) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
    for (int j = 0; j < kChunkKH / 16; ++j) {
        const uint32_t off = j * 2 * kLBO;       // 16 halves = two 16-byte core-matrix columns
        const uint64_t dah = make_smem_desc(ah + off, kLBO, kSBO), dal = make_smem_desc(al + off, kLBO, kSBO);
        const uint64_t dbh = make_smem_desc(bh + off, kLBO, kSBO), dbl = make_smem_desc(bl + off, kLBO, kSBO);
        mma_f16(tmem_d, dal, dbh, idesc, (first_chunk && j == 0) ? 0u : 1u);
        mma_f16(tmem_d, dah, dbl, idesc, 1u);
        mma_f16(tmem_d, dah, dbh, idesc, 1u);
    }
}

// Same, the cross terms (lo scaled by 2^11) in their own accumulator: result = big + small * 2^-11.
__device__ __forceinline__ void mma_chunk_f16_split(uint32_t tmem_big, uint32_t tmem_small, const float* a_hi,
                                                    const float* a_lo, const float* b_hi, const float* b_lo,
                                                    uint32_t idesc, bool first_chunk) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
    for (int j = 0; j < kChunkKH / 16; ++j) {
        const uint32_t off = j * 2 * kLBO;
        const uint64_t dah = make_smem_desc(ah + off, kLBO, kSBO), dal = make_smem_desc(al + off, kLBO, kSBO);
        const uint64_t dbh = make_smem_desc(bh + off, kLBO, kSBO), dbl = make_smem_desc(bl + off, kLBO, kSBO);
        const uint32_t acc = (first_chunk && j == 0) ? 0u : 1u;
        mma_f16(tmem_small, dal, dbh, idesc, acc);
        mma_f16(tmem_small, dah, dbl, idesc, 1u);
        mma_f16(tmem_big, dah, dbh, idesc, acc);
    }
}

}  // namespace tc
}  // namespace plume
