// learner_kernels.cu -- K5 GAE reverse scan + normalisation (train_ppo2.0.py:17-39), K7
// clip_grad_norm + Adam (train_ppo2.0.py:86-87,113), the stateless minibatch permutation
// (replaces torch.randperm, :43) and K8 the curriculum (PPOTrainer.update, model.py:188-221).
//
// Algorithmic bytes (DESIGN.md "K5"): scan reads r,V,done (12 B) and writes A (4 B); the
// normalise pass reads A,V (8 B) and writes A,ret (8 B) => 32 B per transition.
#include "common.cuh"
#include "curriculum.cuh"

namespace plume {

__device__ __forceinline__ double block_sum(double v, double* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) t += scratch[w];
    return t;
}

// ---- K5: one thread per env column, T sequential steps, coalesced across envs ----------------
__global__ void __launch_bounds__(128) gae_scan_kernel(const float* __restrict__ rewards,
                                                       const float* __restrict__ values,
                                                       const float* __restrict__ dones, int T, int N, float gamma,
                                                       float gamma_lam, float* __restrict__ adv,
                                                       double* __restrict__ stats, int variant,
                                                       const float* __restrict__ last_values) {
    __shared__ double scratch[4];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    if (n < N) {
        float last = 0.0f;
        float v_next = 0.0f, d_next = 0.0f;
        // The recurrence is serial in t, the loads are not: fetch kGaeBatch rows ahead so that 3 x kGaeBatch
        // independent loads are in flight per thread (4096 threads cannot hide HBM latency otherwise).
        constexpr int kGaeBatch = 8;
        for (int t1 = T - 1; t1 >= 0; t1 -= kGaeBatch) {
            float rr[kGaeBatch], vv[kGaeBatch], dd[kGaeBatch];
#pragma unroll
            for (int q = 0; q < kGaeBatch; ++q) {
                const int t = t1 - q;
                if (t >= 0) {
                    const size_t i = (size_t)t * N + n;
                    rr[q] = __ldg(rewards + i);
                    vv[q] = __ldg(values + i);
                    dd[q] = __ldg(dones + i);
                }
            }
#pragma unroll
            for (int q = 0; q < kGaeBatch; ++q) {
                const int t = t1 - q;
                if (t < 0) break;
                const size_t i = (size_t)t * N + n;
                const float r = rr[q], v = vv[q], d = dd[q];
                float a;
                if (variant == PLUME_GAE_QUIRK) {
                    float nnt, nv;
                    if (t == T - 1) {                      // :22-24 self-bootstrap
                        nnt = __fsub_rn(1.0f, d);
                        nv = __fmul_rn(v, nnt);
                    } else {                               // :26-27 masks with dones[t+1]
                        nnt = __fsub_rn(1.0f, d_next);
                        nv = __fmul_rn(v_next, nnt);
                    }
                    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, nv)), v);          // :29
                    a = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lam, nnt), last));              // :30
                } else if (variant == PLUME_GAE_BOOTSTRAP) {   // PPOV1.1/train_ppo1.0.py:75-85
                    const float nnt = __fsub_rn(1.0f, t == T - 1 ? d : d_next);
                    const float nv = (t == T - 1) ? last_values[n] : v_next;
                    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, nv), nnt)), v);
                    a = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lam, nnt), last));
                } else {                                       // PPOV1.2: masks with dones[t], no bootstrap (:368-376)
                    const float nt = __fsub_rn(1.0f, d);
                    const float nv = (t == T - 1) ? 0.0f : __fmul_rn(v_next, nt);
                    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(gamma, nv)), v);
                    a = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lam, last), nt));
                }
                adv[i] = a;
                last = a;
                v_next = v;
                d_next = d;
                s1 += (double)a;
                s2 += (double)a * (double)a;
            }
        }
    }
    const double b1 = block_sum(s1, scratch);
    const double b2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
        atomicAdd(stats + 0, b1);
        atomicAdd(stats + 1, b2);
        int cnt = N - blockIdx.x * blockDim.x;
        cnt = cnt > (int)blockDim.x ? (int)blockDim.x : cnt;
        atomicAdd(stats + 2, (double)cnt * (double)T);
    }
}

__global__ void __launch_bounds__(256) gae_normalise_kernel(float* __restrict__ adv, const float* __restrict__ values,
                                                            long long count, const double* __restrict__ stats,
                                                            float* __restrict__ returns, int variant) {
    const double cnt = stats[2];
    const double mean = stats[0] / cnt;
    double var = (stats[1] - cnt * mean * mean) / (cnt - 1.0);       // unbiased, torch .std()
    var = var < 0.0 ? 0.0 : var;
    double sd = sqrt(var);
    if (variant == PLUME_GAE_QUIRK && !(sd >= 1e-6)) sd = 1.0;         // :36-37 (also catches NaN)
    const float mean32 = (float)mean;
    // :38 divides by std + 1e-6; the older drivers by std + 1e-8 (train_ppo1.0.py:89, ppo注释版.py:379)
    const float denom = __fadd_rn((float)sd, variant == PLUME_GAE_QUIRK ? 1e-6f : 1e-8f);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float raw = adv[i];
        const float a = __fdiv_rn(__fsub_rn(raw, mean32), denom);
        adv[i] = a;
        // :39 (sic) returns = NORMALISED advantage + value; train_ppo1.0.py:86 adds the raw advantage
        returns[i] = __fadd_rn(variant == PLUME_GAE_BOOTSTRAP ? raw : a, values[i]);
    }
}

// ---- K7: global-norm clip + Adam over the flat parameter buffer ------------------------------------
// Every CTA recomputes the global gradient norm from L2 (145 KB, same summation order in every CTA, so
// the value is bitwise identical and no grid-wide synchronisation is needed) and then updates its own
// 1024-element slice.
__global__ void __launch_bounds__(1024) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, int n,
                                                         float max_norm, float lr, float b1, float b2, float eps,
                                                         float bc1, float bc2_sqrt, float* grad_norm_out) {
    __shared__ double scratch[32];
    double ss = 0.0;
    const int n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 x = __ldg(g4 + i);
        ss += (double)x.x * (double)x.x + (double)x.y * (double)x.y + (double)x.z * (double)x.z +
              (double)x.w * (double)x.w;
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) ss += (double)g[i] * (double)g[i];
    const float norm = (float)sqrt(block_sum(ss, scratch));
    if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
    float coef = max_norm / (norm + 1e-6f);       // torch.nn.utils.clip_grad_norm_
    coef = coef > 1.0f ? 1.0f : coef;
    const float step_size = lr / bc1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float gi = g[i] * coef;
        const float mi = m[i] + (gi - m[i]) * (1.0f - b1);            // exp_avg.lerp_(grad, 1-beta1)
        const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;           // mul_(beta2).addcmul_(g, g, 1-beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

__global__ void permutation_kernel(long long total, unsigned long long seed, int epoch, long long start,
                                   long long count, long long* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = (long long)feistel_permute((uint64_t)(start + i), (uint64_t)total, seed, (uint32_t)epoch);
}

// ---- K8: curriculum over the finished episodes of a [T][N] segment, canonical order ----------

// PPOTrainer.update (model.py:188-221) for the `total_eps` episodes of a segment whose successes have been binned
// per curriculum window (blk_succ[b] = successes among episodes [b*window, (b+1)*window) counted from the start
// of the carried partial window).  One thread.
__device__ void curriculum_apply(const int* blk_succ, bool overflow, long long total_eps, double* state,
                                 double* curriculum, double initial_radius, double min_radius, double radius_decay,
                                 double thr, int window, double decay_factor, double* win_out) {
    if (overflow) {
        state[4] = -1.0;   // host raises
        return;
    }
    const long long hist_len = (long long)state[4];
    double radius = state[0], eb = state[1];
    double env_radius = state[2], env_eb = state[3];
    long long succ_total = 0;
    const long long n_done = hist_len + total_eps;
    const long long full = n_done / window;
    if (win_out) {
        win_out[0] = (double)hist_len;
        win_out[1] = (double)((full < kCurMaxBlocks ? full : kCurMaxBlocks - 1) + 1);
    }
    for (long long b = 0; b <= full && b < kCurMaxBlocks; ++b) {
        long long s = blk_succ[b];
        succ_total += s;
        if (b == 0) s += (long long)state[5];
        if (win_out) win_out[2 + b] = radius;       // what train_ppo2.0.py:247 logs for the episodes of window b
        if (b < full) {
            // the env sees the trainer's values from before this update (model.py:189-190)
            env_radius = radius;
            env_eb = eb;
            const double rate = (double)s / (double)window;
            eb *= pow(decay_factor, 1.0 + rate);                         // :197-199
            eb = eb > 0.1 ? eb : 0.1;                                    // :201
            if (rate > thr) {                                            // :205-209
                const double r2 = radius * pow(radius_decay, 2.0 + 3.0 * (rate - thr));
                radius = r2 > min_radius ? r2 : min_radius;
            } else if (rate < 0.25) {                                    // :210-214
                const double r2 = radius * 1.1;
                radius = r2 < initial_radius ? r2 : initial_radius;
            }
            if (fabs(radius - env_radius) > 5.0)                         // :217-218
                radius = env_radius + 5.0 * ((radius > env_radius) - (radius < env_radius));
        } else {
            state[5] = (double)s;                                        // partial window carried over
        }
    }
    (void)env_eb;
    state[0] = radius;
    state[1] = eb;
    state[2] = radius;      // envs latch the trainer's current values at their next reset
    state[3] = eb;
    state[4] = (double)(n_done % window);
    state[6] += (double)total_eps;
    state[7] += (double)succ_total;
    curriculum[0] = radius;
    curriculum[1] = eb;
}

// Where the (done, reached) flags of canonical position p come from: per-warp cursors that walk 32 positions
// at a time (seek once per warp, no per-element division).
struct FlagsSeparate {            // float dones + uint8 reached, one rank: position = index
    const float* dones;
    const uint8_t* reached;
    long long pos;
    __device__ __forceinline__ void seek(long long p) { pos = p; }
    __device__ __forceinline__ unsigned get(int lane) const {
        const unsigned d = dones[pos + lane] != 0.0f;
        return d ? (d | ((reached[pos + lane] != 0) << 1)) : 0u;
    }
    __device__ __forceinline__ void next() { pos += 32; }
};
struct FlagsPacked {              // uint8 code (bit 0 done, bit 1 reached), one [T][N] array per rank; canonical
    CodeSrc code;                 // order = step-major, then GLOBAL env id = rank * N + local id
    int T, N, world;
    int t, r, n;                  // cursor: position of lane 0
    __device__ __forceinline__ void seek(long long p) {
        const long long row = (long long)world * N;
        const long long tt = p / row, g = p - tt * row;
        t = (int)tt;
        r = (int)(g / N);
        n = (int)(g - (long long)r * N);
    }
    __device__ __forceinline__ unsigned get(int lane) const {
        int tt = t, rr = r, nn = n + lane;
        while (nn >= N) {          // at most once when N >= 32 (never when N % 32 == 0)
            nn -= N;
            if (++rr == world) {
                rr = 0;
                ++tt;
            }
        }
        return code.base[rr][(size_t)tt * N + nn];
    }
    __device__ __forceinline__ void next() {
        n += 32;
        while (n >= N) {
            n -= N;
            if (++r == world) {
                r = 0;
                ++t;
            }
        }
    }
};

template <typename Flags>
__global__ void __launch_bounds__(1024) curriculum_kernel(Flags flags, long long total,
                                                          double* state, double* curriculum, double initial_radius,
                                                          double min_radius, double radius_decay, double thr,
                                                          int window, double decay_factor, double* win_out,
                                                          const uint32_t* comm_error) {
    // One CTA (the episode order is a serial dependency), 32 warps; warp w owns the contiguous range
    // [w*chunk, (w+1)*chunk) and walks it 32 flags at a time with coalesced loads + ballots.
    __shared__ int ep_cnt[32];
    __shared__ int blk_succ[kCurMaxBlocks];
    __shared__ int overflow_any;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) overflow_any = 0;
    const long long chunk = (((total + 31) / 32) + 31) / 32 * 32;     // multiple of 32 flags per warp
    const long long lo = chunk * warp, hi = (lo + chunk < total) ? lo + chunk : total;
    int eps = 0;
    if (lo < hi) flags.seek(lo);
#pragma unroll 4
    for (long long i = lo; i < hi; i += 32) {
        const bool d = (i + lane < hi) && (flags.get(lane) & 1u);
        eps += __popc(__ballot_sync(0xffffffffu, d));
        flags.next();
    }
    if (lane == 0) ep_cnt[warp] = eps;
    for (int i = tid; i < kCurMaxBlocks; i += blockDim.x) blk_succ[i] = 0;
    __syncthreads();
    const long long hist_len = (long long)state[4];
    long long ord = hist_len;                                          // ordinal of this warp's first episode
    for (int w = 0; w < warp; ++w) ord += ep_cnt[w];
    bool overflow = false;
    if (lo < hi) flags.seek(lo);
    for (long long i = lo; i < hi; i += 32) {
        const unsigned f = (i + lane < hi) ? flags.get(lane) : 0u;
        const bool d = f & 1u;
        const unsigned mask = __ballot_sync(0xffffffffu, d);
        if (d) {
            const long long b = (ord + __popc(mask & ((1u << lane) - 1u))) / window;
            if (b >= kCurMaxBlocks) overflow = true;
            else if (f & 2u) atomicAdd(&blk_succ[(int)b], 1);
        }
        ord += __popc(mask);
        flags.next();
    }
    if (overflow) atomicOr(&overflow_any, 1);
    __syncthreads();
    if (tid == (int)blockDim.x - 1 && !(comm_error && *comm_error))
        curriculum_apply(blk_succ, overflow_any != 0, ord - hist_len, state, curriculum, initial_radius, min_radius,
                         radius_decay, thr, window, decay_factor, win_out);
}


// ---- K8, scalable form for the packed [world][T][N] flags (N % 512 == 0): four small kernels ---------------
// count (many warps, 16 flags per lane and load) -> exclusive scan over the chunks (one CTA) -> bin the
// successes per curriculum window (many warps, global atomics on the few done&reached flags) -> apply the
// rule (one thread).  Independent of the world size up to the all-gather itself.
constexpr int kCurChunk = 2048;                 // canonical positions per warp (4 loads of 512)
constexpr int kCurMaxChunks = 1 << 16;

__device__ __forceinline__ const uint4* packed_chunk_ptr(const CodeSrc& code, int N, int world, long long p) {
    const long long row = (long long)world * N;
    const long long t = p / row, g = p - t * row;
    const long long r = g / N, n = g - r * N;
    return reinterpret_cast<const uint4*>(code.base[r] + (size_t)t * N + n);
}
__device__ __forceinline__ int count_done16(const uint4 v) {
    return __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
           __popc(v.w & 0x01010101u);
}

__global__ void __launch_bounds__(256) cur_count_kernel(const CodeSrc code, int N, int world,
                                                        long long total, int* __restrict__ chunk_cnt, int n_chunks) {
    const int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    int c = 0;
#pragma unroll
    for (int q = 0; q < kCurChunk / 512; ++q) {
        const long long p = (long long)chunk * kCurChunk + q * 512;
        if (p < total) c += count_done16(__ldg(packed_chunk_ptr(code, N, world, p) + lane));
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) chunk_cnt[chunk] = c;
}

__global__ void __launch_bounds__(1024) cur_scan_kernel(int* __restrict__ chunk_cnt, int n_chunks, int* __restrict__ bins,
                                                        int* __restrict__ total_out) {
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int per = (n_chunks + 1023) / 1024;
    int s = 0;
    for (int i = tid * per; i < (tid + 1) * per && i < n_chunks; ++i) s += chunk_cnt[i];
    part[tid] = s;
    for (int i = tid; i < kCurMaxBlocks; i += 1024) bins[i] = 0;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < 1024; ++i) {
            const int c = part[i];
            part[i] = run;
            run += c;
        }
        total_out[0] = run;
        total_out[1] = 0;        // overflow flag
    }
    __syncthreads();
    int run = part[tid];
    for (int i = tid * per; i < (tid + 1) * per && i < n_chunks; ++i) {
        const int c = chunk_cnt[i];
        chunk_cnt[i] = run;      // exclusive prefix
        run += c;
    }
}

__global__ void __launch_bounds__(256) cur_bin_kernel(const CodeSrc code, int N, int world,
                                                      long long total, const int* __restrict__ chunk_base, int n_chunks,
                                                      const double* __restrict__ state, int window,
                                                      int* __restrict__ bins, int* __restrict__ total_out) {
    const int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (chunk >= n_chunks) return;
    long long ord = (long long)state[4] + chunk_base[chunk];     // ordinal of the chunk's first episode
    for (int q = 0; q < kCurChunk / 512; ++q) {
        const long long p = (long long)chunk * kCurChunk + q * 512;
        if (p >= total) break;
        const uint4 v = __ldg(packed_chunk_ptr(code, N, world, p) + lane);
        const int c = count_done16(v);
        int incl = c;                                            // inclusive warp scan of the per-lane counts
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (c) {
            long long mine = ord + incl - c;
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned f = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                if (f & 1u) {
                    const long long b = mine / window;
                    if (b >= kCurMaxBlocks) atomicExch(total_out + 1, 1);
                    else if (f & 2u) atomicAdd(bins + (int)b, 1);
                    ++mine;
                }
            }
        }
        ord += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void cur_apply_kernel(const int* bins, const int* total_in, double* state, double* curriculum,
                                 double initial_radius, double min_radius, double radius_decay, double thr, int window,
                                 double decay_factor, double* win_out, const uint32_t* comm_error) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !(comm_error && *comm_error))
        curriculum_apply(bins, total_in[1] != 0, (long long)total_in[0], state, curriculum, initial_radius, min_radius,
                         radius_decay, thr, window, decay_factor, win_out);
}

static int* curriculum_scratch() {       // chunk counts + window bins + {total, overflow}; stream-ordered reuse
    static int* buf = nullptr;
    if (!buf && cudaMalloc(&buf, (kCurMaxChunks + kCurMaxBlocks + 8) * sizeof(int)) != cudaSuccess) buf = nullptr;
    return buf;
}

int launch_curriculum_packed(const CodeSrc& src, int horizon, int n_envs, int world, double* state, double* curriculum,
                             double initial_radius, double min_radius, double radius_decay, double success_threshold,
                             int window, double decay_factor, double* window_radius_out, const uint32_t* comm_error,
                             cudaStream_t s) {
    const long long total = (long long)horizon * n_envs * world;
    const long long n_chunks = (total + kCurChunk - 1) / kCurChunk;
    bool aligned = true;
    for (int r = 0; r < world; ++r) aligned = aligned && (reinterpret_cast<uintptr_t>(src.base[r]) & 15) == 0;
    if (n_envs % 512 == 0 && n_chunks <= kCurMaxChunks && aligned) {
        int* scratch = curriculum_scratch();
        if (!scratch) return fail("curriculum: cannot allocate scratch");
        int *chunk_cnt = scratch, *bins = scratch + kCurMaxChunks, *tot = bins + kCurMaxBlocks;
        const int blocks = (int)((n_chunks * 32 + 255) / 256);
        cur_count_kernel<<<blocks, 256, 0, s>>>(src, n_envs, world, total, chunk_cnt, (int)n_chunks);
        cur_scan_kernel<<<1, 1024, 0, s>>>(chunk_cnt, (int)n_chunks, bins, tot);
        cur_bin_kernel<<<blocks, 256, 0, s>>>(src, n_envs, world, total, chunk_cnt, (int)n_chunks, state, window, bins,
                                              tot);
        cur_apply_kernel<<<1, 32, 0, s>>>(bins, tot, state, curriculum, initial_radius, min_radius, radius_decay,
                                          success_threshold, window, decay_factor, window_radius_out, comm_error);
        PLUME_LAUNCH_CHECK();
        return 0;
    }
    curriculum_kernel<<<1, 1024, 0, s>>>(FlagsPacked{src, horizon, n_envs, world, 0, 0, 0}, total, state, curriculum,
                                         initial_radius, min_radius, radius_decay, success_threshold, window,
                                         decay_factor, window_radius_out, comm_error);
    PLUME_LAUNCH_CHECK();
    return 0;
}

}  // namespace plume

using namespace plume;

extern "C" int plume_gae_scan_variant(const float* rewards, const float* values, const float* dones,
                                      const float* last_values, int32_t horizon, int32_t n_envs, double gamma,
                                      double lam, int32_t variant, float* advantages, double* stats, void* stream) {
    PLUME_CHECK_ARG(rewards && values && dones && advantages && stats, "null pointer");
    PLUME_CHECK_ARG(variant >= PLUME_GAE_QUIRK && variant <= PLUME_GAE_V12, "unknown GAE variant");
    PLUME_CHECK_ARG(variant != PLUME_GAE_BOOTSTRAP || last_values, "the bootstrap variant needs last_values");
    if (horizon <= 0 || n_envs <= 0) return 0;
    gae_scan_kernel<<<(n_envs + 127) / 128, 128, 0, as_stream(stream)>>>(
        rewards, values, dones, horizon, n_envs, (float)gamma, (float)(gamma * lam), advantages, stats, variant,
        last_values);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_gae_scan(const float* rewards, const float* values, const float* dones, int32_t horizon,
                              int32_t n_envs, double gamma, double lam, float* advantages, double* stats,
                              void* stream) {
    return plume_gae_scan_variant(rewards, values, dones, nullptr, horizon, n_envs, gamma, lam, PLUME_GAE_QUIRK,
                                  advantages, stats, stream);
}

extern "C" int plume_gae_normalise(float* advantages, const float* values, int64_t count, const double* stats,
                                   float* returns, void* stream) {
    return plume_gae_normalise_variant(advantages, values, count, stats, PLUME_GAE_QUIRK, returns, stream);
}

extern "C" int plume_gae_normalise_variant(float* advantages, const float* values, int64_t count, const double* stats,
                                           int32_t variant, float* returns, void* stream) {
    PLUME_CHECK_ARG(advantages && values && stats && returns, "null pointer");
    PLUME_CHECK_ARG(variant >= PLUME_GAE_QUIRK && variant <= PLUME_GAE_V12, "unknown GAE variant");
    if (count <= 0) return 0;
    long long blocks = (count + 255) / 256;
    const long long cap = 8LL * (sm_count() > 0 ? sm_count() : 148);
    if (blocks > cap) blocks = cap;
    gae_normalise_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(advantages, values, count, stats, returns,
                                                                          variant);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t n,
                               float max_norm, float lr, float beta1, float beta2, float eps, int32_t step,
                               float* grad_norm_out, void* stream) {
    PLUME_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "null pointer");
    PLUME_CHECK_ARG(step >= 1, "Adam step is 1-based");
    if (n <= 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    PLUME_CHECK_ARG((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "grads must be 16-byte aligned");
    clip_adam_kernel<<<(n + 1023) / 1024, 1024, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, max_norm, lr, beta1,
                                                        beta2, eps, (float)bc1, (float)sqrt(bc2), grad_norm_out);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_permutation(int64_t total, uint64_t seed, int32_t epoch, int64_t start, int64_t count,
                                 int64_t* out, void* stream) {
    PLUME_CHECK_ARG(out, "null pointer");
    PLUME_CHECK_ARG(total > 0 && start >= 0 && start + count <= total, "range outside [0,total)");
    if (count <= 0) return 0;
    permutation_kernel<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(
        total, seed, epoch, start, count, reinterpret_cast<long long*>(out));
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_curriculum_update(const float* dones, const uint8_t* reached, int32_t horizon, int32_t n_envs,
                                       double* state, double* curriculum, double initial_radius, double min_radius,
                                       double radius_decay, double success_threshold, int32_t window,
                                       double decay_factor, void* stream) {
    PLUME_CHECK_ARG(dones && reached && state && curriculum, "null pointer");
    PLUME_CHECK_ARG(window > 0, "window must be positive");
    if (horizon <= 0 || n_envs <= 0) return 0;
    curriculum_kernel<<<1, 1024, 0, as_stream(stream)>>>(FlagsSeparate{dones, reached, 0}, (long long)horizon * n_envs,
                                                         state, curriculum, initial_radius, min_radius, radius_decay,
                                                         success_threshold, window, decay_factor, nullptr, nullptr);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_curriculum_update_packed(const uint8_t* flag_code, int32_t horizon, int32_t n_envs, int32_t world,
                                              double* state, double* curriculum, double initial_radius,
                                              double min_radius, double radius_decay, double success_threshold,
                                              int32_t window, double decay_factor, double* window_radius_out,
                                              void* stream) {
    PLUME_CHECK_ARG(flag_code && state && curriculum, "null pointer");
    PLUME_CHECK_ARG(window > 0 && world >= 1 && world <= kCommMaxWorld, "window and world must be positive");
    if (horizon <= 0 || n_envs <= 0) return 0;
    CodeSrc src;
    for (int r = 0; r < kCommMaxWorld; ++r)
        src.base[r] = r < world ? flag_code + (size_t)r * horizon * n_envs : nullptr;
    return launch_curriculum_packed(src, horizon, n_envs, world, state, curriculum, initial_radius, min_radius,
                                    radius_decay, success_threshold, window, decay_factor, window_radius_out, nullptr,
                                    as_stream(stream));
}
