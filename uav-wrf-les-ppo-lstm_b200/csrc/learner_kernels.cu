// learner_kernels.cu -- K5 GAE reverse scan + normalisation (train_ppo2.0.py:17-39), K7
// clip_grad_norm + Adam (train_ppo2.0.py:86-87,113), the stateless minibatch permutation
// (replaces torch.randperm, :43) and K8 the curriculum (PPOTrainer.update, model.py:188-221).
//
// Algorithmic bytes (DESIGN.md "K5"): scan reads r,V,done (12 B) and writes A (4 B); the
// normalise pass reads A,V (8 B) and writes A,ret (8 B) => 32 B per transition.
#include "common.cuh"

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
        for (int t = T - 1; t >= 0; --t) {
            const size_t i = (size_t)t * N + n;
            const float r = rewards[i], v = values[i], d = dones[i];
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
constexpr int kCurMaxBlocks = 8192;

__global__ void __launch_bounds__(1024) curriculum_kernel(const float* __restrict__ dones,
                                                          const uint8_t* __restrict__ reached, long long total,
                                                          double* state, double* curriculum, double initial_radius,
                                                          double min_radius, double radius_decay, double thr,
                                                          int window, double decay_factor) {
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
#pragma unroll 4
    for (long long i = lo; i < hi; i += 32) {
        const bool d = (i + lane < hi) && dones[i + lane] != 0.0f;
        eps += __popc(__ballot_sync(0xffffffffu, d));
    }
    if (lane == 0) ep_cnt[warp] = eps;
    for (int i = tid; i < kCurMaxBlocks; i += blockDim.x) blk_succ[i] = 0;
    __syncthreads();
    const long long hist_len = (long long)state[4];
    long long ord = hist_len;                                          // ordinal of this warp's first episode
    for (int w = 0; w < warp; ++w) ord += ep_cnt[w];
    bool overflow = false;
    for (long long i = lo; i < hi; i += 32) {
        const bool d = (i + lane < hi) && dones[i + lane] != 0.0f;
        const unsigned mask = __ballot_sync(0xffffffffu, d);
        if (d) {
            const long long b = (ord + __popc(mask & ((1u << lane) - 1u))) / window;
            if (b >= kCurMaxBlocks) overflow = true;
            else if (reached[i + lane]) atomicAdd(&blk_succ[(int)b], 1);
        }
        ord += __popc(mask);
    }
    if (overflow) atomicOr(&overflow_any, 1);
    __syncthreads();
    if (tid == (int)blockDim.x - 1) {
        if (overflow_any) {
            state[4] = -1.0;   // host raises
            return;
        }
        const long long total_eps = ord - hist_len;      // last thread's ordinal = all episodes
        double radius = state[0], eb = state[1];
        double env_radius = state[2], env_eb = state[3];
        long long succ_total = 0;
        const long long n_done = hist_len + total_eps;
        const long long full = n_done / window;
        for (long long b = 0; b <= full && b < kCurMaxBlocks; ++b) {
            long long s = blk_succ[b];
            succ_total += s;
            if (b == 0) s += (long long)state[5];
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
    curriculum_kernel<<<1, 1024, 0, as_stream(stream)>>>(dones, reached, (long long)horizon * n_envs, state,
                                                         curriculum, initial_radius, min_radius, radius_decay,
                                                         success_threshold, window, decay_factor);
    PLUME_LAUNCH_CHECK();
    return 0;
}
