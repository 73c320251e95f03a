// lstm_kernels.cu -- K4: LSTM stop heads (P4L) and trend features (P4t) as standalone batch
// kernels.  V2.1 PeakAndStopPredictor: PPOV2.1/evaluate_with_lstm.py:11-27; V2.0
// ConcentrationThresholdPredictor LSTM stack: PPOV2.0/model.py:203-240; trend label:
// PPOV2.1/model.py:113-127.
#include <cstdlib>

#include "lstm_tile.cuh"

namespace plume {

// ---- V2.1 head, weights in shared memory, 32 windows per tile ------------------------------
template <int H>
__global__ void __launch_bounds__(256, 1)
lstm_stop_head_kernel(LstmWeights w, const float* __restrict__ windows, int batch, int steps,
                      float* __restrict__ peak, float* __restrict__ stop_prob) {
    extern __shared__ __align__(16) float sm[];
    lstm_load_weights<H>(sm, w);
    const int tid = threadIdx.x;
    const int tiles = (batch + 31) / 32;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int base = tile * 32;
        __syncthreads();
        for (int i = tid; i < steps * 32; i += 256) {     // xs[t][s] <- windows[base+s][t]
            const int s = i / steps, t = i - s * steps;
            sm[LstmSmem<H>::xs + t * 32 + s] = (base + s < batch) ? windows[(size_t)(base + s) * steps + t] : 0.0f;
        }
        __syncthreads();
        lstm_window_tile<H>(sm, steps);
        const float v = lstm_heads<H>(sm, steps);
        if (tid < 64) {
            const int s = tid & 31;
            if (base + s < batch) {
                if (tid < 32) peak[base + s] = v;
                else stop_prob[base + s] = v;
            }
        }
    }
}

// ---- generic multi-layer LSTM (any H <= 256), weights streamed from L2 ----------------------
// One CTA = 8 sequences, one thread per hidden unit.  seq [T][8][H] in shared memory is
// rewritten in place layer by layer.
constexpr int kGenTS = 8;

__global__ void lstm_generic_kernel(const float* __restrict__ params, int layers, int H,
                                    const float* __restrict__ windows, int batch, int steps,
                                    float* __restrict__ h_out) {
    extern __shared__ __align__(16) float sm[];
    float* seq = sm;                              // [steps][8][H]
    float* xin = sm + (size_t)steps * kGenTS * H; // [steps][8] layer-0 scalar inputs
    const int j = threadIdx.x;
    const int base = blockIdx.x * kGenTS;
    for (int i = j; i < steps * kGenTS; i += blockDim.x) {
        const int s = i / steps, t = i - s * steps;
        xin[t * kGenTS + s] = (base + s < batch) ? windows[(size_t)(base + s) * steps + t] : 0.0f;
    }
    __syncthreads();
    const float* p = params;
    for (int l = 0; l < layers; ++l) {
        const int in = (l == 0) ? 1 : H;
        const float* w_ih = p;
        const float* w_hh = w_ih + (size_t)4 * H * in;
        const float* b_ih = w_hh + (size_t)4 * H * H;
        const float* b_hh = b_ih + 4 * H;
        p = b_hh + 4 * H;
        float c[kGenTS];
#pragma unroll
        for (int s = 0; s < kGenTS; ++s) c[s] = 0.0f;
        for (int t = 0; t < steps; ++t) {
            float acc[kGenTS][4];
            if (j < H) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float b = b_ih[g * H + j] + b_hh[g * H + j];
#pragma unroll
                    for (int s = 0; s < kGenTS; ++s) acc[s][g] = b;
                }
                if (l == 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float wv = w_ih[g * H + j];
#pragma unroll
                        for (int s = 0; s < kGenTS; ++s) acc[s][g] = fmaf(xin[t * kGenTS + s], wv, acc[s][g]);
                    }
                } else {
                    const float* xt = seq + (size_t)t * kGenTS * H;
                    for (int k = 0; k < H; ++k) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float wv = w_ih[(size_t)(g * H + j) * H + k];
#pragma unroll
                            for (int s = 0; s < kGenTS; ++s) acc[s][g] = fmaf(xt[s * H + k], wv, acc[s][g]);
                        }
                    }
                }
                if (t > 0) {
                    const float* hp = seq + (size_t)(t - 1) * kGenTS * H;
                    for (int k = 0; k < H; ++k) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float wv = w_hh[(size_t)(g * H + j) * H + k];
#pragma unroll
                            for (int s = 0; s < kGenTS; ++s) acc[s][g] = fmaf(hp[s * H + k], wv, acc[s][g]);
                        }
                    }
                }
            }
            __syncthreads();      // everyone has read seq[t] (layer input) before it is overwritten
            if (j < H) {
#pragma unroll
                for (int s = 0; s < kGenTS; ++s) {
                    const float ig = sigmoidf_acc(acc[s][0]), fg = sigmoidf_acc(acc[s][1]);
                    const float gg = tanhf_acc(acc[s][2]), og = sigmoidf_acc(acc[s][3]);
                    c[s] = fmaf(fg, c[s], ig * gg);
                    seq[((size_t)t * kGenTS + s) * H + j] = og * tanhf_acc(c[s]);
                }
            }
            __syncthreads();
        }
    }
    if (j < H) {
        for (int s = 0; s < kGenTS; ++s)
            if (base + s < batch) h_out[(size_t)(base + s) * H + j] = seq[((size_t)(steps - 1) * kGenTS + s) * H + j];
    }
}

__global__ void linear_heads_kernel(const float* __restrict__ h, int batch, int H, const float* __restrict__ w_peak,
                                    const float* __restrict__ b_peak, const float* __restrict__ w_stop,
                                    const float* __restrict__ b_stop, float* peak, float* stop_prob) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    float a = 0.0f, b = 0.0f;
    for (int k = 0; k < H; ++k) {
        const float v = h[(size_t)i * H + k];
        a = fmaf(v, w_peak[k], a);
        b = fmaf(v, w_stop[k], b);
    }
    peak[i] = a + b_peak[0];
    stop_prob[i] = sigmoidf_acc(b + b_stop[0]);
}


// ---- V2.0 threshold predictor head: Linear(H,64) -> LayerNorm(64) -> ReLU -> Linear(64,1) ----------------
// (PPOV2.0/model.py:213-220; dropout is inactive in eval mode).  One 64-thread CTA per sample.
__global__ void __launch_bounds__(64) threshold_head_kernel(const float* __restrict__ h, int batch, int H,
                                                            const float* __restrict__ w1, const float* __restrict__ b1,
                                                            const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                                            float* __restrict__ out) {
    __shared__ float red[4];
    const int i = blockIdx.x, o = threadIdx.x, lane = o & 31, warp = o >> 5;
    if (i >= batch) return;
    float a = b1[o];
    for (int k = 0; k < H; ++k) a = fmaf(h[(size_t)i * H + k], w1[o * H + k], a);
    float s = a;
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    const float mean = (red[0] + red[1]) * (1.0f / 64.0f);
    const float d = a - mean;
    float q = d * d;
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    if (lane == 0) red[2 + warp] = q;
    __syncthreads();
    const float rstd = 1.0f / sqrtf((red[2] + red[3]) * (1.0f / 64.0f) + 1e-5f);
    float y = fmaxf(fmaf(d * rstd, ln_g[o], ln_b[o]), 0.0f) * w2[o];
    for (int off = 16; off > 0; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
    __syncthreads();
    if (lane == 0) red[warp] = y;
    __syncthreads();
    if (o == 0) out[i] = red[0] + red[1] + b2[0];
}

// ---- P4t trend features (device function in lstm_tile.cuh) ----
__global__ void trend_kernel(const float* __restrict__ conc, int batch, int W, const float* __restrict__ pos,
                             const double* __restrict__ src, double conc_peak, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    trend_features(conc + (size_t)i * W, W, (double)pos[2 * i], (double)pos[2 * i + 1], src[2 * i], src[2 * i + 1],
                   conc_peak, out + 4 * (size_t)i);
}


// ---- deferred stop head over a whole rollout segment -------------------------------------------------
// One tile = 32 consecutive envs at one step t; the window of (t, env) is read straight out of the
// [T][N] conc_sample rows t-W+1..t (coalesced 128 B rows), continued into the carried ring for t < W-1.
// Algorithmic bytes per (t, env): reads 4 (sample; the other W-1 window rows hit L1/L2) + 1 (fill) +
// 8 (dist); writes 4 + 1 + 4 + 16 = 25.  2*4H*(1+H)*W FLOP (H=32, W=20: 168 960).
struct SegmentArgs {
    const float* conc_sample;
    const uint8_t* fill_t;
    const double* src_dist;
    int horizon, n_envs, W;
    const float* window_in;
    float* window_out;
    double conc_peak;
    float threshold;
    float *stop_prob, *peak_pred, *trend;
    uint8_t* stop_flag;
};

template <int H>
__global__ void __launch_bounds__(256, 3) stop_head_segment_kernel(LstmWeights w, SegmentArgs a) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float s_out[64];
    lstm_load_weights<H>(sm, w);
    const int tid = threadIdx.x;
    const int N = a.n_envs, W = a.W;
    const int env_tiles = (N + 31) / 32;
    const long long tiles = (long long)env_tiles * a.horizon;
    float* xs = sm + LstmSmem<H>::xs;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t = (int)(tile / env_tiles), env0 = (int)(tile - (long long)t * env_tiles) * 32;
        __syncthreads();
        for (int i = tid; i < W * 32; i += 256) {              // xs[k][s], k = 0 oldest
            const int k = i >> 5, s = i & 31, env = env0 + s;
            const int tt = t - (W - 1) + k;
            float v = 0.0f;
            if (env < N) v = tt >= 0 ? a.conc_sample[(size_t)tt * N + env] : a.window_in[(size_t)env * W + (W + tt)];
            xs[k * 32 + s] = v;
        }
        int fill = 0;
        if (tid < 32 && env0 + tid < N) fill = a.fill_t[(size_t)t * N + env0 + tid];
        const bool full = fill >= W;
        const bool any = __syncthreads_or(full);               // also publishes xs
        if (any) {
            lstm_window_tile<H>(sm, W);
            const float hv = lstm_heads<H>(sm, W);
            if (tid < 64) s_out[tid] = hv;                     // [0,32) peak, [32,64) stop probability
            __syncthreads();
        }
        if (tid < 32 && env0 + tid < N) {
            const size_t i = (size_t)t * N + env0 + tid;
            const float peak = full ? s_out[tid] : 0.0f, stop_p = full ? s_out[32 + tid] : 0.0f;
            if (a.stop_prob) a.stop_prob[i] = stop_p;
            if (a.stop_flag) a.stop_flag[i] = (full && stop_p > a.threshold) ? 1 : 0;   // evaluate_with_lstm.py:77
            if (a.peak_pred) a.peak_pred[i] = peak;
            if (a.trend) {
                float tr[4] = {0, 0, 0, 0};
                if (full && W >= 4)
                    trend_from_last4(100.0 * (double)xs[(W - 4) * 32 + tid], 100.0 * (double)xs[(W - 3) * 32 + tid],
                                     100.0 * (double)xs[(W - 2) * 32 + tid], 100.0 * (double)xs[(W - 1) * 32 + tid],
                                     a.src_dist[i], a.conc_peak, tr);
                *reinterpret_cast<float4*>(a.trend + i * 4) = make_float4(tr[0], tr[1], tr[2], tr[3]);
            }
            if (t == a.horizon - 1 && a.window_out)            // the ring the next segment continues from
                for (int k = 0; k < W; ++k) a.window_out[(size_t)(env0 + tid) * W + k] = xs[k * 32 + tid];
        }
    }
}

template <int H>
static int launch_segment(const LstmWeights& w, const SegmentArgs& a, cudaStream_t s) {
    static bool configured = false;
    const int smem = LstmSmem<H>::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(stop_head_segment_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return fail("stop-head segment kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const long long tiles = (long long)((a.n_envs + 31) / 32) * a.horizon;
    long long grid = 3LL * sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    stop_head_segment_kernel<H><<<(int)grid, 256, smem, s>>>(w, a);
    if (cudaGetLastError() != cudaSuccess) return fail("stop-head segment kernel launch failed");
    return 0;
}

// ---- deferred stop head for any hidden size <= 256 (BASELINE configs[4]: hidden 256) -------------------------
// Windows of a chunk of (t, env) pairs are assembled into a scratch array, the generic LSTM kernel (weights
// streamed from L2) produces their last hidden states and one warp per window applies fc_peak / fc_stop and the
// trend features.  Same outputs and window semantics as stop_head_segment_kernel.
constexpr int kSegChunk = 32768;

__global__ void segment_windows_kernel(SegmentArgs a, long long first, int count, float* __restrict__ windows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long idx = first + i;
    const int t = (int)(idx / a.n_envs), env = (int)(idx - (long long)t * a.n_envs), W = a.W;
    for (int k = 0; k < W; ++k) {
        const int tt = t - (W - 1) + k;
        const float v = tt >= 0 ? a.conc_sample[(size_t)tt * a.n_envs + env] : a.window_in[(size_t)env * W + (W + tt)];
        windows[(size_t)i * W + k] = v;
        if (t == a.horizon - 1 && a.window_out) a.window_out[(size_t)env * W + k] = v;
    }
}

__global__ void segment_heads_kernel(SegmentArgs a, LstmWeights w, int H, long long first, int count,
                                     const float* __restrict__ h, const float* __restrict__ windows) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= count) return;
    const long long idx = first + i;
    float p = 0.0f, q = 0.0f;
    for (int k = lane; k < H; k += 32) {
        const float v = h[(size_t)i * H + k];
        p = fmaf(v, w.w_peak[k], p);
        q = fmaf(v, w.w_stop[k], q);
    }
    for (int off = 16; off > 0; off >>= 1) {
        p += __shfl_xor_sync(0xffffffffu, p, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane != 0) return;
    const int W = a.W;
    const bool full = a.fill_t[idx] >= W;
    const float peak = full ? p + w.b_peak[0] : 0.0f, stop_p = full ? sigmoidf_acc(q + w.b_stop[0]) : 0.0f;
    if (a.stop_prob) a.stop_prob[idx] = stop_p;
    if (a.stop_flag) a.stop_flag[idx] = (full && stop_p > a.threshold) ? 1 : 0;
    if (a.peak_pred) a.peak_pred[idx] = peak;
    if (a.trend) {
        float tr[4] = {0, 0, 0, 0};
        const float* x = windows + (size_t)i * W;
        if (full && W >= 4)
            trend_from_last4(100.0 * (double)x[W - 4], 100.0 * (double)x[W - 3], 100.0 * (double)x[W - 2],
                             100.0 * (double)x[W - 1], a.src_dist[idx], a.conc_peak, tr);
        *reinterpret_cast<float4*>(a.trend + idx * 4) = make_float4(tr[0], tr[1], tr[2], tr[3]);
    }
}

static int launch_generic(const float* params, int layers, int H, const float* windows, int batch, int steps,
                          float* h_out, cudaStream_t s);

// scratch of the generic-hidden-size paths: stream-ordered reuse, grown on demand (cudaFree synchronises)
static float* generic_scratch(size_t need_floats) {
    static float* scratch = nullptr;
    static size_t scratch_floats = 0;
    if (need_floats > scratch_floats) {
        if (scratch) cudaFree(scratch);
        scratch = nullptr;
        scratch_floats = 0;
        if (cudaMalloc(&scratch, need_floats * sizeof(float)) != cudaSuccess) return nullptr;
        scratch_floats = need_floats;
    }
    return scratch;
}

// torch layout of layer 0: weight_ih [4H][1], weight_hh [4H][H], bias_ih [4H], bias_hh [4H]
static int pack_layer0(float* par, const LstmWeights& w, int H, cudaStream_t s) {
    PLUME_CUDA(cudaMemcpyAsync(par, w.w_ih, sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, s));
    PLUME_CUDA(cudaMemcpyAsync(par + 4 * H, w.w_hh, sizeof(float) * 4 * H * H, cudaMemcpyDeviceToDevice, s));
    PLUME_CUDA(cudaMemcpyAsync(par + 4 * H + (size_t)4 * H * H, w.b_ih, sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, s));
    PLUME_CUDA(cudaMemcpyAsync(par + 8 * H + (size_t)4 * H * H, w.b_hh, sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, s));
    return 0;
}

// PeakAndStopPredictor.forward on explicit windows [batch][steps] for any hidden size <= 256
static int launch_stop_head_generic(const LstmWeights& w, int H, const float* windows, int batch, int steps,
                                    float* peak, float* stop_prob, cudaStream_t s) {
    const size_t n_par = (size_t)4 * H + (size_t)4 * H * H + (size_t)8 * H;
    float* par = generic_scratch(n_par + (size_t)kSegChunk * H);
    if (!par) return fail("stop head: cannot allocate scratch");
    float* hbuf = par + n_par;
    if (pack_layer0(par, w, H, s)) return 1;
    for (int first = 0; first < batch; first += kSegChunk) {
        const int count = batch - first < kSegChunk ? batch - first : kSegChunk;
        if (launch_generic(par, 1, H, windows + (size_t)first * steps, count, steps, hbuf, s)) return 1;
        linear_heads_kernel<<<(count + 127) / 128, 128, 0, s>>>(hbuf, count, H, w.w_peak, w.b_peak, w.w_stop, w.b_stop,
                                                               peak + first, stop_prob + first);
    }
    PLUME_LAUNCH_CHECK();
    return 0;
}

static int launch_segment_generic(const LstmWeights& w, int H, const SegmentArgs& a, cudaStream_t s) {
    const size_t n_par = (size_t)4 * H + (size_t)4 * H * H + (size_t)8 * H;
    float* par = generic_scratch(n_par + (size_t)kSegChunk * a.W + (size_t)kSegChunk * H);
    if (!par) return fail("stop head: cannot allocate scratch");
    float* win = par + n_par;
    float* hbuf = win + (size_t)kSegChunk * a.W;
    if (pack_layer0(par, w, H, s)) return 1;
    const long long total = (long long)a.horizon * a.n_envs;
    for (long long first = 0; first < total; first += kSegChunk) {
        const int count = (int)((total - first) < kSegChunk ? (total - first) : kSegChunk);
        segment_windows_kernel<<<(count + 255) / 256, 256, 0, s>>>(a, first, count, win);
        if (launch_generic(par, 1, H, win, count, a.W, hbuf, s)) return 1;
        segment_heads_kernel<<<(count + 7) / 8, 256, 0, s>>>(a, w, H, first, count, hbuf, win);
    }
    PLUME_LAUNCH_CHECK();
    return 0;
}

template <int H>
static int launch_stop_head(const LstmWeights& w, const float* windows, int batch, int steps, float* peak,
                            float* stop_prob, cudaStream_t s) {
    static bool configured = false;
    const int smem = LstmSmem<H>::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(lstm_stop_head_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return fail("lstm kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const int tiles = (batch + 31) / 32;
    int grid = 2 * sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    lstm_stop_head_kernel<H><<<grid, 256, smem, s>>>(w, windows, batch, steps, peak, stop_prob);
    if (cudaGetLastError() != cudaSuccess) return fail("lstm kernel launch failed");
    return 0;
}

static int launch_generic(const float* params, int layers, int H, const float* windows, int batch, int steps,
                          float* h_out, cudaStream_t s) {
    const size_t smem = ((size_t)steps * kGenTS * H + (size_t)steps * kGenTS) * sizeof(float);
    if (smem > 227 * 1024) return fail("lstm: window %d x hidden %d does not fit in shared memory", steps, H);
    if (cudaFuncSetAttribute(lstm_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail("lstm generic kernel: cannot reserve %zu B of shared memory", smem);
    const int threads = ((H + 31) / 32) * 32;
    lstm_generic_kernel<<<(batch + kGenTS - 1) / kGenTS, threads, smem, s>>>(params, layers, H, windows, batch, steps,
                                                                             h_out);
    if (cudaGetLastError() != cudaSuccess) return fail("lstm generic kernel launch failed");
    return 0;
}

}  // namespace plume

using namespace plume;

extern "C" int plume_lstm_stop_head(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                    const float* w_peak, const float* b_peak, const float* w_stop, const float* b_stop,
                                    int32_t hidden, const float* windows, int32_t batch, int32_t steps, float* peak,
                                    float* stop_prob, void* stream) {
    PLUME_CHECK_ARG(w_ih && w_hh && b_ih && b_hh && w_peak && b_peak && w_stop && b_stop && windows && peak && stop_prob,
                    "null pointer");
    PLUME_CHECK_ARG(steps >= 1 && steps <= kLstmMaxSteps, "window length must be in [1,32]");
    if (batch <= 0) return 0;
    const LstmWeights w{w_ih, w_hh, b_ih, b_hh, w_peak, b_peak, w_stop, b_stop};
    if (hidden == 32) return launch_stop_head<32>(w, windows, batch, steps, peak, stop_prob, as_stream(stream));
    if (hidden == 64) return launch_stop_head<64>(w, windows, batch, steps, peak, stop_prob, as_stream(stream));
    if (hidden >= 1 && hidden <= 256)
        return launch_stop_head_generic(w, hidden, windows, batch, steps, peak, stop_prob, as_stream(stream));
    return fail("plume_lstm_stop_head: hidden must be in [1,256]");
}

extern "C" int plume_stop_head_segment(const plume_lstm_params* lstm, const float* conc_sample, const uint8_t* fill_t,
                                       const double* src_dist, int32_t horizon, int32_t n_envs,
                                       const float* window_in, float* window_out, double conc_peak, float* stop_prob,
                                       uint8_t* stop_flag, float* peak_pred, float* trend, int32_t kernel_path,
                                       void* stream) {
    PLUME_CHECK_ARG(lstm && conc_sample && fill_t && window_in, "null pointer");
    PLUME_CHECK_ARG(kernel_path >= PLUME_KERNEL_AUTO && kernel_path <= PLUME_KERNEL_SIMT, "unknown kernel_path");
    PLUME_CHECK_ARG(lstm->w_ih && lstm->w_hh && lstm->b_ih && lstm->b_hh && lstm->w_peak && lstm->b_peak &&
                        lstm->w_stop && lstm->b_stop, "null LSTM parameter");
    PLUME_CHECK_ARG(lstm->window >= 1 && lstm->window <= kLstmMaxSteps, "stop-head window must be in [1,32]");
    PLUME_CHECK_ARG(!trend || src_dist, "trend features need src_dist");
    PLUME_CHECK_ARG(window_out != window_in, "window_out must not alias window_in");
    if (horizon <= 0 || n_envs <= 0) return 0;
    // gate GEMM on the tensor cores where the hidden size has a tcgen05 kernel (lstm_tc_kernels.cu); kernel_path =
    // PLUME_KERNEL_SIMT selects the CUDA-core kernel, whose arithmetic is bit-identical to the in-loop head of
    // plume_rollout
    const bool resident = stop_head_segment_tc_supports(lstm->hidden);
    const bool streamed = stop_head_segment_stream_supports(lstm->hidden) && lstm->window <= 20;
    PLUME_CHECK_ARG(kernel_path != PLUME_KERNEL_TENSOR || resident || streamed,
                    "no tensor-core stop-head kernel for this hidden size");
    if ((resident || streamed) && kernel_path != PLUME_KERNEL_SIMT) {
        LtArgs t;
        t.conc_sample = conc_sample;
        t.fill_t = fill_t;
        t.src_dist = src_dist;
        t.horizon = horizon;
        t.n_envs = n_envs;
        t.W = lstm->window;
        t.window_in = window_in;
        t.window_out = window_out;
        t.conc_peak = conc_peak;
        t.threshold = lstm->threshold;
        t.stop_prob = stop_prob;
        t.peak_pred = peak_pred;
        t.trend = trend;
        t.stop_flag = stop_flag;
        t.w_ih = lstm->w_ih; t.w_hh = lstm->w_hh; t.b_ih = lstm->b_ih; t.b_hh = lstm->b_hh;
        t.w_peak = lstm->w_peak; t.b_peak = lstm->b_peak; t.w_stop = lstm->w_stop; t.b_stop = lstm->b_stop;
        return resident ? launch_stop_head_segment_tc(t, lstm->hidden, as_stream(stream))
                        : launch_stop_head_segment_stream(t, lstm->hidden, as_stream(stream));
    }
    const LstmWeights w{lstm->w_ih, lstm->w_hh, lstm->b_ih, lstm->b_hh, lstm->w_peak, lstm->b_peak, lstm->w_stop,
                        lstm->b_stop};
    SegmentArgs a;
    a.conc_sample = conc_sample;
    a.fill_t = fill_t;
    a.src_dist = src_dist;
    a.horizon = horizon;
    a.n_envs = n_envs;
    a.W = lstm->window;
    a.window_in = window_in;
    a.window_out = window_out;
    a.conc_peak = conc_peak;
    a.threshold = lstm->threshold;
    a.stop_prob = stop_prob;
    a.peak_pred = peak_pred;
    a.trend = trend;
    a.stop_flag = stop_flag;
    if (lstm->hidden == 32) return launch_segment<32>(w, a, as_stream(stream));
    if (lstm->hidden == 64) return launch_segment<64>(w, a, as_stream(stream));
    if (lstm->hidden >= 1 && lstm->hidden <= 256) return launch_segment_generic(w, lstm->hidden, a, as_stream(stream));
    return fail("plume_stop_head_segment: hidden must be in [1,256]");
}

extern "C" int plume_lstm_forward(const float* params, int32_t layers, int32_t hidden, const float* windows,
                                  int32_t batch, int32_t steps, float* h_out, void* stream) {
    PLUME_CHECK_ARG(params && windows && h_out, "null pointer");
    PLUME_CHECK_ARG(layers >= 1 && layers <= 8, "layers must be in [1,8]");
    PLUME_CHECK_ARG(hidden >= 1 && hidden <= 256, "hidden must be in [1,256]");
    PLUME_CHECK_ARG(steps >= 1, "empty window");
    if (batch <= 0) return 0;
    return launch_generic(params, layers, hidden, windows, batch, steps, h_out, as_stream(stream));
}

extern "C" int plume_threshold_head(const float* h, int32_t batch, int32_t hidden, const float* w1, const float* b1,
                                    const float* ln_weight, const float* ln_bias, const float* w2, const float* b2,
                                    float* out, void* stream) {
    PLUME_CHECK_ARG(h && w1 && b1 && ln_weight && ln_bias && w2 && b2 && out, "null pointer");
    PLUME_CHECK_ARG(hidden >= 1, "hidden must be positive");
    if (batch <= 0) return 0;
    threshold_head_kernel<<<batch, 64, 0, as_stream(stream)>>>(h, batch, hidden, w1, b1, ln_weight, ln_bias, w2, b2,
                                                              out);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_trend_features(const float* conc, int32_t batch, int32_t window, const float* pos_last,
                                    const double* src, double conc_peak, float* out, void* stream) {
    PLUME_CHECK_ARG(conc && pos_last && src && out, "null pointer");
    PLUME_CHECK_ARG(window >= 2, "window must hold at least 2 samples");
    if (batch <= 0) return 0;
    trend_kernel<<<(batch + 127) / 128, 128, 0, as_stream(stream)>>>(conc, batch, window, pos_last, src, conc_peak, out);
    PLUME_LAUNCH_CHECK();
    return 0;
}
