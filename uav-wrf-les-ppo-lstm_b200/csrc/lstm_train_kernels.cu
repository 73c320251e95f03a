// lstm_train_kernels.cu -- supervised training of the V2.1 stop head on the GPU.
//
// Reference: PPOV2.1/train_lstm.py
//   :28-65   TrajectoryDataset._preprocess: per selected episode, from its FIRST window of `window` steps
//            (the segment list is sliding windows and the code takes ep_segs[0]): a "negative" sample
//            (features = conc[:window]/100, labels = [conc[window-1]/100, 0]) and a "positive" sample with the
//            SAME features (conc[-window:] of a window-long segment) and labels
//            [conc[window-1]/100, 1 if ||pos[window-1] - source|| <= stop_radius else 0]
//   :84-100  PeakAndStopPredictor: LSTM(1 -> H) from zero state, fc_peak, fc_stop + sigmoid on h_T
//   :102-125 loop: DataLoader(batch 64, shuffle), loss = MSE(peak) + BCE(stop), clip_grad_norm_(1.0),
//            AdamW(lr 1e-3, weight_decay 1e-4), ReduceLROnPlateau on the epoch mean (host side, lstm_train.py)
//
// One launch = one optimiser step: forward, back-propagation through time, gradient reduction, clip and AdamW.
// A warp owns one sequence and a lane one hidden unit: the four gate rows of the unit (4 x 32 recurrent weights)
// live in registers for the forward pass, the transposed columns for the backward pass; h and the gate
// derivatives are exchanged through a warp-private shared-memory row read as broadcast float4.  Activations of
// all steps stay in shared memory (window x 6 x 32 floats per sequence), nothing but the final partial gradients
// touches HBM.  Each CTA (8 sequences) writes one partial gradient; the last CTA to arrive (ticket) adds the
// partials in CTA order -- the result does not depend on scheduling -- and applies clip + AdamW from registers.
//
// Algorithmic work per sequence: forward 2*4H*(1+H)*T = 169 kFLOP (H=32, T=20), backward twice that.
#include <cmath>

#include "common.cuh"

namespace plume {

constexpr int kLtH = 32;                 // hidden size (train_lstm.py:85)
constexpr int kLtWarps = 8;              // sequences per CTA
constexpr int kLtThreads = kLtWarps * 32;
constexpr int kLtRow = 6 * kLtH;         // per (sequence, step): i, f, g, o (later their pre-activation grads), c, h_prev
constexpr int kLtWs = 33;                // padded row of the staged recurrent matrix
constexpr int kLtMaxSteps = 32;

// flat parameter layout = torch named_parameters() order of PeakAndStopPredictor
constexpr int kLtOffWih = 0;
constexpr int kLtOffWhh = kLtOffWih + 4 * kLtH;
constexpr int kLtOffBih = kLtOffWhh + 4 * kLtH * kLtH;
constexpr int kLtOffBhh = kLtOffBih + 4 * kLtH;
constexpr int kLtOffWp = kLtOffBhh + 4 * kLtH;
constexpr int kLtOffBp = kLtOffWp + kLtH;
constexpr int kLtOffWs = kLtOffBp + 1;
constexpr int kLtOffBs = kLtOffWs + kLtH;
constexpr int kLtParams = kLtOffBs + 1;  // 4546
static_assert(kLtParams == PLUME_LSTM_TRAIN_PARAMS, "flat layout disagrees with include/plume_b200.h");
constexpr int kLtStride = (kLtParams + 3) & ~3;      // per-CTA partial gradient, 16-byte aligned rows
constexpr int kLtPerThread = (kLtParams + kLtThreads - 1) / kLtThreads;

struct LstmTrainArgs {
    float* params;
    float* exp_avg;
    float* exp_avg_sq;
    const float* features;   // [n][T]
    const float* labels;     // [n][2]
    const int32_t* order;    // sample ids of this minibatch, or NULL = base .. base+batch-1
    int32_t base, batch, T;
    float max_norm, lr, b1, b2, eps, wd, bc1, bc2_sqrt;
    float* partial;          // [grid][kLtStride]
    float* partial_loss;     // [grid]
    unsigned int* ticket;
    float* loss_out;         // [1] mean minibatch loss
    float* grad_norm_out;    // [1] or NULL
    float* grad_out;         // [kLtParams] or NULL: the (unclipped) minibatch gradient
};

// Gate activations on the critical path of a 20-step recurrence that runs one warp per scheduler: ex2.approx and
// the fast divide (relative error ~3e-7, the same functions as the inference head in lstm_tile.cuh) instead of
// expf / IEEE division / tanhf, whose 100+ dependent instructions per call dominated the first ncu capture.
__device__ __forceinline__ float lt_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float lt_tanh(float x) {
    const float t = __expf(-2.0f * fabsf(x));
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kLtThreads, 1) lstm_train_kernel(LstmTrainArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* Ws = sm;                                   // [4H][33]
    float* hist = Ws + 4 * kLtH * kLtWs;              // [warps][T][6][32]
    float* xs = hist + kLtWarps * a.T * kLtRow;       // [warps][32]
    float* headg = xs + kLtWarps * 32;                // [warps][2H + 2]
    float* lossw = headg + kLtWarps * (2 * kLtH + 2); // [warps]
    __shared__ double red[kLtWarps];
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.T;
    for (int i = tid; i < 4 * kLtH * kLtH; i += kLtThreads) Ws[(i >> 5) * kLtWs + (i & 31)] = a.params[kLtOffWhh + i];
    const int b = blockIdx.x * kLtWarps + warp;
    const bool live = b < a.batch;
    const int sid = live ? (a.order ? a.order[b] : a.base + b) : 0;
    if (lane < T) xs[warp * 32 + lane] = live ? a.features[(size_t)sid * T + lane] : 0.0f;
    __syncthreads();

    // ---- forward: lane = hidden unit, rows g*H+lane of w_hh in registers -------------------------------------
    float* hw = hist + (size_t)warp * T * kLtRow;
    float h = 0.0f, c = 0.0f;
    {
        float wr[4][kLtH];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int k = 0; k < kLtH; ++k) wr[g][k] = Ws[(g * kLtH + lane) * kLtWs + k];
        float wx[4], bb[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            wx[g] = a.params[kLtOffWih + g * kLtH + lane];
            bb[g] = a.params[kLtOffBih + g * kLtH + lane] + a.params[kLtOffBhh + g * kLtH + lane];
        }
        for (int t = 0; t < T; ++t) {
            float* row = hw + t * kLtRow;
            row[5 * kLtH + lane] = h;                 // h_{t-1}
            __syncwarp();
            const float x = xs[warp * 32 + t];
            float acc[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[g] = fmaf(wx[g], x, bb[g]);
#pragma unroll
            for (int k4 = 0; k4 < kLtH / 4; ++k4) {
                const float4 hv = *reinterpret_cast<const float4*>(row + 5 * kLtH + 4 * k4);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    acc[g] = fmaf(wr[g][4 * k4 + 0], hv.x, acc[g]);
                    acc[g] = fmaf(wr[g][4 * k4 + 1], hv.y, acc[g]);
                    acc[g] = fmaf(wr[g][4 * k4 + 2], hv.z, acc[g]);
                    acc[g] = fmaf(wr[g][4 * k4 + 3], hv.w, acc[g]);
                }
            }
            const float gi = lt_sigmoid(acc[0]), gf = lt_sigmoid(acc[1]), gg = lt_tanh(acc[2]), go = lt_sigmoid(acc[3]);
            c = fmaf(gf, c, gi * gg);
            h = go * lt_tanh(c);
            row[0 * kLtH + lane] = gi;
            row[1 * kLtH + lane] = gf;
            row[2 * kLtH + lane] = gg;
            row[3 * kLtH + lane] = go;
            row[4 * kLtH + lane] = c;
        }
    }

    // ---- heads and loss (train_lstm.py:94-98,113-116) ----------------------------------------------------------
    const float wp = a.params[kLtOffWp + lane], ws = a.params[kLtOffWs + lane];
    const float peak = warp_sum(wp * h) + a.params[kLtOffBp];
    const float logit = warp_sum(ws * h) + a.params[kLtOffBs];
    const float p = 1.0f / (1.0f + expf(-logit));          // once per sequence: the accurate form
    float dpeak = 0.0f, dlogit = 0.0f, loss = 0.0f;
    if (live) {
        const float y0 = a.labels[(size_t)sid * 2], y1 = a.labels[(size_t)sid * 2 + 1];
        const float inv_b = 1.0f / (float)a.batch;
        const float e0 = peak - y0;
        // nn.BCELoss clamps both logs at -100
        const float lp = fmaxf(logf(p), -100.0f), lq = fmaxf(logf(1.0f - p), -100.0f);
        loss = e0 * e0 - (y1 * lp + (1.0f - y1) * lq);
        dpeak = 2.0f * e0 * inv_b;
        // binary_cross_entropy backward: (p - y) / max((1 - p) p, 1e-12), then the sigmoid's p (1 - p)
        const float dp = (p - y1) / fmaxf((1.0f - p) * p, 1e-12f) * inv_b;
        dlogit = dp * p * (1.0f - p);
    }
    {
        float* hg = headg + warp * (2 * kLtH + 2);
        hg[lane] = dpeak * h;
        hg[kLtH + lane] = dlogit * h;
        if (lane == 0) {
            hg[2 * kLtH] = dpeak;
            hg[2 * kLtH + 1] = dlogit;
            lossw[warp] = loss;
        }
    }

    // ---- back-propagation through time: lane = input unit k, column k of w_hh in registers ---------------------
    {
        float wc[4 * kLtH];
#pragma unroll
        for (int r = 0; r < 4 * kLtH; ++r) wc[r] = Ws[r * kLtWs + lane];
        float dh = dpeak * wp + dlogit * ws, dc = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            float* row = hw + t * kLtRow;
            const float gi = row[lane], gf = row[kLtH + lane], gg = row[2 * kLtH + lane], go = row[3 * kLtH + lane];
            const float ct = row[4 * kLtH + lane];
            const float cprev = t > 0 ? row[4 * kLtH + lane - kLtRow] : 0.0f;
            const float tc = lt_tanh(ct);
            const float d_o = dh * tc;
            dc = fmaf(dh * go, 1.0f - tc * tc, dc);
            const float d_i = dc * gg, d_g = dc * gi, d_f = dc * cprev;
            row[lane] = d_i * gi * (1.0f - gi);
            row[kLtH + lane] = d_f * gf * (1.0f - gf);
            row[2 * kLtH + lane] = d_g * (1.0f - gg * gg);
            row[3 * kLtH + lane] = d_o * go * (1.0f - go);
            dc *= gf;
            __syncwarp();
            float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
            for (int r4 = 0; r4 < kLtH; r4 += 2) {
                const float4 d0 = *reinterpret_cast<const float4*>(row + 4 * r4);
                const float4 d1 = *reinterpret_cast<const float4*>(row + 4 * r4 + 4);
                acc0 = fmaf(wc[4 * r4 + 0], d0.x, acc0);
                acc0 = fmaf(wc[4 * r4 + 1], d0.y, acc0);
                acc0 = fmaf(wc[4 * r4 + 2], d0.z, acc0);
                acc0 = fmaf(wc[4 * r4 + 3], d0.w, acc0);
                acc1 = fmaf(wc[4 * r4 + 4], d1.x, acc1);
                acc1 = fmaf(wc[4 * r4 + 5], d1.y, acc1);
                acc1 = fmaf(wc[4 * r4 + 6], d1.z, acc1);
                acc1 = fmaf(wc[4 * r4 + 7], d1.w, acc1);
            }
            dh = acc0 + acc1;
        }
    }
    __syncthreads();

    // ---- this CTA's partial gradient: dW_hh[r][k] = sum over (sequence, step) dpre[r] h_prev[k] ------------------
    float* part = a.partial + (size_t)blockIdx.x * kLtStride;
    {
        const int r = tid & (4 * kLtH - 1), half = tid >> 7;          // gate row, 16-column half
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
        float ax = 0.0f, ab = 0.0f;
        for (int w = 0; w < kLtWarps; ++w) {
#pragma unroll 4
            for (int t = 0; t < T; ++t) {
                const float* row = hist + ((size_t)w * T + t) * kLtRow;
                const float d = row[r];
                const float4* hp = reinterpret_cast<const float4*>(row + 5 * kLtH + half * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = hp[q];
                    acc[4 * q + 0] = fmaf(d, v.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(d, v.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(d, v.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(d, v.w, acc[4 * q + 3]);
                }
                ax = fmaf(d, xs[w * 32 + t], ax);
                ab += d;
            }
        }
        float4* dst = reinterpret_cast<float4*>(part + kLtOffWhh + r * kLtH + half * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        if (half == 0) {
            part[kLtOffWih + r] = ax;
            part[kLtOffBih + r] = ab;
            part[kLtOffBhh + r] = ab;
        }
        if (tid < 2 * kLtH + 2) {
            float s = 0.0f;
            for (int w = 0; w < kLtWarps; ++w) s += headg[w * (2 * kLtH + 2) + tid];
            const int off = tid < kLtH ? kLtOffWp + tid
                                       : (tid < 2 * kLtH ? kLtOffWs + tid - kLtH : (tid == 2 * kLtH ? kLtOffBp : kLtOffBs));
            part[off] = s;
        }
        if (tid == 0) {
            float s = 0.0f;
            for (int w = 0; w < kLtWarps; ++w) s += lossw[w];
            a.partial_loss[blockIdx.x] = s;
        }
    }

    // ---- last CTA: reduce the partials in CTA order, clip_grad_norm_, AdamW -------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float g[kLtPerThread];
    double ss = 0.0;
#pragma unroll
    for (int q = 0; q < kLtPerThread; ++q) {
        const int i = q * kLtThreads + tid;
        float s = 0.0f;
        if (i < kLtParams) {
#pragma unroll 8
            for (unsigned int cta = 0; cta < gridDim.x; ++cta) s += __ldcg(a.partial + (size_t)cta * kLtStride + i);
        }
        g[q] = s;
        ss += (double)s * (double)s;
        if (a.grad_out && i < kLtParams) a.grad_out[i] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < kLtWarps; ++w) tot += red[w];
    const float norm = (float)sqrt(tot);
    float coef = a.max_norm / (norm + 1e-6f);                 // torch.nn.utils.clip_grad_norm_
    coef = coef > 1.0f ? 1.0f : coef;
    const float step_size = a.lr / a.bc1;
#pragma unroll
    for (int q = 0; q < kLtPerThread; ++q) {
        const int i = q * kLtThreads + tid;
        if (i < kLtParams) {
            const float gi = g[q] * coef;
            float pv = a.params[i];
            pv -= pv * (a.lr * a.wd);                                    // AdamW: param.mul_(1 - lr * weight_decay)
            const float mi = a.exp_avg[i] + (gi - a.exp_avg[i]) * (1.0f - a.b1);
            const float vi = a.exp_avg_sq[i] * a.b2 + (1.0f - a.b2) * gi * gi;
            a.exp_avg[i] = mi;
            a.exp_avg_sq[i] = vi;
            const float denom = sqrtf(vi) / a.bc2_sqrt + a.eps;
            a.params[i] = pv - step_size * (mi / denom);
        }
    }
    if (tid == 0) {
        float s = 0.0f;
        for (unsigned int cta = 0; cta < gridDim.x; ++cta) s += __ldcg(a.partial_loss + cta);
        *a.loss_out = s / (float)a.batch;
        if (a.grad_norm_out) *a.grad_norm_out = norm;
        *a.ticket = 0u;                                                  // ready for the next launch
    }
}

// TrajectoryDataset._preprocess (train_lstm.py:28-65) for the selected episodes: two samples per episode.
__global__ void lstm_dataset_kernel(const float* conc, const float* x, const float* y, const float* src_x,
                                    const float* src_y, int max_steps, const int32_t* episode_ids, int n_sel, int window,
                                    float stop_radius, float* features, float* labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sel * window) return;
    const int e = i / window, t = i - e * window;
    const int ep = episode_ids[e];
    const float cv = conc[(size_t)ep * max_steps + t];
    const float f = __fdiv_rn(cv, 100.0f);                               // float32 / python float -> float32
    features[((size_t)2 * e) * window + t] = f;                          // negative sample  (:44-52)
    features[((size_t)2 * e + 1) * window + t] = f;                      // positive sample: same window (:54-63)
    if (t == window - 1) {
        const float dx = __fsub_rn(x[(size_t)ep * max_steps + t], src_x[ep]);
        const float dy = __fsub_rn(y[(size_t)ep * max_steps + t], src_y[ep]);
        const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        labels[4 * (size_t)e + 0] = f;
        labels[4 * (size_t)e + 1] = 0.0f;
        labels[4 * (size_t)e + 2] = f;
        labels[4 * (size_t)e + 3] = dist <= stop_radius ? 1.0f : 0.0f;
    }
}

static size_t lt_smem_bytes(int T) {
    return sizeof(float) * (size_t)(4 * kLtH * kLtWs + kLtWarps * T * kLtRow + kLtWarps * 32 +
                                    kLtWarps * (2 * kLtH + 2) + kLtWarps);
}

}  // namespace plume

using namespace plume;

extern "C" int64_t plume_lstm_train_workspace_bytes(int32_t batch_size) {
    if (batch_size <= 0) return 0;
    const int64_t grid = (batch_size + kLtWarps - 1) / kLtWarps;
    return 256 + 4 * ((grid + 63) / 64 * 64) + 4 * grid * (int64_t)kLtStride;
}

extern "C" int plume_lstm_dataset(const float* conc, const float* x, const float* y, const float* src_x,
                                  const float* src_y, int32_t max_steps, const int32_t* episode_ids, int32_t n_selected,
                                  int32_t window, float stop_radius, float* features, float* labels, void* stream) {
    PLUME_CHECK_ARG(conc && x && y && src_x && src_y && episode_ids && features && labels, "null pointer");
    PLUME_CHECK_ARG(window >= 1 && window <= max_steps, "window must be in [1,max_steps]");
    if (n_selected <= 0) return 0;
    const long long n = (long long)n_selected * window;
    lstm_dataset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        conc, x, y, src_x, src_y, max_steps, episode_ids, n_selected, window, stop_radius, features, labels);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_lstm_train_epoch(float* params, float* exp_avg, float* exp_avg_sq, int32_t hidden,
                                      const float* features, const float* labels, const int32_t* order,
                                      int32_t n_samples, int32_t window, int32_t batch_size, float max_norm, float lr,
                                      float beta1, float beta2, float eps, float weight_decay, int32_t first_step,
                                      void* workspace, int64_t workspace_bytes, float* batch_losses, float* grad_norms,
                                      float* grad_out, void* stream) {
    PLUME_CHECK_ARG(params && exp_avg && exp_avg_sq && features && labels && workspace && batch_losses, "null pointer");
    PLUME_CHECK_ARG(hidden == kLtH, "the training kernel implements the reference's hidden_dim=32 (train_lstm.py:85)");
    PLUME_CHECK_ARG(window >= 1 && window <= kLtMaxSteps, "window must be in [1,32]");
    PLUME_CHECK_ARG(batch_size >= 1 && first_step >= 1, "batch_size and the 1-based optimiser step must be positive");
    PLUME_CHECK_ARG(workspace_bytes >= plume_lstm_train_workspace_bytes(batch_size), "workspace too small");
    PLUME_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    if (n_samples <= 0) return 0;
    cudaStream_t s = as_stream(stream);
    const size_t smem = lt_smem_bytes(window);
    static size_t configured = 0;
    if (smem > configured) {
        PLUME_CUDA(cudaFuncSetAttribute(lstm_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int64_t grid_max = (batch_size + kLtWarps - 1) / kLtWarps;
    char* ws = static_cast<char*>(workspace);
    LstmTrainArgs a;
    a.params = params;
    a.exp_avg = exp_avg;
    a.exp_avg_sq = exp_avg_sq;
    a.features = features;
    a.labels = labels;
    a.T = window;
    a.max_norm = max_norm;
    a.lr = lr;
    a.b1 = beta1;
    a.b2 = beta2;
    a.eps = eps;
    a.wd = weight_decay;
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.partial_loss = reinterpret_cast<float*>(ws + 256);
    a.partial = reinterpret_cast<float*>(ws + 256 + 4 * ((grid_max + 63) / 64 * 64));
    PLUME_CUDA(cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), s));
    int step = first_step, idx = 0;
    for (int b0 = 0; b0 < n_samples; b0 += batch_size, ++step, ++idx) {
        a.batch = n_samples - b0 < batch_size ? n_samples - b0 : batch_size;
        a.order = order ? order + b0 : nullptr;
        a.base = b0;
        a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
        a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
        a.loss_out = batch_losses + idx;
        a.grad_norm_out = grad_norms ? grad_norms + idx : nullptr;
        a.grad_out = grad_out;
        const int grid = (a.batch + kLtWarps - 1) / kLtWarps;
        lstm_train_kernel<<<grid, kLtThreads, smem, s>>>(a);
    }
    PLUME_LAUNCH_CHECK();
    return 0;
}
