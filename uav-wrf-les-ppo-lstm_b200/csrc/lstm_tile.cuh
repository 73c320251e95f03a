// lstm_tile.cuh -- CTA-cooperative single-layer LSTM (input size 1) over a window, from zero
// state, for a tile of 32 sequences with the recurrent weights resident in shared memory
// (P4L: PeakAndStopPredictor, PPOV2.1/evaluate_with_lstm.py:11-27,73-80).  256 threads.
//
// Gate order i,f,g,o (torch.nn.LSTM).  FLOP per cell step = 2*4H*(1+H).
//
// Shared-memory plan (floats), H = hidden size (32 or 64):
//   Wt  [H][H][4]   w_hh transposed & gate-interleaved: Wt[k][j][g] = w_hh[g*H+j][k]
//   wx  [H][4]      w_ih[g*H+j][0]
//   bs  [H][4]      b_ih + b_hh
//   hd  [2][H]      head weights: fc_peak.weight, fc_stop.0.weight ; hb [2] biases
//   h   [2][H][32]  hidden state, double buffered, k-major (sample fastest)
//   xs  [32][32]    window values, step-major: xs[t][s]
#pragma once
#include "common.cuh"

namespace plume {

constexpr int kLstmMaxSteps = 32;

template <int H>
struct LstmSmem {
    static constexpr int Wt = 0;
    static constexpr int wx = Wt + H * H * 4;
    static constexpr int bs = wx + H * 4;
    static constexpr int hd = bs + H * 4;
    static constexpr int hb = hd + 2 * H;
    static constexpr int h = hb + 4;
    static constexpr int xs = h + 2 * H * 32;
    static constexpr int total = xs + kLstmMaxSteps * 32;     // H=32: 7428 floats; H=64: 22660 floats
};

struct LstmWeights {
    const float *w_ih, *w_hh, *b_ih, *b_hh, *w_peak, *b_peak, *w_stop, *b_stop;
};

template <int H>
__device__ __forceinline__ void lstm_load_weights(float* sm, const LstmWeights& w) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * H * H; i += 256) {          // w_hh[row = g*H+j][k], coalesced read along k
        const int row = i / H, k = i - row * H;
        const int g = row / H, j = row - g * H;
        sm[LstmSmem<H>::Wt + (k * H + j) * 4 + g] = w.w_hh[i];
    }
    for (int i = tid; i < 4 * H; i += 256) {
        const int g = i / H, j = i - g * H;
        sm[LstmSmem<H>::wx + j * 4 + g] = w.w_ih[i];
        sm[LstmSmem<H>::bs + j * 4 + g] = w.b_ih[i] + w.b_hh[i];
    }
    for (int i = tid; i < H; i += 256) {
        sm[LstmSmem<H>::hd + i] = w.w_peak[i];
        sm[LstmSmem<H>::hd + H + i] = w.w_stop[i];
    }
    if (tid == 0) {
        sm[LstmSmem<H>::hb] = w.b_peak[0];
        sm[LstmSmem<H>::hb + 1] = w.b_stop[0];
    }
}

// Gate activations: ex2.approx-based exp and the fast divide (both ~2 ulp) -- relative error
// ~3e-7, well inside the fp32 rel 1e-5 parity budget, at a quarter of the instructions of
// expf()/IEEE division/tanhf() (ncu r1: 30 % of the rollout kernel's samples were in those).
__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_acc(float x) {
    const float t = __expf(-2.0f * fabsf(x));
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// Runs `steps` cell steps over sm[xs][t][s] from zero state.  Afterwards the final hidden
// state is in sm[h][steps & 1].  All 256 threads must call; ends with __syncthreads().
template <int H>
__device__ __forceinline__ void lstm_window_tile(float* sm, int steps) {
    constexpr int U = H / 32;                     // hidden units per thread
    const int tid = threadIdx.x, sg = tid & 7, j0 = tid >> 3;
    float cst[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) cst[u][i] = 0.0f;
    // h_0 = 0
    for (int i = tid; i < H * 32; i += 256) sm[LstmSmem<H>::h + i] = 0.0f;
    __syncthreads();
    for (int t = 0; t < steps; ++t) {
        const float* hin = sm + LstmSmem<H>::h + (t & 1) * H * 32 + 4 * sg;
        float* hout = sm + LstmSmem<H>::h + ((t + 1) & 1) * H * 32 + 4 * sg;
        const float4 xv = *reinterpret_cast<const float4*>(sm + LstmSmem<H>::xs + t * 32 + 4 * sg);
        const float x[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + 32 * u;
            const float4 wxv = *reinterpret_cast<const float4*>(sm + LstmSmem<H>::wx + j * 4);
            const float4 bv = *reinterpret_cast<const float4*>(sm + LstmSmem<H>::bs + j * 4);
            float acc[4][4];     // [sample][gate]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(x[i], wxv.x, bv.x);
                acc[i][1] = fmaf(x[i], wxv.y, bv.y);
                acc[i][2] = fmaf(x[i], wxv.z, bv.z);
                acc[i][3] = fmaf(x[i], wxv.w, bv.w);
            }
            const float* wp = sm + LstmSmem<H>::Wt + j * 4;
#pragma unroll 8
            for (int k = 0; k < H; ++k) {
                const float4 hv = *reinterpret_cast<const float4*>(hin + k * 32);
                const float4 wv = *reinterpret_cast<const float4*>(wp + k * H * 4);
                const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(hh[i], wv.x, acc[i][0]);
                    acc[i][1] = fmaf(hh[i], wv.y, acc[i][1]);
                    acc[i][2] = fmaf(hh[i], wv.z, acc[i][2]);
                    acc[i][3] = fmaf(hh[i], wv.w, acc[i][3]);
                }
            }
            float hn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float ig = sigmoidf_acc(acc[i][0]);
                const float fg = sigmoidf_acc(acc[i][1]);
                const float gg = tanhf_acc(acc[i][2]);
                const float og = sigmoidf_acc(acc[i][3]);
                cst[u][i] = fmaf(fg, cst[u][i], ig * gg);
                hn[i] = og * tanhf_acc(cst[u][i]);
            }
            *reinterpret_cast<float4*>(hout + j * 32) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        }
        __syncthreads();
    }
}

// fc_peak / fc_stop on the final hidden state: threads 0..31 -> peak of sample tid,
// threads 32..63 -> stop probability of sample tid-32.  Returns the value for those threads.
template <int H>
__device__ __forceinline__ float lstm_heads(const float* sm, int steps) {
    const int tid = threadIdx.x;
    if (tid >= 64) return 0.0f;
    const int which = tid >> 5, s = tid & 31;
    const float* hf = sm + LstmSmem<H>::h + (steps & 1) * H * 32 + s;
    const float* w = sm + LstmSmem<H>::hd + which * H;
    float a = 0.0f;
#pragma unroll 8
    for (int k = 0; k < H; ++k) a = fmaf(hf[k * 32], w[k], a);
    a += sm[LstmSmem<H>::hb + which];
    return which ? sigmoidf_acc(a) : a;
}

// ---- tensor-core deferred stop head (lstm_tc_kernels.cu) ---------------------------------------------
struct LtArgs {
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
    const float *w_ih, *w_hh, *b_ih, *b_hh, *w_peak, *b_peak, *w_stop, *b_stop;
};
bool stop_head_segment_tc_supports(int hidden);
int launch_stop_head_segment_tc(const LtArgs& a, int hidden, cudaStream_t s);
// hidden 128 / 256: gate weights streamed from L2 (lstm_tc_stream.cu)
bool stop_head_segment_stream_supports(int hidden);
int launch_stop_head_segment_stream(const LtArgs& a, int hidden, cudaStream_t s);

// ---- P4t trend features (calculate_dynamic_label, PPOV2.1/model.py:113-127) ------------------------
// m = mean of the last three np.gradient values of the window; dist = ||pos[-1] - src||
__device__ __forceinline__ void trend_finish(double m, double c_last, double dist, double conc_peak, float* out) {
    const double trend = tanh(m / 5.0);
    const double dist_score = exp(-dist / 50.0);
    double cs = c_last / conc_peak;
    cs = cs < 0.0 ? 0.0 : (cs > 1.0 ? 1.0 : cs);
    double label = 0.4 * dist_score + 0.3 * (trend + 1.0) / 2.0 + 0.3 * cs;
    label = label < 0.01 ? 0.01 : (label > 0.99 ? 0.99 : label);
    out[0] = (float)label;
    out[1] = (float)trend;
    out[2] = (float)dist_score;
    out[3] = (float)cs;
}

// np.linalg.norm(pos - src) in float64, not contracted (same arithmetic as env:155)
__device__ __forceinline__ double source_distance(double px, double py, double sx, double sy) {
    const double ddx = dsub(px, sx), ddy = dsub(py, sy);
    return dsqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy)));
}

// window of >= 4 samples: only the last four matter (central differences inside, one-sided at the end)
__device__ __forceinline__ void trend_from_last4(double c4, double c3, double c2, double c1, double dist,
                                                 double conc_peak, float* out) {
    const double g0 = (c2 - c4) / 2.0, g1 = (c1 - c3) / 2.0, g2 = c1 - c2;
    trend_finish(((g0 + g1) + g2) / 3.0, c1, dist, conc_peak, out);
}

__device__ __forceinline__ void trend_features(const float* conc, int W, double px, double py, double sx, double sy,
                                               double conc_peak, float* out) {
    const double dist = source_distance(px, py, sx, sy);
    if (W >= 4) {
        trend_from_last4((double)conc[W - 4], (double)conc[W - 3], (double)conc[W - 2], (double)conc[W - 1], dist,
                         conc_peak, out);
        return;
    }
    // W in {2,3}: np.gradient is one-sided at both ends
    double m;
    if (W == 2) {
        m = (double)conc[1] - (double)conc[0];
    } else {
        const double g0 = (double)conc[1] - (double)conc[0], g1 = ((double)conc[2] - (double)conc[0]) / 2.0,
                     g2 = (double)conc[2] - (double)conc[1];
        m = ((g0 + g1) + g2) / 3.0;
    }
    trend_finish(m, (double)conc[W - 1], dist, conc_peak, out);
}

}  // namespace plume
