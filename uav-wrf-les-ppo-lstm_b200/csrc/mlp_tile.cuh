// mlp_tile.cuh -- CTA-cooperative forward of the actor-critic MLP (P4, model.py:16-46) for a
// tile of up to 32 samples, all weights resident in shared memory (fp32: 6->256->LN->ReLU->
// 128->LN->ReLU->{5 logits, 1 value}; 70 144 FLOP/sample).  256 threads.
//
// Used by the standalone policy kernels (K3) and by the persistent rollout kernel, which
// loads the weights once and reuses them for every lockstep iteration.
//
// Shared-memory plan (floats):
//   W1t [6][256]      feature.0.weight transposed (k-major)           1536
//   P1  [3][256]      feature.0.bias, feature.1.weight, feature.1.bias  768
//   W2t [256][128]    feature.3.weight transposed (k-major)          32768
//   P2  [3][128]      feature.3.bias, feature.4.weight, feature.4.bias  384
//   Wh  [128][8]      heads: cols 0-4 actor, 5 critic, 6-7 zero       1024
//   bh  [8]                                                               8
//   x   [32][8]       input tile                                        256
//   h1  [256][32]     layer-1 activations, k-major                     8192
//   h2  [32][132]     layer-2 activations, row padded (bank-conflict free float4 rows) 4224
//   red [2][8][32]    LayerNorm partial sums                            512
//   out [32][8]       logits (0-4), value (5)                           256
//   stat[4][32]       LayerNorm statistics kept for the backward pass   128
#pragma once
#include "common.cuh"

namespace plume {

constexpr int kTileM = 32;
constexpr int kMlpThreads = 256;
constexpr int kH2Stride = 132;

struct MlpSmem {
    static constexpr int W1t = 0;
    static constexpr int P1 = W1t + 6 * 256;
    static constexpr int W2t = P1 + 3 * 256;
    static constexpr int P2 = W2t + 256 * 128;
    static constexpr int Wh = P2 + 3 * 128;
    static constexpr int bh = Wh + 128 * 8;
    static constexpr int x = bh + 8;
    static constexpr int h1 = x + kTileM * 8;
    static constexpr int h2 = h1 + 256 * kTileM;
    static constexpr int red = h2 + kTileM * kH2Stride;
    static constexpr int out = red + 2 * 8 * kTileM;
    static constexpr int stat = out + kTileM * 8;       // [4][32]: LN1 mean, LN1 rstd, LN2 rstd (training)
    static constexpr int total = stat + 4 * kTileM;     // 50056 floats = 200 224 B
};

constexpr float kLnEps = 1e-5f;    // torch.nn.LayerNorm default

// One-time: flat parameters (include/plume_b200.h layout) -> shared memory.
__device__ __forceinline__ void mlp_load_weights(float* sm, const float* __restrict__ p) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 6 * 256; i += kMlpThreads) {            // W1[o][k] -> W1t[k][o]
        const int o = i / 6, k = i - o * 6;
        sm[MlpSmem::W1t + k * 256 + o] = p[PLUME_OFF_W1 + i];
    }
    for (int i = tid; i < 256; i += kMlpThreads) {
        sm[MlpSmem::P1 + i] = p[PLUME_OFF_B1 + i];
        sm[MlpSmem::P1 + 256 + i] = p[PLUME_OFF_G1 + i];
        sm[MlpSmem::P1 + 512 + i] = p[PLUME_OFF_BE1 + i];
    }
    // W2[o][k] (row-major [128][256]) -> W2t[k][o]; 32x32 register-free transpose through
    // coalesced float4 reads along k and scalar stores (one-time cost per kernel)
    for (int i = tid; i < 128 * 64; i += kMlpThreads) {
        const int o = i >> 6, k4 = (i & 63) << 2;
        const float4 w = *reinterpret_cast<const float4*>(p + PLUME_OFF_W2 + o * 256 + k4);
        sm[MlpSmem::W2t + (k4 + 0) * 128 + o] = w.x;
        sm[MlpSmem::W2t + (k4 + 1) * 128 + o] = w.y;
        sm[MlpSmem::W2t + (k4 + 2) * 128 + o] = w.z;
        sm[MlpSmem::W2t + (k4 + 3) * 128 + o] = w.w;
    }
    for (int i = tid; i < 128; i += kMlpThreads) {
        sm[MlpSmem::P2 + i] = p[PLUME_OFF_B2 + i];
        sm[MlpSmem::P2 + 128 + i] = p[PLUME_OFF_G2 + i];
        sm[MlpSmem::P2 + 256 + i] = p[PLUME_OFF_BE2 + i];
    }
    for (int i = tid; i < 128 * 8; i += kMlpThreads) {
        const int k = i >> 3, o = i & 7;
        float w = 0.0f;
        if (o < 5) w = p[PLUME_OFF_WA + o * 128 + k];
        else if (o == 5) w = p[PLUME_OFF_WC + k];
        sm[MlpSmem::Wh + i] = w;
    }
    if (tid < 8) sm[MlpSmem::bh + tid] = tid < 5 ? p[PLUME_OFF_BA + tid] : (tid == 5 ? p[PLUME_OFF_BC] : 0.0f);
}

// Forward of the tile in sm[x] (rows >= n_valid must hold finite values, e.g. zeros).
// On return sm[out][s][0..4] = logits, [5] = value, sm[h1], sm[h2] hold the activations.
// kTrain: sm[h2] holds the *normalised* layer-2 pre-activations x_hat2 (the ReLU input is
// recomputed where needed) and sm[stat] the LayerNorm statistics, for the backward pass.
// Every thread of the 256-thread CTA must call it; ends with a __syncthreads().
template <bool kTrain = false>
__device__ __forceinline__ void mlp_forward_tile(float* sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();   // x tile visible
    // ---- layer 1: thread = (sample lane, 32 outputs of chunk `warp`) -----------------------
    {
        float xr[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) xr[k] = sm[MlpSmem::x + lane * 8 + k];
        float z[32];
        float part = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int o = warp * 32 + j;
            float a = 0.0f;
#pragma unroll
            for (int k = 0; k < 6; ++k) a = fmaf(xr[k], sm[MlpSmem::W1t + k * 256 + o], a);
            a += sm[MlpSmem::P1 + o];
            z[j] = a;
            part += a;
        }
        sm[MlpSmem::red + warp * 32 + lane] = part;
        __syncthreads();
        float mean = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) mean += sm[MlpSmem::red + w * 32 + lane];
        mean *= (1.0f / 256.0f);
        float sq = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float d = z[j] - mean;
            sq = fmaf(d, d, sq);
        }
        sm[MlpSmem::red + 256 + warp * 32 + lane] = sq;
        __syncthreads();
        float var = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) var += sm[MlpSmem::red + 256 + w * 32 + lane];
        const float rstd = 1.0f / sqrtf(var * (1.0f / 256.0f) + kLnEps);
        if (kTrain && warp == 0) {
            sm[MlpSmem::stat + lane] = mean;
            sm[MlpSmem::stat + kTileM + lane] = rstd;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int o = warp * 32 + j;
            const float y = (z[j] - mean) * rstd * sm[MlpSmem::P1 + 256 + o] + sm[MlpSmem::P1 + 512 + o];
            sm[MlpSmem::h1 + o * kTileM + lane] = fmaxf(y, 0.0f);
        }
    }
    __syncthreads();
    // ---- layer 2: thread = (4 samples sg, 4 outputs og), K = 256 ----------------------------
    {
        const int sg = tid & 7, og = tid >> 3;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
        const float* h1p = sm + MlpSmem::h1 + 4 * sg;
        const float* w2p = sm + MlpSmem::W2t + 4 * og;
#pragma unroll 8
        for (int k = 0; k < 256; ++k) {
            const float4 h = *reinterpret_cast<const float4*>(h1p + k * kTileM);
            const float4 w = *reinterpret_cast<const float4*>(w2p + k * 128);
            const float hv[4] = {h.x, h.y, h.z, h.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(hv[i], wv[j], acc[i][j]);
        }
        float part[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            part[i] = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] += sm[MlpSmem::P2 + 4 * og + j];
                part[i] += acc[i][j];
            }
            part[i] += __shfl_xor_sync(0xffffffffu, part[i], 8);
            part[i] += __shfl_xor_sync(0xffffffffu, part[i], 16);
        }
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sm[MlpSmem::red + warp * 32 + 4 * sg + i] = part[i];
        }
        __syncthreads();
        float mean[4], sq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) m += sm[MlpSmem::red + w * 32 + 4 * sg + i];
            mean[i] = m * (1.0f / 128.0f);
            sq[i] = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = acc[i][j] - mean[i];
                sq[i] = fmaf(d, d, sq[i]);
            }
            sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], 8);
            sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], 16);
        }
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sm[MlpSmem::red + 256 + warp * 32 + 4 * sg + i] = sq[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sm[MlpSmem::red + 256 + w * 32 + 4 * sg + i];
            const float rstd = 1.0f / sqrtf(v * (1.0f / 128.0f) + kLnEps);
            if (kTrain && og == 0) sm[MlpSmem::stat + 2 * kTileM + 4 * sg + i] = rstd;
            float4 y;
            float* yp = reinterpret_cast<float*>(&y);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = 4 * og + j;
                const float xh = (acc[i][j] - mean[i]) * rstd;
                yp[j] = kTrain ? xh : fmaxf(xh * sm[MlpSmem::P2 + 128 + o] + sm[MlpSmem::P2 + 256 + o], 0.0f);
            }
            *reinterpret_cast<float4*>(sm + MlpSmem::h2 + (4 * sg + i) * kH2Stride + 4 * og) = y;
        }
    }
    __syncthreads();
    // ---- heads: thread = (sample lane, output warp<6), K = 128 ---------------------------------
    if (warp < 6) {
        float a = 0.0f;
        const float* hp = sm + MlpSmem::h2 + lane * kH2Stride;
        const float* wp = sm + MlpSmem::Wh + warp;
#pragma unroll 8
        for (int k = 0; k < 128; k += 4) {
            float4 h = *reinterpret_cast<const float4*>(hp + k);
            if (kTrain) {
                const float* g2 = sm + MlpSmem::P2 + 128 + k;
                const float* be2 = sm + MlpSmem::P2 + 256 + k;
                h.x = fmaxf(fmaf(h.x, g2[0], be2[0]), 0.0f);
                h.y = fmaxf(fmaf(h.y, g2[1], be2[1]), 0.0f);
                h.z = fmaxf(fmaf(h.z, g2[2], be2[2]), 0.0f);
                h.w = fmaxf(fmaf(h.w, g2[3], be2[3]), 0.0f);
            }
            a = fmaf(h.x, wp[(k + 0) * 8], a);
            a = fmaf(h.y, wp[(k + 1) * 8], a);
            a = fmaf(h.z, wp[(k + 2) * 8], a);
            a = fmaf(h.w, wp[(k + 3) * 8], a);
        }
        sm[MlpSmem::out + lane * 8 + warp] = a + sm[MlpSmem::bh + warp];
    }
    __syncthreads();
}

// softmax over the 5 logits of sample s (model.py:44)
__device__ __forceinline__ void softmax5(const float* logits, float* probs) {
    float m = logits[0];
#pragma unroll
    for (int k = 1; k < 5; ++k) m = fmaxf(m, logits[k]);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        probs[k] = expf(logits[k] - m);
        s += probs[k];
    }
    const float inv = 1.0f / s;
#pragma unroll
    for (int k = 0; k < 5; ++k) probs[k] *= inv;
}

// Categorical(probs): renormalise, inverse-CDF draw with uniform u (or argmax / forced action),
// log_prob = log(clamp(p, eps, 1-eps))  (train_ppo2.0.py:161-162,185; torch Categorical).
__device__ __forceinline__ int categorical_pick(const float* probs, float u, bool greedy, int forced, float& logp) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 5; ++k) s += probs[k];
    float pn[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) pn[k] = probs[k] / s;
    int a = 4;
    if (forced >= 0) {
        a = forced;
    } else if (greedy) {
        a = 0;
        float best = probs[0];
#pragma unroll
        for (int k = 1; k < 5; ++k)
            if (probs[k] > best) {
                best = probs[k];
                a = k;
            }
    } else {
        float c = 0.0f;
        bool found = false;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            c += pn[k];
            if (!found && u < c) {
                a = k;
                found = true;
            }
        }
    }
    const float eps = 1.1920928955078125e-07f;
    float pa = pn[0];
#pragma unroll
    for (int k = 1; k < 5; ++k)
        if (a == k) pa = pn[k];
    logp = logf(fminf(fmaxf(pa, eps), 1.0f - eps));
    return a;
}

}  // namespace plume
