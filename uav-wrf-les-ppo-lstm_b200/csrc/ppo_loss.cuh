// ppo_loss.cuh -- per-sample PPO loss and its gradient w.r.t. the 5 logits and the value
// (train_ppo2.0.py:63-82), shared by the CUDA-core and the tensor-core update kernels.
#pragma once
#include "mlp_tile.cuh"

namespace plume {

// arguments of one minibatch gradient (both the CUDA-core and the tensor-core kernels)
struct PpoArgs {
    plume_ppo_batch batch;
    const long long* perm;
    unsigned long long perm_seed;
    int epoch;
    long long mb_start, mb_size;
    float inv_global;         // 1 / mb_size_global
    float clip_eps, entropy_beta;
    float* grads;
    double* loss_out;
    int32_t* nan_flag;
    float* ws_dz2;            // [mb_size][128]
    float* ws_x;              // [mb_size][8]
    float* ws_stat;           // [mb_size][2]  LN1 mean, rstd
};

// tensor-core path (ppo_tc_kernels.cu)
int64_t ppo_tc_workspace_bytes();
// zero_grads: the operand-preparation kernel also clears a.grads (the optimiser loop's memset folded into it)
int launch_ppo_tc(const float* params, const PpoArgs& a, void* workspace, cudaStream_t s, bool zero_grads = false);
int launch_ppo_pack(const plume_ppo_batch& b, float* packed, cudaStream_t s);

struct SampleLoss {
    float dout[6];      // d total / d logits[0..4], d total / d value, already divided by the global batch
    float pol, val, ent;
    bool nan;
};

// The per-sample loss in two independent halves (each starts from the softmax of the 5 logits), so that the
// tensor-core kernel can run them in two warps of the same scheduler instead of one long instruction stream.

// Clipped surrogate + clipped value loss (train_ppo2.0.py:63-77).  dpol[k] = d pol / d logits[k], dv = d val / d value
// (not yet divided by the batch size).
struct PolicyValuePart {
    float dpol[5], dv, pol, val;
};
__device__ __forceinline__ PolicyValuePart ppo_policy_value_part(const float* p, float v, int act, float adv, float ret,
                                                                 float vold, float lpold, float clip_eps) {
    PolicyValuePart r;
    float S = 0.0f;
#pragma unroll
    for (int k = 0; k < 5; ++k) S += p[k];
    float pa = p[0];
#pragma unroll
    for (int k = 1; k < 5; ++k)
        if (act == k) pa = p[k];
    const float q = pa / S;
    const float eps = 1.1920928955078125e-07f;
    const bool q_inside = (q >= eps) && (q <= 1.0f - eps);
    const float lp = logf(fminf(fmaxf(q, eps), 1.0f - eps));          // :63-64
    const float ratio = expf(lp - lpold);                             // :67
    const float lo = 1.0f - clip_eps, hi = 1.0f + clip_eps;
    const float s1 = ratio * adv;
    const float s2 = fminf(fmaxf(ratio, lo), hi) * adv;               // :68-69
    const bool inside = (ratio >= lo) && (ratio <= hi);
    float dratio;                                                     // d(-min(s1,s2))/d ratio
    if (inside) dratio = -adv;
    else if (s1 < s2) dratio = -adv;
    else if (s1 > s2) dratio = 0.0f;
    else dratio = -0.5f * adv;
    const float dlp = q_inside ? dratio * ratio : 0.0f;
    r.pol = -fminf(s1, s2);                                           // :70
    // value loss :73-77
    const float dvv = v - vold;
    const bool v_inside = (dvv >= -clip_eps) && (dvv <= clip_eps);
    const float vclip = vold + fminf(fmaxf(dvv, -clip_eps), clip_eps);
    const float e1 = (v - ret) * (v - ret), e2 = (vclip - ret) * (vclip - ret);
    if (e1 > e2) r.dv = (v - ret);
    else if (e2 > e1) r.dv = v_inside ? (vclip - ret) : 0.0f;
    else r.dv = 0.5f * (v - ret) + (v_inside ? 0.5f * (vclip - ret) : 0.0f);
    r.val = 0.5f * fmaxf(e1, e2);
    // log q = log p_a - log S: the renormalisation of Categorical(probs) contributes -(1 - S)/S p_k (S = 1 up to rounding)
    const float renorm = (1.0f - S) / S;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float onehot = (act == k) ? 1.0f : 0.0f;
        r.dpol[k] = dlp * ((onehot - p[k]) - p[k] * renorm);
    }
    return r;
}

// Entropy bonus -mean(sum p log(p + 1e-8)) (:80).  dent[k] = d(-beta ent) / d logits[k] (not yet divided by the batch).
struct EntropyPart {
    float dent[5], ent;
};
__device__ __forceinline__ EntropyPart ppo_entropy_part(const float* p, float entropy_beta) {
    EntropyPart r;
    float ent = 0.0f, gbar = 0.0f, gk[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float lg = logf(p[k] + 1e-8f);
        ent -= p[k] * lg;
        gk[k] = lg + p[k] / (p[k] + 1e-8f);     // d/dp_k sum p log(p+1e-8)
        gbar += gk[k] * p[k];
    }
    r.ent = ent;
#pragma unroll
    for (int k = 0; k < 5; ++k) r.dent[k] = entropy_beta * p[k] * (gk[k] - gbar);
    return r;
}

// o[0..4] logits, o[5] value; total = pol + val - beta*ent (:82), all means over the batch.
__device__ __forceinline__ SampleLoss ppo_sample_loss(const float* o, int act, float adv, float ret, float vold,
                                                      float lpold, float clip_eps, float entropy_beta,
                                                      float inv_global) {
    SampleLoss r;
    r.nan = false;
#pragma unroll
    for (int k = 0; k < 5; ++k) r.nan |= isnan(o[k]);                    // :57-61
    float p[5];
    softmax5(o, p);
    const PolicyValuePart pv = ppo_policy_value_part(p, o[5], act, adv, ret, vold, lpold, clip_eps);
    const EntropyPart en = ppo_entropy_part(p, entropy_beta);
    r.pol = pv.pol;
    r.val = pv.val;
    r.ent = en.ent;
#pragma unroll
    for (int k = 0; k < 5; ++k) r.dout[k] = pv.dpol[k] * inv_global + en.dent[k] * inv_global;
    r.dout[5] = pv.dv * inv_global;
    return r;
}

// warp "transpose-reduce": v[32] per lane -> returns sum over lanes of v[lane index]
__device__ __forceinline__ float warp_reduce_by_index(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float send = up ? v[i] : v[i + step];
            const float keep = up ? v[i + step] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return v[0];
}

}  // namespace plume
