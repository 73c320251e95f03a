// ppo_kernels.cu -- K6: the PPO minibatch gradient (train_ppo2.0.py:42-85): forward of the
// actor-critic, clipped-surrogate + clipped value + entropy loss, and the full backward pass,
// fused per 32-sample tile with all weights resident in shared memory.
//
//   kernel A (ppo_fwd_bwd_kernel): gather the minibatch samples (explicit permutation or the
//     stateless Feistel bijection), forward (mlp_forward_tile<true>), loss and its gradient,
//     backward through heads / LN2 / layer 2 / LN1; accumulates every gradient except
//     d feature.3.weight in registers across tiles (flushed with one atomicAdd per element
//     per CTA) and writes dz2 [B][128] + the gathered inputs [B][8] + LN1 statistics [B][2]
//     to the workspace.
//   kernel B (ppo_wgrad2_kernel): d feature.3.weight = dz2^T . h1 (128 x 256 x B), h1
//     recomputed from the gathered inputs; split over samples, 128 accumulators per thread.
//
// Algorithmic bytes per sample (DESIGN.md "K6"): gather obs 24 + action 4 + old logp 4 + adv 4 +
// ret 4 + old value 4 = 44 B read; workspace dz2 512 + x 32 + stats 8 = 552 B written by A and
// read by B.  FLOP per sample: forward 70 144, backward ~140 288.
#include <cstdlib>

#include "ppo_loss.cuh"

namespace plume {

constexpr int kWsFloatsPerSample = 128 + 8 + 2;

__global__ void __launch_bounds__(kMlpThreads, 1) ppo_fwd_bwd_kernel(const float* __restrict__ params, PpoArgs a) {
    extern __shared__ __align__(16) float sm[];
    // per-sample scalars of the tile, after the forward regions
    float* s_adv = sm + MlpSmem::total;          // [32]
    float* s_ret = s_adv + 32;
    float* s_vold = s_ret + 32;
    float* s_lpold = s_vold + 32;
    int* s_act = reinterpret_cast<int*>(s_lpold + 32);
    mlp_load_weights(sm, params);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // persistent gradient accumulators (meaning depends on the phase that owns them)
    float g_w1[6] = {0, 0, 0, 0, 0, 0};        // d feature.0.weight[o = tid][0..5]
    float g_b1 = 0.0f, g_g1 = 0.0f, g_be1 = 0.0f;   // d feature.0.bias / feature.1.weight / feature.1.bias [o = tid]
    float g_b2 = 0.0f, g_g2 = 0.0f, g_be2 = 0.0f;   // [o = tid & 127], half tid >> 7 of the samples
    float g_wh[6] = {0, 0, 0, 0, 0, 0};        // d heads[j][o = tid & 127]
    float g_bh = 0.0f;                         // d head bias j = tid (tid < 6)
    double l_tot = 0.0, l_pol = 0.0, l_val = 0.0, l_ent = 0.0;

    const long long tiles = (a.mb_size + kTileM - 1) / kTileM;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long base = tile * kTileM;
        const int n_valid = (int)((a.mb_size - base) < kTileM ? (a.mb_size - base) : kTileM);
        __syncthreads();
        // ---- gather ------------------------------------------------------------------------
        if (tid < kTileM) {
            long long idx = -1;
            if (tid < n_valid) {
                const long long pos = a.mb_start + base + tid;
                idx = a.perm ? a.perm[pos]
                             : (long long)feistel_permute((uint64_t)pos, (uint64_t)a.batch.total, a.perm_seed,
                                                          (uint32_t)a.epoch);
            }
            float xv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (idx >= 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) xv[k] = a.batch.obs[idx * 6 + k];
                s_adv[tid] = a.batch.advantages[idx];
                s_ret[tid] = a.batch.returns[idx];
                s_vold[tid] = a.batch.old_values[idx];
                s_lpold[tid] = a.batch.old_log_probs[idx];
                s_act[tid] = a.batch.actions[idx];
            } else {
                s_adv[tid] = s_ret[tid] = s_vold[tid] = s_lpold[tid] = 0.0f;
                s_act[tid] = 0;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) sm[MlpSmem::x + tid * 8 + k] = xv[k];
            if (idx >= 0) {
                float4* wx = reinterpret_cast<float4*>(a.ws_x + (base + tid) * 8);
                wx[0] = make_float4(xv[0], xv[1], xv[2], xv[3]);
                wx[1] = make_float4(xv[4], xv[5], 0.0f, 0.0f);
            }
        }
        mlp_forward_tile<true>(sm);
        // ---- loss and d(loss)/d(logits, value): one thread per sample ---------------------------
        if (tid < kTileM) {
            float dout[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (tid < n_valid) {
                const SampleLoss L = ppo_sample_loss(sm + MlpSmem::out + tid * 8, s_act[tid], s_adv[tid], s_ret[tid],
                                                     s_vold[tid], s_lpold[tid], a.clip_eps, a.entropy_beta,
                                                     a.inv_global);
                if (L.nan) atomicExch(a.nan_flag, 1);                            // :57-61
#pragma unroll
                for (int k = 0; k < 6; ++k) dout[k] = L.dout[k];
                l_pol += (double)L.pol;
                l_val += (double)L.val;
                l_ent += (double)L.ent;
                l_tot += (double)L.pol + (double)L.val - (double)a.entropy_beta * (double)L.ent;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) sm[MlpSmem::out + tid * 8 + k] = dout[k];
        }
        __syncthreads();
        if (tid < 6) {
            float sacc = 0.0f;
            for (int s = 0; s < kTileM; ++s) sacc += sm[MlpSmem::out + s * 8 + tid];
            g_bh += sacc;
        }
        // ---- heads + LN2 backward: thread = (output o, half of the samples) ------------------
        {
            const int o = tid & 127, half = tid >> 7;
            const float g2 = sm[MlpSmem::P2 + 128 + o], be2 = sm[MlpSmem::P2 + 256 + o];
            float wrow[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) wrow[j] = sm[MlpSmem::Wh + o * 8 + j];
            float xh[16], dxh[16];
            float* xh2 = sm + MlpSmem::h2;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = half * 16 + q;
                const float x_hat = xh2[s * kH2Stride + o];
                const float y = fmaf(x_hat, g2, be2);
                const float h2v = fmaxf(y, 0.0f);
                const float* d = sm + MlpSmem::out + s * 8;
                float dh = 0.0f;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    dh = fmaf(d[j], wrow[j], dh);
                    g_wh[j] = fmaf(d[j], h2v, g_wh[j]);
                }
                const float dy = (y > 0.0f) ? dh : 0.0f;
                g_g2 = fmaf(dy, x_hat, g_g2);
                g_be2 += dy;
                xh[q] = x_hat;
                dxh[q] = dy * g2;
            }
            // means over the 128 outputs of dxh and dxh*xh, per sample: warp shuffle + smem
            float r1[16], r2[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                r1[q] = dxh[q];
                r2[q] = dxh[q] * xh[q];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    r1[q] += __shfl_xor_sync(0xffffffffu, r1[q], off);
                    r2[q] += __shfl_xor_sync(0xffffffffu, r2[q], off);
                }
            }
            // 4 warps per half: partials in red[warp_in_half][...]
            const int wih = warp & 3;
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    sm[MlpSmem::red + wih * 32 + half * 16 + q] = r1[q];
                    sm[MlpSmem::red + 256 + wih * 32 + half * 16 + q] = r2[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int s = half * 16 + q;
                float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    m1 += sm[MlpSmem::red + w * 32 + s];
                    m2 += sm[MlpSmem::red + 256 + w * 32 + s];
                }
                m1 *= (1.0f / 128.0f);
                m2 *= (1.0f / 128.0f);
                const float rstd = sm[MlpSmem::stat + 2 * kTileM + s];
                const float dz = rstd * (dxh[q] - m1 - xh[q] * m2);
                g_b2 += dz;
                xh2[s * kH2Stride + o] = dz;                 // dz2 replaces x_hat2 in shared memory
                if (s < n_valid) a.ws_dz2[(base + s) * 128 + o] = dz;
            }
        }
        if (tid < kTileM && tid < n_valid) {
            a.ws_stat[(base + tid) * 2 + 0] = sm[MlpSmem::stat + tid];
            a.ws_stat[(base + tid) * 2 + 1] = sm[MlpSmem::stat + kTileM + tid];
        }
        __syncthreads();
        // ---- dh1 = dz2 . W2: thread = (samples sg+8i, 8 inputs k) -> written over h1 ---------
        {
            const int sg = tid & 7, kg = tid >> 3;
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
            const float* dzp = sm + MlpSmem::h2 + sg * kH2Stride;
            const float* wp = sm + MlpSmem::W2t + (8 * kg) * 128;
#pragma unroll 2
            for (int o = 0; o < 128; o += 4) {
                float4 dz[4], w[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) dz[i] = *reinterpret_cast<const float4*>(dzp + (8 * i) * kH2Stride + o);
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const float4*>(wp + j * 128 + o);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[i][j] = fmaf(dz[i].x, w[j].x, acc[i][j]);
                        acc[i][j] = fmaf(dz[i].y, w[j].y, acc[i][j]);
                        acc[i][j] = fmaf(dz[i].z, w[j].z, acc[i][j]);
                        acc[i][j] = fmaf(dz[i].w, w[j].w, acc[i][j]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) sm[MlpSmem::h1 + (8 * kg + j) * kTileM + sg + 8 * i] = acc[i][j];
        }
        __syncthreads();
        // ---- LN1 backward + d feature.0.*: thread = (sample lane, 32 outputs of chunk warp) ---
        {
            float xr[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) xr[k] = sm[MlpSmem::x + lane * 8 + k];
            const float mean = sm[MlpSmem::stat + lane], rstd = sm[MlpSmem::stat + kTileM + lane];
            float xh[32], dyv[32];
            float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int o = warp * 32 + j;
                float z = 0.0f;
#pragma unroll
                for (int k = 0; k < 6; ++k) z = fmaf(xr[k], sm[MlpSmem::W1t + k * 256 + o], z);
                z += sm[MlpSmem::P1 + o];
                const float g1 = sm[MlpSmem::P1 + 256 + o];
                const float x_hat = (z - mean) * rstd;
                const float y = fmaf(x_hat, g1, sm[MlpSmem::P1 + 512 + o]);
                const float dy = (y > 0.0f) ? sm[MlpSmem::h1 + o * kTileM + lane] : 0.0f;
                xh[j] = x_hat;
                dyv[j] = dy;
                const float dxh = dy * g1;
                p1 += dxh;
                p2 = fmaf(dxh, x_hat, p2);
            }
            sm[MlpSmem::red + warp * 32 + lane] = p1;
            sm[MlpSmem::red + 256 + warp * 32 + lane] = p2;
            __syncthreads();
            float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                m1 += sm[MlpSmem::red + w * 32 + lane];
                m2 += sm[MlpSmem::red + 256 + w * 32 + lane];
            }
            m1 *= (1.0f / 256.0f);
            m2 *= (1.0f / 256.0f);
            // reductions over the samples (lanes); lane j ends up owning output warp*32+j = tid
            float t[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = dyv[j] * xh[j];
            g_g1 += warp_reduce_by_index(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = dyv[j];
            g_be1 += warp_reduce_by_index(t, lane);
            // dz1 (kept in dyv)
#pragma unroll
            for (int j = 0; j < 32; ++j)
                dyv[j] = rstd * (dyv[j] * sm[MlpSmem::P1 + 256 + warp * 32 + j] - m1 - xh[j] * m2);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = dyv[j];
            g_b1 += warp_reduce_by_index(t, lane);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
#pragma unroll
                for (int j = 0; j < 32; ++j) t[j] = dyv[j] * xr[k];
                g_w1[k] += warp_reduce_by_index(t, lane);
            }
        }
    }
    // ---- flush the CTA's partial gradients --------------------------------------------------------
    float* g = a.grads;
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(g + PLUME_OFF_W1 + tid * 6 + k, g_w1[k]);
    atomicAdd(g + PLUME_OFF_B1 + tid, g_b1);
    atomicAdd(g + PLUME_OFF_G1 + tid, g_g1);
    atomicAdd(g + PLUME_OFF_BE1 + tid, g_be1);
    {
        const int o = tid & 127;
        atomicAdd(g + PLUME_OFF_B2 + o, g_b2);
        atomicAdd(g + PLUME_OFF_G2 + o, g_g2);
        atomicAdd(g + PLUME_OFF_BE2 + o, g_be2);
#pragma unroll
        for (int j = 0; j < 5; ++j) atomicAdd(g + PLUME_OFF_WA + j * 128 + o, g_wh[j]);
        atomicAdd(g + PLUME_OFF_WC + o, g_wh[5]);
    }
    if (tid < 5) atomicAdd(g + PLUME_OFF_BA + tid, g_bh);
    if (tid == 5) atomicAdd(g + PLUME_OFF_BC, g_bh);
    if (tid < kTileM) {
        // loss sums: warp 0 holds them
        for (int off = 16; off > 0; off >>= 1) {
            l_tot += __shfl_xor_sync(0xffffffffu, l_tot, off);
            l_pol += __shfl_xor_sync(0xffffffffu, l_pol, off);
            l_val += __shfl_xor_sync(0xffffffffu, l_val, off);
            l_ent += __shfl_xor_sync(0xffffffffu, l_ent, off);
        }
        if (tid == 0) {
            const double inv = (double)a.inv_global;
            atomicAdd(a.loss_out + 0, l_tot * inv);
            atomicAdd(a.loss_out + 1, l_pol * inv);
            atomicAdd(a.loss_out + 2, l_val * inv);
            atomicAdd(a.loss_out + 3, l_ent * inv);
        }
    }
}

// ---- kernel B: d feature.3.weight[o][k] = sum_s dz2[s][o] * h1[s][k] ---------------------------------
constexpr int kH1Stride = 260;   // [32][260]: conflict-free float4 rows

__global__ void __launch_bounds__(256, 1) ppo_wgrad2_kernel(const float* __restrict__ params, PpoArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* W1t = sm;                      // [6][256]
    float* P1 = W1t + 6 * 256;            // [3][256]
    float* dz = P1 + 3 * 256;             // [32][128]
    float* h1 = dz + 32 * 128;            // [32][260]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 6 * 256; i += 256) {
        const int o = i / 6, k = i - o * 6;
        W1t[k * 256 + o] = params[PLUME_OFF_W1 + i];
    }
    for (int i = tid; i < 256; i += 256) {
        P1[i] = params[PLUME_OFF_B1 + i];
        P1[256 + i] = params[PLUME_OFF_G1 + i];
        P1[512 + i] = params[PLUME_OFF_BE1 + i];
    }
    const int og = tid & 15, kg = tid >> 4;       // 8 outputs o, 16 inputs k per thread
    float acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.0f;

    const long long tiles = (a.mb_size + kTileM - 1) / kTileM;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long base = tile * kTileM;
        const int n_valid = (int)((a.mb_size - base) < kTileM ? (a.mb_size - base) : kTileM);
        __syncthreads();
        // dz2 tile (zero rows beyond the minibatch)
        for (int i = tid; i < 32 * 32; i += 256) {
            const int s = i >> 5, c = (i & 31) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < n_valid) v = *reinterpret_cast<const float4*>(a.ws_dz2 + (base + s) * 128 + c);
            *reinterpret_cast<float4*>(dz + s * 128 + c) = v;
        }
        // recompute h1 = relu(LN1(W1 x + b1)) for the tile: thread = (sample lane, chunk warp)
        {
            float xr[6] = {0, 0, 0, 0, 0, 0};
            float mean = 0.0f, rstd = 0.0f;
            if (lane < n_valid) {
                const float4 x0 = *reinterpret_cast<const float4*>(a.ws_x + (base + lane) * 8);
                const float4 x1 = *reinterpret_cast<const float4*>(a.ws_x + (base + lane) * 8 + 4);
                xr[0] = x0.x; xr[1] = x0.y; xr[2] = x0.z; xr[3] = x0.w; xr[4] = x1.x; xr[5] = x1.y;
                mean = a.ws_stat[(base + lane) * 2];
                rstd = a.ws_stat[(base + lane) * 2 + 1];
            }
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
                float hv[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int o = warp * 32 + j4 + jj;
                    float z = 0.0f;
#pragma unroll
                    for (int k = 0; k < 6; ++k) z = fmaf(xr[k], W1t[k * 256 + o], z);
                    z += P1[o];
                    hv[jj] = fmaxf(fmaf((z - mean) * rstd, P1[256 + o], P1[512 + o]), 0.0f);
                }
                *reinterpret_cast<float4*>(h1 + lane * kH1Stride + warp * 32 + j4) =
                    make_float4(hv[0], hv[1], hv[2], hv[3]);
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int s = 0; s < kTileM; ++s) {
            const float4 d0 = *reinterpret_cast<const float4*>(dz + s * 128 + 8 * og);
            const float4 d1 = *reinterpret_cast<const float4*>(dz + s * 128 + 8 * og + 4);
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            float hv[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 h = *reinterpret_cast<const float4*>(h1 + s * kH1Stride + 16 * kg + 4 * q);
                hv[4 * q] = h.x; hv[4 * q + 1] = h.y; hv[4 * q + 2] = h.z; hv[4 * q + 3] = h.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(dv[i], hv[j], acc[i][j]);
        }
    }
    float* g = a.grads + PLUME_OFF_W2;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(g + (8 * og + i) * 256 + 16 * kg + j, acc[i][j]);
}

}  // namespace plume

using namespace plume;

// Which kernels compute a minibatch is the CALLER's choice (kernel_path argument of plume_ppo_grad): the tcgen05
// kernel (ppo_tc_kernels.cu, 128-sample tiles, one CTA per SM) or the fp32-FMA kernels below (32-sample tiles, more
// CTAs for tiny batches such as the reference's 256).  PLUME_KERNEL_AUTO is a size rule, not a fallback: tcgen05 from
// kTcMinBatch samples on.
constexpr int64_t kTcMinBatch = 1024;
static bool use_tc_path(int64_t mb_size, int32_t kernel_path) {
    if (kernel_path == PLUME_KERNEL_TENSOR) return true;
    if (kernel_path == PLUME_KERNEL_SIMT) return false;
    return mb_size >= kTcMinBatch;
}

extern "C" int64_t plume_ppo_workspace_bytes(int64_t mb_size) {
    const int64_t padded = ((mb_size + kTileM - 1) / kTileM) * kTileM;
    const int64_t cuda_path = padded * kWsFloatsPerSample * (int64_t)sizeof(float) + 256;
    const int64_t tc_path = ppo_tc_workspace_bytes();
    return cuda_path > tc_path ? cuda_path : tc_path;
}

extern "C" int plume_ppo_pack(const plume_ppo_batch* batch, float* packed, void* stream) {
    PLUME_CHECK_ARG(batch && packed, "null pointer");
    PLUME_CHECK_ARG(batch->obs && batch->actions && batch->old_log_probs && batch->advantages && batch->returns &&
                        batch->old_values, "null batch pointer");
    PLUME_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(batch->obs) & 7) == 0,
                    "packed must be 16-byte aligned, obs 8-byte aligned");
    return launch_ppo_pack(*batch, packed, as_stream(stream));
}

// zero_grads: the launch clears `grads` itself before accumulating (the optimiser loop; plume_ppo_grad accumulates)
static int ppo_grad_impl(const float* params, const plume_ppo_batch* batch, const int64_t* perm, uint64_t perm_seed,
                         int32_t epoch, int64_t mb_start, int64_t mb_size, int64_t mb_size_global, float clip_eps,
                         float entropy_beta, float* grads, double* loss_out, int32_t* nan_flag, void* workspace,
                         int64_t workspace_bytes, int32_t kernel_path, void* stream, bool zero_grads) {
    PLUME_CHECK_ARG(params && batch && grads && loss_out && nan_flag && workspace, "null pointer");
    PLUME_CHECK_ARG(kernel_path >= PLUME_KERNEL_AUTO && kernel_path <= PLUME_KERNEL_SIMT, "unknown kernel_path");
    PLUME_CHECK_ARG(batch->obs && batch->actions && batch->old_log_probs && batch->advantages && batch->returns &&
                        batch->old_values, "null batch pointer");
    PLUME_CHECK_ARG(mb_start >= 0 && mb_size >= 0 && mb_start + mb_size <= batch->total, "minibatch outside [0,total)");
    PLUME_CHECK_ARG(mb_size_global >= mb_size && mb_size_global > 0, "mb_size_global must be >= mb_size");
    PLUME_CHECK_ARG(workspace_bytes >= plume_ppo_workspace_bytes(mb_size), "workspace too small");
    PLUME_CHECK_ARG((reinterpret_cast<uintptr_t>(batch->packed) & 15) == 0, "packed records must be 16-byte aligned");
    if (mb_size == 0) {
        if (zero_grads) PLUME_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * PLUME_MLP_PARAMS, as_stream(stream)));
        return 0;
    }
    static bool configured = false;
    const int smem_a = (MlpSmem::total + 7 * 32 + 64) * (int)sizeof(float);
    const int smem_b = (6 * 256 + 3 * 256 + 32 * 128 + 32 * kH1Stride) * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(ppo_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a) != cudaSuccess ||
            cudaFuncSetAttribute(ppo_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b) != cudaSuccess)
            return fail("ppo kernels: cannot reserve shared memory");
        configured = true;
    }
    PpoArgs a;
    a.batch = *batch;
    a.perm = reinterpret_cast<const long long*>(perm);
    a.perm_seed = perm_seed;
    a.epoch = epoch;
    a.mb_start = mb_start;
    a.mb_size = mb_size;
    a.inv_global = (float)(1.0 / (double)mb_size_global);
    a.clip_eps = clip_eps;
    a.entropy_beta = entropy_beta;
    a.grads = grads;
    a.loss_out = loss_out;
    a.nan_flag = nan_flag;
    a.ws_dz2 = a.ws_x = a.ws_stat = nullptr;
    if (use_tc_path(mb_size, kernel_path)) return launch_ppo_tc(params, a, workspace, as_stream(stream), zero_grads);
    if (zero_grads) PLUME_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * PLUME_MLP_PARAMS, as_stream(stream)));
    const int64_t padded = ((mb_size + kTileM - 1) / kTileM) * kTileM;
    uintptr_t wsp = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
    a.ws_dz2 = reinterpret_cast<float*>(wsp);
    a.ws_x = a.ws_dz2 + padded * 128;
    a.ws_stat = a.ws_x + padded * 8;
    const long long tiles = padded / kTileM;
    int grid = sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = (int)tiles;
    cudaStream_t s = as_stream(stream);
    ppo_fwd_bwd_kernel<<<grid, kMlpThreads, smem_a, s>>>(params, a);
    PLUME_LAUNCH_CHECK();
    ppo_wgrad2_kernel<<<grid, 256, smem_b, s>>>(params, a);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_ppo_grad(const float* params, const plume_ppo_batch* batch, const int64_t* perm,
                              uint64_t perm_seed, int32_t epoch, int64_t mb_start, int64_t mb_size,
                              int64_t mb_size_global, float clip_eps, float entropy_beta, float* grads,
                              double* loss_out, int32_t* nan_flag, void* workspace, int64_t workspace_bytes,
                              int32_t kernel_path, void* stream) {
    return ppo_grad_impl(params, batch, perm, perm_seed, epoch, mb_start, mb_size, mb_size_global, clip_eps, entropy_beta,
                         grads, loss_out, nan_flag, workspace, workspace_bytes, kernel_path, stream, false);
}

// The whole optimiser loop of _update_model (train_ppo2.0.py:42-87) behind one call: `epochs` passes over the M
// transitions in minibatches of mb_size, each step = zero the gradient (inside the gradient launch), plume_ppo_grad, clip + Adam (fused with the
// all-reduce over peer memory when `comm` is given).  The launches are the same as when the host drives the steps one
// by one; the loop only lives on this side of the ABI, because at the reference's BATCH_SIZE = 256 an iteration is
// 20 480 steps of ~15 us and a Python call per launch costs more than the kernels.
extern "C" int plume_allreduce_clip_adam(void* comm, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                         int32_t n, float max_norm, float lr, float beta1, float beta2, float eps,
                                         int32_t step, float* grad_norm_out, void* stream);
extern "C" int plume_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t n,
                               float max_norm, float lr, float beta1, float beta2, float eps, int32_t step,
                               float* grad_norm_out, void* stream);

extern "C" int plume_ppo_update(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                const plume_ppo_batch* batch, const int64_t* perms, uint64_t perm_seed, int32_t epochs,
                                int64_t mb_size, int32_t world, float clip_eps, float entropy_beta, float max_norm,
                                float lr, float beta1, float beta2, float eps, int32_t first_step, void* comm,
                                double* losses, float* grad_norm_out, int32_t* nan_flag, void* workspace,
                                int64_t workspace_bytes, int32_t kernel_path, void* stream) {
    PLUME_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && batch && losses && nan_flag && workspace, "null pointer");
    PLUME_CHECK_ARG(epochs >= 1 && mb_size >= 1 && world >= 1 && first_step >= 1, "bad epochs / minibatch / world / step");
    const int64_t M = batch->total;
    if (M <= 0) return 0;
    int32_t step = first_step;
    int64_t row = 0;
    for (int32_t epoch = 0; epoch < epochs; ++epoch) {
        const int64_t* perm = perms ? perms + (int64_t)epoch * M : nullptr;
        for (int64_t start = 0; start < M; start += mb_size, ++step, ++row) {
            const int64_t size = (M - start) < mb_size ? (M - start) : mb_size;
            int rc = ppo_grad_impl(params, batch, perm, perm_seed, epoch, start, size, size * world, clip_eps,
                                   entropy_beta, grads, losses + 4 * row, nan_flag, workspace, workspace_bytes,
                                   kernel_path, stream, true);
            if (rc) return rc;
            rc = comm ? plume_allreduce_clip_adam(comm, params, grads, exp_avg, exp_avg_sq, PLUME_MLP_PARAMS, max_norm,
                                                  lr, beta1, beta2, eps, step, grad_norm_out, stream)
                      : plume_clip_adam(params, grads, exp_avg, exp_avg_sq, PLUME_MLP_PARAMS, max_norm, lr, beta1, beta2,
                                        eps, step, grad_norm_out, stream);
            if (rc) return rc;
        }
    }
    return 0;
}
