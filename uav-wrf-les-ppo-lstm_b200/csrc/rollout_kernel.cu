// rollout_kernel.cu -- the fused persistent rollout: for a tile of 32 environments one CTA runs
// T lockstep iterations of
//     policy forward + Categorical sample   (train_ppo2.0.py:158-162,185; model.py:38-46)
//     MethaneEnv.step                        (environment.py:89-178)
//     LSTM stop head over the sliding window (evaluate_with_lstm.py:67-80), trend features
//     auto-reset of finished episodes        (environment.py:42-50)
// without returning to the host: MLP (fp16 hi/lo split, 141 KB) and LSTM (17 KB) weights stay in shared memory,
// the per-env state stays in the registers of the tile's first warp, and only the [T][N]
// rollout buffers are written to HBM.  Procedural field mode only (cells are evaluated from
// the Philox stream, so a reset costs O(1)).
//
// Algorithmic bytes per env-step (DESIGN.md "rollout"): obs 24 + action 4 + reward 4 + value 4 +
// logp 4 + done 4 + reached 1 = 45 B written (+ stop 9, info 20, episode 4 when requested); the visit tables are read
// into shared memory once per segment and written back at its end (2 x 208 B per env and segment); nothing else
// touches HBM.
#include "lstm_tile.cuh"
#include "mlp_tc_tile.cuh"

namespace plume {

struct RolloutArgs {
    Cfg c;
    plume_env_state st;
    const float* mlp;
    plume_lstm_params lstm;
    plume_rollout_buffers buf;
    int horizon;
    uint32_t flags;
    int32_t* nan_flag;
};

template <int H>
struct RolloutSmem {
    static constexpr int lstm = PolicySmem::total;
    static constexpr int misc = lstm + (H > 0 ? LstmSmem<(H > 0 ? H : 32)>::total : kLstmMaxSteps * 32);
    static constexpr int vis = misc + 96;       // [32] stop prob, [32] peak, [32] scratch; then the tile's visit tables
    static constexpr int total = vis + kTileM * PLUME_VISIT_STRIDE / 2;     // [32][104] uint16 (16-byte aligned rows)
};

static_assert(RolloutSmem<0>::vis % 4 == 0 && RolloutSmem<32>::vis % 4 == 0, "visit-table rows are copied as uint4");
static_assert(RolloutSmem<32>::total * 4 <= 227 * 1024, "rollout kernel: shared memory plan exceeds 227 KB");

// kSpec (as in the K2 step kernel): 0 = plume model / reward mode / division mode read at run time,
// 1 = reference code model with exact float64 reward, 2 = code model with PLUME_FLAG_FAST_REWARD; 1 and 2 use the
// constant divisions make_cfg() validated and contain only the per-env code of their mode.
template <int H, int kSpec>
__global__ void __launch_bounds__(kMlpThreads, 1) rollout_kernel(RolloutArgs a) {
    extern __shared__ __align__(16) float sm[];
    constexpr int HH = H > 0 ? H : 32;
    float* lsm = sm + RolloutSmem<H>::lstm;
    // the window doubles as the LSTM input xs[t][s]; without a head it is the only part allocated
    float* win = H > 0 ? lsm + LstmSmem<HH>::xs : lsm;
    float* s_stop = sm + RolloutSmem<H>::misc;
    float* s_peak = s_stop + 32;

    Cfg c = a.c;
    if (kSpec != 0) {               // constant-propagated through the inlined per-env code
        c.plume_model = PLUME_MODEL_ISOTROPIC;
        c.fastdiv = 1;
    }
    const int tid = threadIdx.x;
    const int N = a.st.n_envs;
    const int W = a.lstm.window;
    policy_load_weights(sm, a.mlp);
    if (H > 0) {
        const LstmWeights lw{a.lstm.w_ih, a.lstm.w_hh, a.lstm.b_ih, a.lstm.b_hh,
                             a.lstm.w_peak, a.lstm.b_peak, a.lstm.w_stop, a.lstm.b_stop};
        lstm_load_weights<HH>(lsm, lw);
    }
    const ProceduralField field{a.st.sin_tab, a.st.cos_tab};
    const double cur_radius = a.st.curriculum[0], cur_bonus = a.st.curriculum[1];
    const bool greedy = (a.flags & PLUME_FLAG_GREEDY) != 0;
    const bool stop_terminates = (a.flags & PLUME_FLAG_STOP_TERMINATES) != 0;
    // deferred stop head: only record the window inputs; plume_stop_head_segment does the rest
    const bool defer = (a.flags & PLUME_FLAG_DEFER_STOP_HEAD) != 0;
    const bool fast = kSpec == 0 ? (a.flags & PLUME_FLAG_FAST_REWARD) != 0 : kSpec == 2;   // float32 reward terms

    const int tiles = (N + kTileM - 1) / kTileM;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int env = tile * kTileM + tid;          // meaningful for tid < 32
        const bool owner = tid < kTileM && env < N;
        const uint32_t gid = (uint32_t)(a.st.env_id_base + env);
        EnvRegs e{};
        uint16_t* vis = nullptr;
        double cell_conc = 0.0, cell_tke = 0.0;      // field at the float32 cell of the current position
        int fill = 0;
        __syncthreads();
        if (tid < kTileM) {
            float o[6] = {0, 0, 0, 0, 0, 0};
            if (owner) {
                e = load_env(a.st, env);
                // the visit counters live in shared memory for the whole segment: their read-modify-write sits on the
                // serial env-step chain of every lockstep iteration (an L2 round trip per step when left in global memory)
                vis = reinterpret_cast<uint16_t*>(sm + RolloutSmem<H>::vis) + tid * PLUME_VISIT_STRIDE;
                {
                    const uint4* src = reinterpret_cast<const uint4*>(a.st.visited + (size_t)env * PLUME_VISIT_STRIDE);
                    uint4* dst = reinterpret_cast<uint4*>(vis);
#pragma unroll
                    for (int q = 0; q < PLUME_VISIT_STRIDE / 8; ++q) dst[q] = src[q];
                }
                int x, y;
                cell32_of(c, e, x, y);
                field.eval(c, env, gid, e.episode, e.sx, e.sy, x, y, cell_conc, cell_tke);
                make_obs(c, e, cell_conc, cell_tke,
                         vis[(x / c.cell_size) * PLUME_MAX_GRID_DIVISIONS + (y / c.cell_size)], o, gid);
                fill = a.buf.window_fill ? a.buf.window_fill[env] : 0;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) sm[PolicySmem::x + tid * 8 + k] = o[k];
            sm[PolicySmem::x + tid * 8 + 6] = 0.0f;
            sm[PolicySmem::x + tid * 8 + 7] = 0.0f;
            if (!defer)
                for (int k = 0; k < W; ++k)
                    win[k * 32 + tid] = (owner && a.buf.conc_window) ? a.buf.conc_window[(size_t)env * W + k] : 0.0f;
        }

        for (int t = 0; t < a.horizon; ++t) {
            const size_t row = (size_t)t * N;
            // the observation the policy acts on, coalesced [32][6] -> buf.obs[t][tile*32 ..][6]
            __syncthreads();
            if (tid < 6 * kTileM) {
                const int s = tid / 6, k = tid - s * 6;
                if (tile * kTileM + s < N) a.buf.obs[(row + (size_t)tile * kTileM) * 6 + tid] = sm[PolicySmem::x + s * 8 + k];
            }
            policy_forward_tile(sm);
            // ---- part A: sample, step, push the window ------------------------------------------------
            StepResult r{};
            int action = 0;
            float logp = 0.0f, value = 0.0f;
            uint32_t ep_of_transition = e.episode;
            if (owner) {
                const float* o = sm + PolicySmem::out + tid * 8;
                bool bad = false;
#pragma unroll
                for (int k = 0; k < 5; ++k) bad |= isnan(o[k]);
                if (bad) atomicExch(a.nan_flag, 1);
                float p[5];
                softmax5(o, p);
                value = o[5];
                const int forced = a.buf.forced_actions ? a.buf.forced_actions[row + env] : -1;
                const float u = (forced < 0 && !greedy) ? action_uniform(c, gid, e) : 0.0f;
                action = categorical_pick(p, u, greedy, forced, logp);
                double z0, z1;
                if (a.buf.step_noise) {
                    const double2 z = reinterpret_cast<const double2*>(a.buf.step_noise)[row + env];
                    z0 = z.x;
                    z1 = z.y;
                } else {
                    step_noise(c, gid, e, z0, z1);
                }
                if (a.buf.noise_out) reinterpret_cast<double2*>(a.buf.noise_out)[row + env] = make_double2(z0, z1);
                if (fast) env_step_fast(c, field, env, gid, e, vis, action, z0, z1, true, (float)cell_conc, cell_tke, r);
                else env_step(c, field, env, gid, e, vis, action, z0, z1, true, cell_conc, cell_tke, r);
                cell_conc = r.cell_conc;
                cell_tke = r.cell_tke;
                // sliding window of obs[2] (= conc_field[int(x),int(y)]/100 as float32, evaluate_with_lstm.py:67-74)
                fill = fill < W ? fill + 1 : W;
                if (a.buf.conc_sample) a.buf.conc_sample[row + env] = r.obs[2];
                if (a.buf.fill_t) a.buf.fill_t[row + env] = (uint8_t)fill;
                if (a.buf.src_dist) a.buf.src_dist[row + env] = r.distance;
                if (!defer) {
                    for (int k = 0; k + 1 < W; ++k) win[k * 32 + tid] = win[(k + 1) * 32 + tid];
                    win[(W - 1) * 32 + tid] = r.obs[2];
                }
            }
            // ---- LSTM stop head over the window, all threads ------------------------------------------------
            // (skipped while no env of the tile has a full window: the first W-1 steps after a cold start)
            if (H > 0 && __syncthreads_or(owner && fill >= W)) {
                lstm_window_tile<HH>(lsm, W);
                const float hv = lstm_heads<HH>(lsm, W);
                if (tid < 32) s_peak[tid] = hv;
                else if (tid < 64) s_stop[tid - 32] = hv;
                __syncthreads();
            }
            // ---- part B: stop decision, outputs, auto-reset --------------------------------------------------
            if (owner) {
                float stop_p = 0.0f, peak = 0.0f;
                bool stop = false;
                if (H > 0 && fill >= W) {                                    // evaluate_with_lstm.py:73
                    stop_p = s_stop[tid];
                    peak = s_peak[tid];
                    stop = stop_p > a.lstm.threshold;                        // :77
                }
                const bool done = r.done || (stop_terminates && stop);
                const size_t i = row + env;
                a.buf.actions[i] = action;
                a.buf.rewards[i] = (float)r.reward;
                a.buf.values[i] = value;
                a.buf.log_probs[i] = logp;
                a.buf.dones[i] = done ? 1.0f : 0.0f;
                a.buf.reached[i] = r.reached ? 1 : 0;
                if (a.buf.flag_code) a.buf.flag_code[i] = (uint8_t)((done ? 1 : 0) | (r.reached ? 2 : 0));
                if (!defer) {
                    if (a.buf.stop_prob) a.buf.stop_prob[i] = stop_p;
                    if (a.buf.stop_flag) a.buf.stop_flag[i] = stop ? 1 : 0;
                    if (a.buf.peak_pred) a.buf.peak_pred[i] = peak;
                }
                if (a.buf.episode_idx) a.buf.episode_idx[i] = (int32_t)ep_of_transition;
                if (a.buf.pos_out) reinterpret_cast<float2*>(a.buf.pos_out)[i] = make_float2(e.px, e.py);
                if (a.buf.src_out && done) reinterpret_cast<float2*>(a.buf.src_out)[i] = make_float2((float)e.sx, (float)e.sy);
                if (a.buf.conc_out) a.buf.conc_out[i] = (float)cell_conc;       // train_ppo2.0.py:167-173
                if (a.buf.info) {
                    float* inf = a.buf.info + (size_t)t * 5 * N + env;
                    inf[0] = r.conc_reward;
                    inf[(size_t)N] = r.explore_reward;
                    inf[2 * (size_t)N] = (float)r.move_penalty;
                    inf[3 * (size_t)N] = r.tke_penalty;
                    inf[4 * (size_t)N] = (float)r.boundary_penalty;
                }
                if (a.buf.trend && !defer) {
                    float tr[4] = {0, 0, 0, 0};
                    if (fill >= W && W >= 4) {
                        trend_from_last4(100.0 * (double)win[(W - 4) * 32 + tid], 100.0 * (double)win[(W - 3) * 32 + tid],
                                         100.0 * (double)win[(W - 2) * 32 + tid], 100.0 * (double)win[(W - 1) * 32 + tid],
                                         r.distance, c.conc_peak, tr);
                    }
                    *reinterpret_cast<float4*>(a.buf.trend + i * 4) = make_float4(tr[0], tr[1], tr[2], tr[3]);
                }
                if (done) {
                    env_reset(c, gid, e, vis, nullptr, cur_radius, cur_bonus);
                    fill = 0;
                    field.eval(c, env, gid, e.episode, e.sx, e.sy, 0, 0, cell_conc, cell_tke);
                    make_obs(c, e, cell_conc, cell_tke, 0, r.obs, gid);
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) sm[PolicySmem::x + tid * 8 + k] = r.obs[k];
            }
        }
        // ---- persist the tile's state ----------------------------------------------------------------------
        if (owner) {
            store_env(a.st, env, e);
            {
                uint4* dst = reinterpret_cast<uint4*>(a.st.visited + (size_t)env * PLUME_VISIT_STRIDE);
                const uint4* src = reinterpret_cast<const uint4*>(vis);
#pragma unroll
                for (int q = 0; q < PLUME_VISIT_STRIDE / 8; ++q) dst[q] = src[q];
            }
            if (a.st.cell_tke && a.st.cell_conc && a.st.cell_key) {   // hand the carried cell to a following plume_env_step
                int x, y;
                cell32_of(c, e, x, y);
                a.st.cell_tke[env] = cell_tke;
                a.st.cell_conc[env] = cell_conc;
                a.st.cell_key[env] = cell_key_of(c, x, y, e.episode);
            }
            if (a.buf.window_fill) a.buf.window_fill[env] = fill;
            if (a.buf.conc_window && !defer)
                for (int k = 0; k < W; ++k) a.buf.conc_window[(size_t)env * W + k] = win[k * 32 + tid];
            if (a.buf.last_obs) {
#pragma unroll
                for (int k = 0; k < 6; ++k) a.buf.last_obs[(size_t)env * 6 + k] = sm[PolicySmem::x + tid * 8 + k];
            }
        }
    }
}

// ---- the pipelined form (no in-loop stop head: training rollouts, deferred head) -----------------------------------
// ncu of the kernel above (profiles/r1u_rollout_ncu_summary.txt): 33 % of the samples sit at the top-of-loop barrier --
// seven warps wait while one warp steps the tile's 32 envs (a ~1200-instruction float64 dependency chain), then that
// warp waits for the policy forward.  Here the tile is two HALVES of 16 envs with their own env warp each, next to the
// eight MLP warps: while the MLP warps run the policy forward of half A (mlp_tc_forward_half), the env warp of half B
// steps its envs, and vice versa.  The hand-over is four named barriers (bar.arrive by the producer, bar.sync by the
// consumer; 256 + 32 threads each): OBS(h) "x[h] holds the observations of the next step", LOGITS(h) "out[h] holds
// the logits".  The action-independent Philox draws of a step (action uniform, step noise) are taken by the env warp
// BEFORE it waits for the logits, i.e. off the critical path.  Same per-env arithmetic as the kernel above: the
// transitions are bit-identical.
constexpr int kPipeThreads = kMlpThreads + 64;       // 8 MLP warps + 2 env warps
constexpr int kBarObs = 2, kBarLogits = 4;           // named barrier ids: kBarObs + h, kBarLogits + h
constexpr int kPipeBarThreads = kMlpThreads + 32;

__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) {
    __threadfence_block();           // the producer's shared-memory writes are ordered before its arrival
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

template <int kSpec>
__global__ void __launch_bounds__(kPipeThreads, 1) rollout_pipe_kernel(RolloutArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* lsm = sm + RolloutSmem<0>::lstm;          // the [W][32] window ring (only maintained without `defer`)
    float* win = lsm;
    Cfg c = a.c;
    if (kSpec != 0) {               // constant-propagated through the inlined per-env code
        c.plume_model = PLUME_MODEL_ISOTROPIC;
        c.fastdiv = 1;
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.st.n_envs;
    const int W = a.lstm.window;
    if (warp < 8) policy_load_weights(sm, a.mlp);
    const ProceduralField field{a.st.sin_tab, a.st.cos_tab};
    const double cur_radius = a.st.curriculum[0], cur_bonus = a.st.curriculum[1];
    const bool greedy = (a.flags & PLUME_FLAG_GREEDY) != 0;
    const bool defer = (a.flags & PLUME_FLAG_DEFER_STOP_HEAD) != 0;
    const bool fast = kSpec == 0 ? (a.flags & PLUME_FLAG_FAST_REWARD) != 0 : kSpec == 2;   // float32 reward terms
    // evaluator stop tests (N1): 0 none, 1 fixed (PPOV1.1/evaluate_model.py:25-37), 2 threshold controller
    // (PPOV2.0/evaluate_with_lstm.py:28-37); the last 10 samples of every env live in the window region
    const int stop_mode = (a.flags & PLUME_FLAG_STOP_FIXED) ? 1 : ((a.flags & PLUME_FLAG_STOP_THRESHOLD) ? 2 : 0);
    const int step_guard = (int)(a.flags >> 16);
    double* const ring = reinterpret_cast<double*>(lsm);          // [10][32] doubles (or float pairs)
    const int tiles = (N + kTileM - 1) / kTileM;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        __syncthreads();            // weights loaded / the previous tile is finished everywhere
        if (warp < 8) {
            // ---- MLP warps: policy forward of half 0, half 1, half 0, ... as their observations arrive -------------
            for (int t = 0; t < a.horizon; ++t) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
#ifdef PLUME_ROLLOUT_TIMELINE
                    const long long tl0 = clock64();
#endif
                    bar_sync_n(kBarObs + h, kPipeBarThreads);           // x[h] of step t is in shared memory
#ifdef PLUME_ROLLOUT_TIMELINE
                    const long long tl1 = clock64();
#endif
                    mlp_tc_forward_half<true>(sm, h);
                    bar_arrive_n(kBarLogits + h, kPipeBarThreads);      // out[h] written by this thread
#ifdef PLUME_ROLLOUT_TIMELINE
                    if (blockIdx.x == 0 && tid == 0 && (t == 100 || t == 101))
                        printf("rollout timeline t=%d half %d: MLP warps waited %lld cycles for the observations, forward %lld\n",
                               t, h, tl1 - tl0, clock64() - tl1);
#endif
                }
            }
            continue;
        }
        // ---- env warps: warp 8 + h owns envs 16 h .. 16 h + 15 of the tile (lanes 0..15) ------------------------------
        const int h = warp - 8;
        const int slot = 16 * h + (lane & 15);       // row of the tile in x / out / vis / win
        const int env = tile * kTileM + slot;
        const bool owner = lane < 16 && env < N;
        const uint32_t gid = (uint32_t)(a.st.env_id_base + env);
        EnvRegs e{};
        uint16_t* vis = reinterpret_cast<uint16_t*>(sm + RolloutSmem<0>::vis) + slot * PLUME_VISIT_STRIDE;
        double cell_conc = 0.0, cell_tke = 0.0;      // field at the float32 cell of the current position
        int fill = 0;
        {
            float o[6] = {0, 0, 0, 0, 0, 0};
            if (owner) {
                e = load_env(a.st, env);
                {
                    const uint4* src = reinterpret_cast<const uint4*>(a.st.visited + (size_t)env * PLUME_VISIT_STRIDE);
                    uint4* dst = reinterpret_cast<uint4*>(vis);
#pragma unroll
                    for (int q = 0; q < PLUME_VISIT_STRIDE / 8; ++q) dst[q] = src[q];
                }
                int x, y;
                cell32_of(c, e, x, y);
                field.eval(c, env, gid, e.episode, e.sx, e.sy, x, y, cell_conc, cell_tke);
                make_obs(c, e, cell_conc, cell_tke,
                         vis[(x / c.cell_size) * PLUME_MAX_GRID_DIVISIONS + (y / c.cell_size)], o, gid);
                fill = a.buf.window_fill ? a.buf.window_fill[env] : 0;
            }
            if (lane < 16) {
#pragma unroll
                for (int k = 0; k < 6; ++k) sm[PolicySmem::x + slot * 8 + k] = o[k];
                sm[PolicySmem::x + slot * 8 + 6] = 0.0f;
                sm[PolicySmem::x + slot * 8 + 7] = 0.0f;
                if (stop_mode) {
                    for (int k = 0; k < 10; ++k)
                        ring[k * 32 + slot] = (owner && a.buf.eval_ring) ? a.buf.eval_ring[(size_t)env * 10 + k] : 0.0;
                } else if (!defer) {
                    for (int k = 0; k < W; ++k)
                        win[k * 32 + slot] = (owner && a.buf.conc_window) ? a.buf.conc_window[(size_t)env * W + k] : 0.0f;
                }
            }
        }
        for (int t = 0; t < a.horizon; ++t) {
            const size_t row = (size_t)t * N;
            // the observation the policy acts on: x[h] -> buf.obs[t][tile*32 + 16h ..][6], 96 consecutive floats
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int j = lane + 32 * q, s16 = j / 6, k = j - 6 * s16;
                if (tile * kTileM + 16 * h + s16 < N)
                    a.buf.obs[(row + (size_t)tile * kTileM + 16 * h) * 6 + j] = sm[PolicySmem::x + (16 * h + s16) * 8 + k];
            }
            bar_arrive_n(kBarObs + h, kPipeBarThreads);
#ifdef PLUME_ROLLOUT_TIMELINE
            const long long te0 = clock64();
#endif
            // the draws of this step do not depend on the action: taken while the MLP warps work
            const int forced = (owner && a.buf.forced_actions) ? a.buf.forced_actions[row + env] : -1;
            float u = 0.0f;
            double z0 = 0.0, z1 = 0.0;
            if (owner) {
                if (forced < 0 && !greedy) u = action_uniform(c, gid, e);
                if (a.buf.step_noise) {
                    const double2 z = reinterpret_cast<const double2*>(a.buf.step_noise)[row + env];
                    z0 = z.x;
                    z1 = z.y;
                } else {
                    step_noise(c, gid, e, z0, z1);
                }
            }
#ifdef PLUME_ROLLOUT_TIMELINE
            const long long te1 = clock64();
#endif
            bar_sync_n(kBarLogits + h, kPipeBarThreads);                // out[h] of step t is complete
#ifdef PLUME_ROLLOUT_TIMELINE
            const long long te2 = clock64();
#endif
            if (owner) {
                const float* o = sm + PolicySmem::out + slot * 8;
                bool bad = false;
#pragma unroll
                for (int k = 0; k < 5; ++k) bad |= isnan(o[k]);
                if (bad) atomicExch(a.nan_flag, 1);
                float p[5];
                softmax5(o, p);
                const float value = o[5];
                float logp;
                const int action = categorical_pick(p, u, greedy, forced, logp);
                const uint32_t ep_of_transition = e.episode;
                if (a.buf.noise_out) reinterpret_cast<double2*>(a.buf.noise_out)[row + env] = make_double2(z0, z1);
                StepResult r{};
                if (fast) env_step_fast(c, field, env, gid, e, vis, action, z0, z1, true, (float)cell_conc, cell_tke, r);
                else env_step(c, field, env, gid, e, vis, action, z0, z1, true, cell_conc, cell_tke, r);
                cell_conc = r.cell_conc;
                cell_tke = r.cell_tke;
                // sliding window of obs[2] (= conc_field[int(x),int(y)]/100 as float32, evaluate_with_lstm.py:67-74)
                fill = fill < W ? fill + 1 : W;
                const size_t i = row + env;
                if (a.buf.conc_sample) a.buf.conc_sample[i] = r.obs[2];
                if (a.buf.fill_t) a.buf.fill_t[i] = (uint8_t)fill;
                if (a.buf.src_dist) a.buf.src_dist[i] = r.distance;
                bool stop = false;
                if (stop_mode) {
                    for (int k = 0; k + 1 < 10; ++k) ring[k * 32 + slot] = ring[(k + 1) * 32 + slot];
                    if (stop_mode == 1) {
                        // np.std(last 10 positions, axis=0).mean() < 2 (float32, like the reference's float32 arrays) and
                        // info['concentration_reward'] * CONC_PEAK * CONC_PEAK (sic) > 0.8 * CONC_PEAK
                        ring[9 * 32 + slot] = __hiloint2double(__float_as_int(e.py), __float_as_int(e.px));
                        if (e.step >= 10) {
                            float mx = 0.0f, my = 0.0f;
                            for (int k = 0; k < 10; ++k) {
                                const double v = ring[k * 32 + slot];
                                mx += __int_as_float(__double2loint(v));
                                my += __int_as_float(__double2hiint(v));
                            }
                            mx *= 0.1f;
                            my *= 0.1f;
                            float vx = 0.0f, vy = 0.0f;
                            for (int k = 0; k < 10; ++k) {
                                const double v = ring[k * 32 + slot];
                                const float dx = __int_as_float(__double2loint(v)) - mx, dy = __int_as_float(__double2hiint(v)) - my;
                                vx = fmaf(dx, dx, vx);
                                vy = fmaf(dy, dy, vy);
                            }
                            const float pos_std = 0.5f * (sqrtf(vx * 0.1f) + sqrtf(vy * 0.1f));
                            const float cur = r.conc_reward * (float)c.conc_peak * (float)c.conc_peak;
                            stop = pos_std < 2.0f && cur > 0.8f * (float)c.conc_peak;
                        }
                    } else {
                        ring[9 * 32 + slot] = cell_conc;
                        // the segment's last step is tested by the host against the threshold refreshed from this very
                        // sample (evaluate_with_lstm.py:89-92: update_threshold precedes should_stop)
                        if (e.step >= 20 && t + 1 < a.horizon && a.buf.stop_threshold) {
                            const double thr = a.buf.stop_threshold[env];
                            if (thr == thr) {
                                const int n = e.step < 10 ? e.step : 10;
                                double sum = 0.0;
                                for (int k = 10 - n; k < 10; ++k) sum += ring[k * 32 + slot];
                                stop = cell_conc >= thr || sum / (double)n >= thr;
                            }
                        }
                    }
                } else if (!defer) {
                    for (int k = 0; k + 1 < W; ++k) win[k * 32 + slot] = win[(k + 1) * 32 + slot];
                    win[(W - 1) * 32 + slot] = r.obs[2];
                }
                const bool done = r.done || stop || (step_guard > 0 && e.step >= step_guard);
                a.buf.actions[i] = action;
                a.buf.rewards[i] = (float)r.reward;
                a.buf.values[i] = value;
                a.buf.log_probs[i] = logp;
                a.buf.dones[i] = done ? 1.0f : 0.0f;
                a.buf.reached[i] = r.reached ? 1 : 0;
                if (a.buf.flag_code) a.buf.flag_code[i] = (uint8_t)((done ? 1 : 0) | (r.reached ? 2 : 0));
                if (!defer) {                       // no stop head in this kernel: its outputs are zero
                    if (a.buf.stop_prob) a.buf.stop_prob[i] = 0.0f;
                    if (a.buf.stop_flag) a.buf.stop_flag[i] = stop ? 1 : 0;
                    if (a.buf.peak_pred) a.buf.peak_pred[i] = 0.0f;
                }
                if (a.buf.episode_idx) a.buf.episode_idx[i] = (int32_t)ep_of_transition;
                if (a.buf.pos_out) reinterpret_cast<float2*>(a.buf.pos_out)[i] = make_float2(e.px, e.py);
                if (a.buf.src_out && done) reinterpret_cast<float2*>(a.buf.src_out)[i] = make_float2((float)e.sx, (float)e.sy);
                if (a.buf.conc_out) a.buf.conc_out[i] = (float)cell_conc;       // train_ppo2.0.py:167-173
                if (a.buf.info) {
                    float* inf = a.buf.info + (size_t)t * 5 * N + env;
                    inf[0] = r.conc_reward;
                    inf[(size_t)N] = r.explore_reward;
                    inf[2 * (size_t)N] = (float)r.move_penalty;
                    inf[3 * (size_t)N] = r.tke_penalty;
                    inf[4 * (size_t)N] = (float)r.boundary_penalty;
                }
                if (a.buf.trend && !defer && !stop_mode) {
                    float tr[4] = {0, 0, 0, 0};
                    if (fill >= W && W >= 4) {
                        trend_from_last4(100.0 * (double)win[(W - 4) * 32 + slot], 100.0 * (double)win[(W - 3) * 32 + slot],
                                         100.0 * (double)win[(W - 2) * 32 + slot], 100.0 * (double)win[(W - 1) * 32 + slot],
                                         r.distance, c.conc_peak, tr);
                    }
                    *reinterpret_cast<float4*>(a.buf.trend + i * 4) = make_float4(tr[0], tr[1], tr[2], tr[3]);
                }
                if (done) {
                    env_reset(c, gid, e, vis, nullptr, cur_radius, cur_bonus);
                    fill = 0;
                    field.eval(c, env, gid, e.episode, e.sx, e.sy, 0, 0, cell_conc, cell_tke);
                    make_obs(c, e, cell_conc, cell_tke, 0, r.obs, gid);
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) sm[PolicySmem::x + slot * 8 + k] = r.obs[k];
            }
#ifdef PLUME_ROLLOUT_TIMELINE
            if (blockIdx.x == 0 && lane == 0 && (t == 100 || t == 101))
                printf("rollout timeline t=%d env warp %d: draws %lld cycles, waited %lld for the logits, step + outputs %lld\n", t,
                       h, te1 - te0, te2 - te1, clock64() - te2);
#endif
        }
        // ---- persist the half tile's state --------------------------------------------------------------------
        if (owner) {
            store_env(a.st, env, e);
            {
                uint4* dst = reinterpret_cast<uint4*>(a.st.visited + (size_t)env * PLUME_VISIT_STRIDE);
                const uint4* src = reinterpret_cast<const uint4*>(vis);
#pragma unroll
                for (int q = 0; q < PLUME_VISIT_STRIDE / 8; ++q) dst[q] = src[q];
            }
            if (a.st.cell_tke && a.st.cell_conc && a.st.cell_key) {   // hand the carried cell to a following plume_env_step
                int x, y;
                cell32_of(c, e, x, y);
                a.st.cell_tke[env] = cell_tke;
                a.st.cell_conc[env] = cell_conc;
                a.st.cell_key[env] = cell_key_of(c, x, y, e.episode);
            }
            if (a.buf.window_fill) a.buf.window_fill[env] = fill;
            if (stop_mode) {
                if (a.buf.eval_ring)
                    for (int k = 0; k < 10; ++k) a.buf.eval_ring[(size_t)env * 10 + k] = ring[k * 32 + slot];
            } else if (a.buf.conc_window && !defer) {
                for (int k = 0; k < W; ++k) a.buf.conc_window[(size_t)env * W + k] = win[k * 32 + slot];
            }
            if (a.buf.last_obs) {
#pragma unroll
                for (int k = 0; k < 6; ++k) a.buf.last_obs[(size_t)env * 6 + k] = sm[PolicySmem::x + slot * 8 + k];
            }
        }
    }
}

// ---- N1: evaluator bookkeeping -----------------------------------------------------------------------------------------
__global__ void eval_collect_kernel(plume_rollout_buffers buf, int T, int N, int base_step, uint8_t* finished, int32_t* steps,
                                    uint8_t* early, int32_t* stop_step, double* deviation, const double* pending) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || finished[n]) return;
    for (int t = 0; t < T; ++t) {
        const size_t i = (size_t)t * N + n;
        const bool stop = buf.stop_flag && buf.stop_flag[i] != 0;
        if (stop || buf.dones[i] != 0.0f) {
            steps[n] = base_step + t + 1;
            early[n] = stop ? 1 : 0;
            stop_step[n] = stop ? base_step + t + 1 : 0;
            deviation[n] = buf.src_dist[i];
            finished[n] = 1;
            return;
        }
    }
    if (pending && buf.eval_ring) {       // ThresholdController.should_stop for the segment's last step
        const int s = base_step + T;
        const double thr = pending[n];
        if (s >= 20 && thr == thr) {
            const double* ring = buf.eval_ring + (size_t)n * 10;
            const int cnt = s < 10 ? s : 10;
            double sum = 0.0;
            for (int k = 10 - cnt; k < 10; ++k) sum += ring[k];
            if (ring[9] >= thr || sum / (double)cnt >= thr) {
                steps[n] = s;
                early[n] = 1;
                stop_step[n] = s;
                deviation[n] = buf.src_dist[(size_t)(T - 1) * N + n];
                finished[n] = 1;
            }
        }
    }
}

template <int kSpec>
static int launch_rollout_pipe_spec(const RolloutArgs& a, cudaStream_t s) {
    static bool configured = false;
    const int smem = RolloutSmem<0>::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(rollout_pipe_kernel<kSpec>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("pipelined rollout kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const int tiles = (a.st.n_envs + kTileM - 1) / kTileM;
    int grid = sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    rollout_pipe_kernel<kSpec><<<grid, kPipeThreads, smem, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return fail("pipelined rollout kernel launch failed");
    return 0;
}

static int launch_rollout_pipe(const RolloutArgs& a, cudaStream_t s) {
    if (a.c.plume_model == PLUME_MODEL_ISOTROPIC && a.c.fastdiv)
        return (a.flags & PLUME_FLAG_FAST_REWARD) ? launch_rollout_pipe_spec<2>(a, s) : launch_rollout_pipe_spec<1>(a, s);
    return launch_rollout_pipe_spec<0>(a, s);
}

template <int H, int kSpec>
static int launch_rollout_spec(const RolloutArgs& a, cudaStream_t s) {
    static bool configured = false;
    const int smem = RolloutSmem<H>::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(rollout_kernel<H, kSpec>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
            return fail("rollout kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const int tiles = (a.st.n_envs + kTileM - 1) / kTileM;
    int grid = sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    rollout_kernel<H, kSpec><<<grid, kMlpThreads, smem, s>>>(a);
    if (cudaGetLastError() != cudaSuccess) return fail("rollout kernel launch failed");
    return 0;
}

template <int H>
static int launch_rollout(const RolloutArgs& a, cudaStream_t s) {
    if (a.c.plume_model == PLUME_MODEL_ISOTROPIC && a.c.fastdiv)
        return (a.flags & PLUME_FLAG_FAST_REWARD) ? launch_rollout_spec<H, 2>(a, s) : launch_rollout_spec<H, 1>(a, s);
    return launch_rollout_spec<H, 0>(a, s);
}

}  // namespace plume

using namespace plume;

extern "C" int plume_rollout(const plume_env_config* cfg, const plume_env_state* st, const float* mlp_params,
                             const plume_lstm_params* lstm, const plume_rollout_buffers* buf, int32_t horizon,
                             uint32_t flags, int32_t* nan_flag, void* stream) {
    PLUME_CHECK_ARG(cfg && st && mlp_params && buf && nan_flag, "null pointer");
    PLUME_CHECK_ARG(cfg->field_mode == PLUME_FIELD_PROCEDURAL,
                    "the fused rollout needs the procedural field mode (no CPU or materialised fallback)");
    PLUME_CHECK_ARG(cfg->grid_size > 0 && cfg->grid_size % 2 == 0, "grid_size must be even");
    PLUME_CHECK_ARG(cfg->grid_divisions > 0 && cfg->grid_divisions <= PLUME_MAX_GRID_DIVISIONS, "grid_divisions");
    PLUME_CHECK_ARG(cfg->max_steps > 0 && cfg->max_steps <= 65535, "max_steps must fit uint16 visit counters");
    PLUME_CHECK_ARG(st->sin_tab && st->cos_tab && st->curriculum, "sin_tab/cos_tab/curriculum missing");
    PLUME_CHECK_ARG(buf->obs && buf->actions && buf->rewards && buf->values && buf->log_probs && buf->dones &&
                        buf->reached, "null rollout buffer");
    if (flags & PLUME_FLAG_DEFER_STOP_HEAD) {
        PLUME_CHECK_ARG(!(flags & PLUME_FLAG_STOP_TERMINATES),
                        "a stop decision that ends the episode cannot be deferred out of the rollout loop");
        PLUME_CHECK_ARG(buf->conc_sample && buf->fill_t && buf->window_fill,
                        "the deferred stop head needs conc_sample, fill_t and window_fill");
    }
    if (flags & (PLUME_FLAG_STOP_FIXED | PLUME_FLAG_STOP_THRESHOLD)) {
        PLUME_CHECK_ARG(!(lstm && lstm->hidden > 0 && !(flags & PLUME_FLAG_DEFER_STOP_HEAD)),
                        "the evaluator stop tests run in the kernel without an in-loop LSTM head");
        PLUME_CHECK_ARG(!(flags & PLUME_FLAG_DEFER_STOP_HEAD), "evaluator stop tests and a deferred stop head exclude each other");
        PLUME_CHECK_ARG(buf->eval_ring, "evaluator stop tests need eval_ring");
        PLUME_CHECK_ARG(!(flags & PLUME_FLAG_STOP_THRESHOLD) || buf->stop_threshold, "STOP_THRESHOLD needs stop_threshold");
    }
    if (horizon <= 0 || st->n_envs <= 0) return 0;
    RolloutArgs a;
    a.c = make_cfg(*cfg, *st);
    a.st = *st;
    a.mlp = mlp_params;
    a.buf = *buf;
    a.horizon = horizon;
    a.flags = flags;
    a.nan_flag = nan_flag;
    if (lstm && lstm->hidden > 0 && !(flags & PLUME_FLAG_DEFER_STOP_HEAD)) {
        a.lstm = *lstm;
        PLUME_CHECK_ARG(lstm->window >= 1 && lstm->window <= kLstmMaxSteps, "stop-head window must be in [1,32]");
        PLUME_CHECK_ARG(lstm->w_ih && lstm->w_hh && lstm->b_ih && lstm->b_hh && lstm->w_peak && lstm->b_peak &&
                            lstm->w_stop && lstm->b_stop, "null LSTM parameter");
        if (lstm->hidden == 32) return launch_rollout<32>(a, as_stream(stream));
        return fail("plume_rollout: fused stop head supports hidden=32 (got %d); use plume_lstm_stop_head "
                    "for other sizes", lstm->hidden);
    }
    a.lstm = plume_lstm_params{};
    a.lstm.window = (lstm && lstm->window > 0 && lstm->window <= kLstmMaxSteps) ? lstm->window : 20;
#ifdef PLUME_ROLLOUT_LOCKSTEP          // A/B build: the single-env-warp kernel also for rollouts without an in-loop head
    return launch_rollout<0>(a, as_stream(stream));
#else
    return launch_rollout_pipe(a, as_stream(stream));
#endif
}

extern "C" int plume_eval_collect(const plume_rollout_buffers* buf, int32_t horizon, int32_t n_envs, int32_t base_step,
                                  uint8_t* finished, int32_t* steps, uint8_t* early, int32_t* stop_step, double* deviation,
                                  const double* pending_threshold, void* stream) {
    PLUME_CHECK_ARG(buf && finished && steps && early && stop_step && deviation, "null pointer");
    PLUME_CHECK_ARG(buf->dones && buf->src_dist, "the evaluator bookkeeping needs dones and src_dist");
    if (horizon <= 0 || n_envs <= 0) return 0;
    eval_collect_kernel<<<(n_envs + 127) / 128, 128, 0, as_stream(stream)>>>(*buf, horizon, n_envs, base_step, finished, steps,
                                                                           early, stop_step, deviation, pending_threshold);
    PLUME_LAUNCH_CHECK();
    return 0;
}
