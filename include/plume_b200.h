/*
 * plume_b200.h -- C ABI of libplume_b200.so: the B200 (sm_100a) data-parallel hot path of
 * su1phurd/UAV-WRF-LES-PPO-LSTM (vectorised plume rollout + PPO update).
 *
 * The reference has no FFI: its boundary is three duck-typed Python surfaces
 * (environment.py reset/step, model.py actor/critic/LSTM forward, train_ppo*.py
 * _update_model).  Every entry point below replaces the arithmetic of one of those
 * reference functions for a whole batch of environments; the citation next to each
 * declaration names it (paths are relative to the reference root, PPOV2.1 unless noted).
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch types.  Unless a parameter is documented as
 *    HOST, every pointer is a DEVICE pointer on the current CUDA device.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All
 *    launches are asynchronous on that stream; nothing synchronises unless stated.
 *  - every function returns 0 on success, non-zero on error; plume_last_error() gives
 *    the message of the last failure on the calling thread.
 *  - there is no CPU fallback: without a CUDA device the calls fail.
 */
#ifndef PLUME_B200_H
#define PLUME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLUME_B200_ABI_VERSION 4

#define PLUME_OBS_DIM 6          /* environment.py:80-87 */
#define PLUME_NUM_ACTIONS 5      /* environment.py:23 */
#define PLUME_INFO_DIM 5         /* environment.py:170-176 */
#define PLUME_VISIT_STRIDE 104   /* 10x10 visit counters (uint16) padded to 13 x 16 B */
#define PLUME_MAX_GRID_DIVISIONS 10

/* field_mode */
#define PLUME_FIELD_PROCEDURAL 0 /* no field in memory: cells are evaluated from the Philox stream */
#define PLUME_FIELD_F32 1        /* conc/tke materialised as float  [N,G,G] */
#define PLUME_FIELD_F64 2        /* conc/tke materialised as double [N,G,G] (reference layout, environment.py:62-63) */

/* plume_model */
#define PLUME_MODEL_ISOTROPIC 0  /* the reference CODE: isotropic Gaussian + noise, five-term reward
                                  * (environment.py:52-63,146-158) -- the parity target */
#define PLUME_MODEL_DISPERSION 1 /* the reference README (README.md:48-53,95-100; no code behind it): Gaussian
                                  * dispersion sigma_y = 0.3 x^0.71 in the wind-rotated frame, observation
                                  * [x, y, CH4, wind_x, step, wind_y], reward R = dCH4 - 0.2 |dtheta|
                                  * (+ the code's arrival bonus).  Oracle = oracle/plume_oracle.py, not pinned
                                  * by the reference. */

/* flags of plume_env_step / plume_rollout */
#define PLUME_FLAG_AUTO_RESET 1u     /* finished envs are reset inside the call (procedural mode) */
#define PLUME_FLAG_GREEDY 2u         /* argmax actions (evaluate_with_lstm.py:65) instead of sampling */
#define PLUME_FLAG_STOP_TERMINATES 4u /* a stop-head decision ends the episode (evaluate_with_lstm.py:77-80) */
#define PLUME_FLAG_DEFER_STOP_HEAD 8u /* plume_rollout only records the stop-head inputs (conc_sample, fill_t,
                                       * src_dist); plume_stop_head_segment evaluates the head and the trend
                                       * features for the whole [T][N] segment afterwards.  Identical results;
                                       * not combinable with PLUME_FLAG_STOP_TERMINATES. */

#define PLUME_FLAG_FAST_REWARD 16u   /* plume_env_step / plume_rollout: flags and indices (position, cells, visit
                                     * counters, reached, done) stay the float64 arithmetic of environment.py, the
                                     * concentration / tke observation entries and the reward terms are float32
                                     * (fp32 rel 1e-5 bar instead of bit-exact float64 rewards) */
/* Evaluator stop tests inside plume_rollout (they end the episode like `done`; the kernel without an in-loop LSTM head
 * only).  STOP_FIXED: PPOV1.1/evaluate_model.py:25-37 -- std of the last 10 positions < 2 px and the (sic, twice-scaled)
 * concentration above 0.8 CONC_PEAK, from step 10 on.  STOP_THRESHOLD: ThresholdController.should_stop
 * (PPOV2.0/evaluate_with_lstm.py:28-37) against the per-env threshold in buf.stop_threshold (NaN = none yet) -- from
 * step 20 on, concentration or mean of the last 10 >= threshold; the test of the segment's LAST step is left to the host,
 * which refreshes the threshold from that step's sample first (:89-92).  Bits 16..31 of flags: step guard (an episode
 * ends when step_count reaches it; evaluate_model.py:52 uses 2000), 0 = none. */
#define PLUME_FLAG_STOP_FIXED 32u
#define PLUME_FLAG_STOP_THRESHOLD 64u

/* Constants of one reference version (PPOV x/config.py; environment.py). HOST struct. */
typedef struct plume_env_config {
    int32_t grid_size;            /* config.py:6 */
    int32_t max_steps;            /* config.py:7 (V1.1: 5000) */
    int32_t grid_divisions;       /* config.py:27 */
    int32_t field_mode;           /* PLUME_FIELD_* */
    double conc_peak;             /* config.py:8,13 */
    double turbulence_intensity;  /* config.py:9 */
    double sigma;                 /* config.py:12; V2.0/V1.1: grid/16 (PPOV2.0/environment.py:54) */
    double clip_hi;               /* environment.py:112; V1.1: grid-1e-6 (PPOV1.1/environment.py:105) */
    double conc_reward_coef;      /* config.py:38 */
    double tke_penalty_factor;    /* config.py:39 */
    double boundary_penalty;      /* config.py:40 */
    double boundary_decay_start;  /* config.py:41 */
    double initial_radius;        /* config.py:31 */
    uint64_t seed;                /* Philox key */
    int32_t plume_model;          /* PLUME_MODEL_* */
    int32_t reserved0;
} plume_env_config;

/* Struct-of-arrays state of n_envs environments (DEVICE pointers, HOST struct).
 * Replaces the attributes of MethaneEnv (environment.py:20-50). */
typedef struct plume_env_state {
    int32_t n_envs;
    int32_t env_id_base;          /* global id of env 0 (rank * n_envs): Philox counters use global ids */
    float* pos_x;                 /* agent_pos after astype(float32), environment.py:113 */
    float* pos_y;
    double* src_x;                /* source_pos, environment.py:44 */
    double* src_y;
    int32_t* step_count;          /* environment.py:47,90 */
    int32_t* episode_idx;         /* number of resets so far (Philox counter, episode index output) */
    uint16_t* visited;            /* [n_envs][PLUME_VISIT_STRIDE], environment.py:38,136 */
    double* radius;               /* current_radius latched at reset, environment.py:32; model.py:189 */
    double* explore_bonus;        /* explore_bonus latched at reset, environment.py:39; model.py:190 */
    void* conc_field;             /* [n_envs][G][G] float/double or NULL (procedural) */
    void* tke_field;              /* idem */
    const double* sin_tab;        /* [G] sin(0.05 x), environment.py:59 */
    const double* cos_tab;        /* [G] cos(0.07 y) */
    const double* curriculum;     /* [2] = {current_radius, explore_bonus} that resets latch (model.py:189-190) */
    int8_t* last_move;            /* last non-zero action (heading) per env, PLUME_MODEL_DISPERSION only; may be NULL */
    /* ABI 2, all optional (NULL = evaluate in the kernel).  Host tables of the two integer-indexed terms:
     * step_frac_tab[s] = (float)(s / MAX_STEPS), s in [0,max_steps] (environment.py:85);
     * visit_denom_tab[v] = (float)(v**0.75 + 1), v in [0,max_steps+1] (environment.py:140). */
    const float* step_frac_tab;
    const float* visit_denom_tab;
    /* procedural mode: tke_field / conc_field at the float32 cell of the current position (the "prev" lookups
     * of the next step, environment.py:93-95,106), tagged with (cell, episode) in cell_key so that a stale
     * entry -- state edited from the host -- is recomputed, never trusted.  Zero-initialise cell_key; the
     * concentration entry also depends on the source position: clear cell_key when src_x/src_y are edited. */
    double* cell_tke;
    double* cell_conc;
    uint32_t* cell_key;
} plume_env_state;

/* ---- library ------------------------------------------------------------------------- */
int plume_abi_version(void);
const char* plume_last_error(void);
/* HOST outputs: SM count, compute capability. */
int plume_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---- P0 reset: MethaneEnv.reset, environment.py:42-50 -------------------------------- */
/* Resets the envs listed in env_list (int32[n_list] device, NULL = all n_envs): draws the
 * source (Philox TAG_SRC) unless u_src (double[n_list][2], the two rand() of :44) is given,
 * zeroes position/step/visit table, bumps episode_idx, latches the curriculum scalars.
 * Materialised fields are NOT regenerated here (call plume_generate_fields next). */
int plume_env_reset(const plume_env_config* cfg, const plume_env_state* st, const int32_t* env_list,
                    int32_t n_list, const double* u_src, void* stream);

/* ---- P1 _generate_plume, environment.py:52-63 ----------------------------------------- */
/* Writes conc/tke fields of the listed envs (field_mode F32 or F64) from the Philox field
 * stream of each env's current episode.  Optionally also dumps the raw draws
 * z_out/u_out (float[n_list][G][G] each, may be NULL) so a CPU oracle can rebuild the
 * identical field. */
int plume_generate_fields(const plume_env_config* cfg, const plume_env_state* st, const int32_t* env_list,
                          int32_t n_list, float* z_out, float* u_out, void* stream);

/* Noise draws of individual cells: z_out/u_out float[n] for (env_local[i], x[i], y[i]) at the
 * env's current episode (what a procedural lookup of that cell uses). */
int plume_field_noise_at(const plume_env_config* cfg, const plume_env_state* st, const int32_t* env_local,
                         const int32_t* x, const int32_t* y, int32_t n, float* z_out, float* u_out, void* stream);

/* conc_field[x[i], y[i]] and tke_field[x[i], y[i]] of env i (int32 x/y [n_envs], clipped to the grid) as double
 * [n_envs]; either output may be NULL.  Works in every field mode (the accessor behind env.conc_field[...] in
 * evaluate_with_lstm.py:67-68). */
int plume_field_at(const plume_env_config* cfg, const plume_env_state* st, const int32_t* x, const int32_t* y,
                   double* conc, double* tke, void* stream);

/* ---- P3 _get_obs, environment.py:71-87 ------------------------------------------------- */
/* obs float[n_envs][6]. */
int plume_env_observe(const plume_env_config* cfg, const plume_env_state* st, float* obs, void* stream);

/* ---- P2 MethaneEnv.step, environment.py:89-178 ----------------------------------------- */
/* One lockstep step of all n_envs.
 *  actions      int32[n_envs]
 *  step_noise   double[n_envs][2] = the randn(2) of :108, or NULL to draw from Philox TAG_STEP
 *  obs          float[n_envs][6]   observation after the step (after the reset if auto-reset fired)
 *  reward       double[n_envs]     (the reference returns float64)
 *  done,reached uint8[n_envs]
 *  info         float[5][n_envs]   concentration_reward, explore_reward, move_penalty, tke_penalty,
 *                                  boundary_penalty (may be NULL)
 *  final_obs    float[n_envs][6]   pre-reset observation, written when auto-reset (may be NULL)
 *  noise_out    double[n_envs][2]  the step noise actually used (may be NULL)
 */
int plume_env_step(const plume_env_config* cfg, const plume_env_state* st, const int32_t* actions,
                   const double* step_noise, uint32_t flags, float* obs, double* reward, uint8_t* done,
                   uint8_t* reached, float* info, float* final_obs, double* noise_out, void* stream);

/* ---- P4 PPOActorCritic.forward, model.py:38-46 ----------------------------------------- */
/* Flat parameter layout (floats), each block padded to 4 floats:
 *   feature.0.weight[256][6] feature.0.bias[256] feature.1.weight[256] feature.1.bias[256]
 *   feature.3.weight[128][256] feature.3.bias[128] feature.4.weight[128] feature.4.bias[128]
 *   actor.weight[5][128] actor.bias[5](+3) critic.weight[1][128] critic.bias[1](+3) */
#define PLUME_MLP_IN 6
#define PLUME_MLP_H1 256
#define PLUME_MLP_H2 128
#define PLUME_OFF_W1 0
#define PLUME_OFF_B1 1536
#define PLUME_OFF_G1 1792
#define PLUME_OFF_BE1 2048
#define PLUME_OFF_W2 2304
#define PLUME_OFF_B2 35072
#define PLUME_OFF_G2 35200
#define PLUME_OFF_BE2 35328
#define PLUME_OFF_WA 35456
#define PLUME_OFF_BA 36096
#define PLUME_OFF_WC 36104
#define PLUME_OFF_BC 36232
#define PLUME_MLP_PARAMS 36236   /* padded size of the flat buffer */

/* probs float[B][5], value float[B]; nan_flag int32[1] is set to 1 if any logit is NaN
 * (model.py:41-43 raises RuntimeError("NaN in model output")). */
int plume_policy_forward(const float* params, const float* x, int32_t batch, float* probs, float* value,
                         int32_t* nan_flag, void* stream);

/* P4s: forward + Categorical sample/log_prob (train_ppo2.0.py:161-162,185).
 *  uniforms float[B] in [0,1) for the inverse-CDF draw, or NULL: Philox TAG_ACT keyed by the
 *  env state (then B must equal st->n_envs); forced_actions int32[B] (may be NULL) overrides the
 *  draw (action-trace replay); flags: PLUME_FLAG_GREEDY. */
int plume_policy_act(const plume_env_config* cfg, const plume_env_state* st, const float* params,
                     const float* obs, int32_t batch, const float* uniforms, const int32_t* forced_actions,
                     uint32_t flags, int32_t* actions, float* logp, float* value, float* probs,
                     int32_t* nan_flag, void* stream);

/* ---- P4L LSTM stop heads ---------------------------------------------------------------- */
/* V2.1 PeakAndStopPredictor (evaluate_with_lstm.py:11-27): single-layer LSTM(1->H) from zero
 * state over windows float[B][T] (already divided by 100), torch parameter layouts:
 * w_ih[4H][1], w_hh[4H][H], b_ih[4H], b_hh[4H] (gate order i,f,g,o), fc_peak w[H] b[1],
 * fc_stop w[H] b[1].  Outputs peak[B], stop_prob[B]. H in {32,64,128}. */
int plume_lstm_stop_head(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                         const float* w_peak, const float* b_peak, const float* w_stop, const float* b_stop,
                         int32_t hidden, const float* windows, int32_t batch, int32_t steps, float* peak,
                         float* stop_prob, void* stream);

/* Generic multi-layer LSTM (input size 1) over windows float[B][T] from zero state; writes the
 * top layer's last hidden state float[B][H] (V2.0 ConcentrationThresholdPredictor,
 * PPOV2.0/model.py:203-240).  params: per layer l: w_ih (4H x in_l), w_hh (4H x H), b_ih, b_hh
 * concatenated in that order. */
int plume_lstm_forward(const float* params, int32_t layers, int32_t hidden, const float* windows,
                       int32_t batch, int32_t steps, float* h_out, void* stream);

/* FC head of the V2.0 ConcentrationThresholdPredictor (PPOV2.0/model.py:213-220, eval mode): h float[B][H] (the
 * last hidden state from plume_lstm_forward) -> Linear(H,64) -> LayerNorm(64) -> ReLU -> Linear(64,1) -> out[B].
 * w1 [64][H], b1 [64], ln_weight/ln_bias [64], w2 [64], b2 [1] (torch layouts). */
int plume_threshold_head(const float* h, int32_t batch, int32_t hidden, const float* w1, const float* b1,
                         const float* ln_weight, const float* ln_bias, const float* w2, const float* b2, float* out,
                         void* stream);

/* P4t trend features, model.py:113-127: conc float[B][W] (raw 0..100), last position float[B][2],
 * source double[B][2] -> out float[B][4] = {label, trend_score, dist_score, conc_score}. */
int plume_trend_features(const float* conc, int32_t batch, int32_t window, const float* pos_last,
                         const double* src, double conc_peak, float* out, void* stream);

/* ---- fused rollout: train_ppo2.0.py:156-192 + evaluate_with_lstm.py:61-82 ---------------- */
typedef struct plume_lstm_params {        /* V2.1 stop head, DEVICE pointers; hidden = 0 disables the head */
    int32_t hidden;
    int32_t window;                       /* evaluate_with_lstm.py:42 */
    float threshold;                      /* evaluate_with_lstm.py:77 */
    const float *w_ih, *w_hh, *b_ih, *b_hh, *w_peak, *b_peak, *w_stop, *b_stop;
} plume_lstm_params;

typedef struct plume_rollout_buffers {    /* DEVICE pointers, [T][N] row-major unless noted */
    float* obs;            /* [T][N][6]  state fed to the policy at step t */
    int32_t* actions;
    float* rewards;        /* float(reward), model.py:152 */
    float* values;
    float* log_probs;
    float* dones;          /* float(done), model.py:155 */
    uint8_t* reached;      /* trajectory[-1]['reached'], environment.py:167 */
    float* stop_prob;      /* LSTM stop probability (0 while the window is not full), may be NULL */
    uint8_t* stop_flag;    /* stop_prob > threshold, may be NULL */
    float* peak_pred;      /* may be NULL */
    float* trend;          /* [T][N][4] trend features over the stop window, may be NULL */
    float* info;           /* [T][5][N] reward components, may be NULL */
    int32_t* episode_idx;  /* [T][N] episode index the transition belongs to, may be NULL */
    const int32_t* forced_actions; /* [T][N] action-trace replay, may be NULL */
    const double* step_noise;      /* [T][N][2] injected randn(2), may be NULL */
    double* noise_out;             /* [T][N][2] the randn(2) actually used, may be NULL */
    float* conc_window;    /* [N][window] ring of obs[2] inside the current episode (persistent state) */
    int32_t* window_fill;  /* [N] samples in the ring (persistent state) */
    float* last_obs;       /* [N][6] observation each env will act on next (persistent state, in/out) */
    /* inputs of the deferred stop head (PLUME_FLAG_DEFER_STOP_HEAD), [T][N]: */
    float* conc_sample;    /* the value pushed into the window at step t: obs[2] after the step, before a reset
                            * (evaluate_with_lstm.py:67-70) */
    uint8_t* fill_t;       /* samples of the current episode in the window after the push (saturates at window) */
    double* src_dist;      /* ||agent_pos - source_pos|| after the step (environment.py:155), for the trend label */
    /* trajectory logging in the training_data.nc layout (train_ppo2.0.py:166-170,207-233), optional: */
    float* pos_out;        /* [T][N][2] agent_pos after the step, before a reset */
    float* src_out;        /* [T][N][2] source_pos of the episode, written at its last transition only */
    float* conc_out;       /* [T][N] conc_field[int(x), int(y)] at the position after the step, as float32: what the
                            * reference driver logs per step (train_ppo2.0.py:167-173); may be NULL */
    double* eval_ring;     /* [N][10] evaluator stop tests: the last 10 samples of every env, oldest first, carried across
                            * segments (STOP_THRESHOLD: concentrations; STOP_FIXED: (x, y) float pairs); may be NULL */
    const double* stop_threshold;   /* [N] STOP_THRESHOLD: 0.95 x predicted source concentration, NaN = none yet */
    uint8_t* flag_code;    /* [T][N] bit 0 = done, bit 1 = reached: what the curriculum (and its multi-GPU
                            * all-gather, 1 B per transition) consumes; may be NULL */
} plume_rollout_buffers;

/* T lockstep iterations of: policy forward + sample, env step, stop head, auto-reset; one
 * persistent kernel, procedural field mode only. nan_flag as in plume_policy_forward. */
int plume_rollout(const plume_env_config* cfg, const plume_env_state* st, const float* mlp_params,
                  const plume_lstm_params* lstm, const plume_rollout_buffers* buf, int32_t horizon,
                  uint32_t flags, int32_t* nan_flag, void* stream);

/* Deferred stop head + trend features of a whole segment (PPOV2.1/evaluate_with_lstm.py:73-80 and
 * model.py:113-127 for every (t, env) of the [T][N] rollout): the window of (t, env) is
 * conc_sample[t-W+1..t][env], continued into window_in [N][W] (the ring carried over from the previous
 * segment) for t < W-1; it is evaluated when fill_t[t][env] >= W, else the outputs are 0.  Writes
 * stop_prob/stop_flag/peak_pred [T][N], trend [T][N][4] (each may be NULL) and the ring for the next
 * segment into window_out [N][W] (must not alias window_in).  Because the stop decision does not feed
 * back into the rollout unless PLUME_FLAG_STOP_TERMINATES is set, taking the head out of the lockstep
 * loop changes no result; it turns 20 latency-bound cell steps per env-step into a throughput kernel. */
/* kernel_path: which of the library's two kernel families runs -- an explicit argument of the caller, never an
 * environment switch.  PLUME_KERNEL_AUTO is a documented shape rule (tcgen05 where the shape has a tensor-core kernel:
 * hidden 32/64/128/256 for the stop head, >= 1024 samples for plume_ppo_grad), PLUME_KERNEL_TENSOR / PLUME_KERNEL_SIMT
 * name one (tests compare the two; SIMT = fp32 FMA on the CUDA cores, the in-loop head's exact arithmetic). */
#define PLUME_KERNEL_AUTO 0
#define PLUME_KERNEL_TENSOR 1
#define PLUME_KERNEL_SIMT 2
int plume_stop_head_segment(const plume_lstm_params* lstm, const float* conc_sample, const uint8_t* fill_t,
                            const double* src_dist, int32_t horizon, int32_t n_envs, const float* window_in,
                            float* window_out, double conc_peak, float* stop_prob, uint8_t* stop_flag,
                            float* peak_pred, float* trend, int32_t kernel_path, void* stream);

/* ---- P5 GAE + normalisation, train_ppo2.0.py:17-39 -------------------------------------- */
/* Per-env reverse scan over [T][N] (the reference's quirks kept: self-bootstrap at T-1, mask with
 * dones[t+1]); writes raw advantages and accumulates {sum, sum of squares, count} of them into
 * stats double[3] (caller zeroes; all-reduce it across ranks for global statistics). */
int plume_gae_scan(const float* rewards, const float* values, const float* dones, int32_t horizon,
                   int32_t n_envs, double gamma, double lam, float* advantages, double* stats, void* stream);
/* adv = (adv-mean)/(std_unbiased+1e-6) (std := 1 if <1e-6 or NaN); returns = adv + values (sic, :39). */
int plume_gae_normalise(float* advantages, const float* values, int64_t count, const double* stats,
                        float* returns, void* stream);

/* P5' the GAE variants of the older drivers (kept as flags, not parity targets of the V2.x update):
 *   PLUME_GAE_QUIRK      train_ppo2.0.py:17-39 (= plume_gae_scan / plume_gae_normalise)
 *   PLUME_GAE_BOOTSTRAP  PPOV1.1/train_ppo1.0.py:66-89 (also V1.0, GAIL): bootstraps the last step with
 *                        last_values[n] = V(next_state), masks with dones[t+1], returns = RAW advantage + value,
 *                        normalises with std + 1e-8
 *   PLUME_GAE_V12        PPOV1.2: next value masked with dones[t], no bootstrap at the last step, returns =
 *                        normalised advantage + value, std + 1e-8 */
#define PLUME_GAE_QUIRK 0
#define PLUME_GAE_BOOTSTRAP 1
#define PLUME_GAE_V12 2
int plume_gae_scan_variant(const float* rewards, const float* values, const float* dones, const float* last_values,
                           int32_t horizon, int32_t n_envs, double gamma, double lam, int32_t variant,
                           float* advantages, double* stats, void* stream);
int plume_gae_normalise_variant(float* advantages, const float* values, int64_t count, const double* stats,
                                int32_t variant, float* returns, void* stream);

/* ---- P6/P7 PPO minibatch update, train_ppo2.0.py:42-87 ---------------------------------- */
typedef struct plume_ppo_batch {          /* DEVICE pointers over the flat [M] transition set */
    int64_t total;                        /* M = T*N */
    const float* obs;                     /* [M][6] */
    const int32_t* actions;
    const float* old_log_probs;
    const float* advantages;
    const float* returns;
    const float* old_values;
    const float* packed;                  /* optional [M][12] sample records written by plume_ppo_pack (NULL = gather
                                           * from the six arrays above) */
} plume_ppo_batch;

/* Interleaves the transition set into 48-byte records {obs[6], advantage, return, old value, old log-prob, action
 * (int bits), 0}: the gradient kernel then gathers a permuted sample with three 16-byte copies (1.5 DRAM sectors)
 * instead of eleven 4-byte ones (7 sectors).  Call after the advantages are final; batch->packed is not read. */
int plume_ppo_pack(const plume_ppo_batch* batch, float* packed, void* stream);

/* Forward + loss + backward of one minibatch: samples are perm[i] for i in [mb_start, mb_start+mb_size)
 * if perm (int64[M] device) is given, else the stateless bijection keyed by (perm_seed, epoch).
 * Accumulates d(loss)/d(params) into grads float[PLUME_MLP_PARAMS] (caller zeroes) and
 * {loss, policy_loss, value_loss, entropy} sums into loss_out double[4] scaled by 1/mb_size_global.
 * mb_size_global is the divisor of the means (= mb_size on one GPU; sum over ranks otherwise).
 * kernel_path: PLUME_KERNEL_AUTO / _TENSOR / _SIMT (above). */
int plume_ppo_grad(const float* params, const plume_ppo_batch* batch, const int64_t* perm, uint64_t perm_seed,
                   int32_t epoch, int64_t mb_start, int64_t mb_size, int64_t mb_size_global, float clip_eps,
                   float entropy_beta, float* grads, double* loss_out, int32_t* nan_flag, void* workspace,
                   int64_t workspace_bytes, int32_t kernel_path, void* stream);
/* The whole optimiser loop of _update_model (train_ppo2.0.py:42-87) in one call: `epochs` passes over the
 * batch->total transitions in minibatches of mb_size; per step: zero grads (folded into the gradient launch's operand
 * preparation on the tensor-core path), plume_ppo_grad (divisor mb * world),
 * plume_clip_adam -- or plume_allreduce_clip_adam when comm != NULL.  perms: int64 [epochs][M] (NULL = the stateless
 * bijection keyed by (perm_seed, epoch)); first_step = the optimiser's step number of the first step (1-based);
 * losses double[epochs * ceil(M / mb_size)][4] (zeroed by the caller); grad_norm_out (may be NULL) receives the last
 * step's gradient norm. */
int plume_ppo_update(float* params, float* grads, float* exp_avg, float* exp_avg_sq, const plume_ppo_batch* batch,
                     const int64_t* perms, uint64_t perm_seed, int32_t epochs, int64_t mb_size, int32_t world,
                     float clip_eps, float entropy_beta, float max_norm, float lr, float beta1, float beta2, float eps,
                     int32_t first_step, void* comm, double* losses, float* grad_norm_out, int32_t* nan_flag,
                     void* workspace, int64_t workspace_bytes, int32_t kernel_path, void* stream);
/* bytes of workspace plume_ppo_grad needs for a minibatch of mb_size samples */
int64_t plume_ppo_workspace_bytes(int64_t mb_size);

/* clip_grad_norm_(0.5) + Adam (train_ppo2.0.py:86-87,113): global L2 norm over grads, scale by
 * max_norm/(norm+1e-6) if norm > max_norm, Adam(lr, b1, b2, eps) bias-corrected with `step`
 * (1-based).  grad_norm_out float[1] (may be NULL). */
int plume_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t n,
                    float max_norm, float lr, float beta1, float beta2, float eps, int32_t step,
                    float* grad_norm_out, void* stream);

/* ---- the exchange steps of a data-parallel iteration over NVLink peer memory (one process per GPU) ----------
 * No host-launched collective is on the path.  Every rank owns one block that all peers have mapped (CUDA IPC);
 * a rank publishes into its own block, signals with system-scope release stores into the peers' flag arrays,
 * waits for the flags of all ranks and reads the peers' blocks.
 *   plume_comm_create   allocates this rank's block (room for two gradients of n_params floats and two flag-code
 *                       segments of code_bytes bytes), returns an opaque communicator and its 64-byte IPC handle
 *   plume_comm_connect  maps the peers' blocks: all_handles = the handles of ranks 0..world-1, 64 bytes each
 *                       (exchange them with any host-side all-gather)
 *   plume_allreduce_clip_adam  the update's exchange step (train_ppo2.0.py:85-87 on N data-parallel ranks) as ONE
 *                       kernel: all-reduce(sum) of grads in rank order (bitwise identical on every rank), then
 *                       plume_clip_adam from registers; grads receives the reduced gradient.
 *   plume_comm_allreduce_small  values[count <= 32] doubles <- sum over the ranks (the advantage statistics,
 *                       train_ppo2.0.py:34-38 over the global batch)
 *   plume_comm_publish_codes    copies this rank's [T][N] flag codes into its block and signals the peers (may run
 *                       on a side stream while the stop-head kernel runs)
 *   plume_curriculum_update_peer  waits for every rank's codes, then plume_curriculum_update_packed with the flags
 *                       read straight from the peers' blocks (no gathered copy)
 *   Every rank must make the same sequence of calls.  Spins are bounded: on a timeout the kernel applies NOTHING
 *   (parameters, moments, curriculum stay as they were), records the error and every later exchange kernel returns
 *   at once.
 *   plume_comm_error    HOST out 0 ok / 1 a peer did not arrive / 2 grid barrier timeout (synchronises the stream)
 *   plume_comm_error_async  queues a copy of that word into PINNED host memory behind the work in `stream`
 *   plume_comm_reset    collective recovery: call on every rank between two host barriers, nothing in flight */
int plume_comm_create(int32_t world, int32_t rank, int32_t n_params, int64_t code_bytes, void** comm_out,
                      uint8_t* handle_out);
int plume_comm_connect(void* comm, const uint8_t* all_handles);
int plume_comm_destroy(void* comm);
int plume_comm_error(void* comm, int32_t* error_out, void* stream);
int plume_comm_error_async(void* comm, int32_t* pinned_error_out, void* stream);
int plume_comm_reset(void* comm);
int plume_allreduce_clip_adam(void* comm, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int32_t n,
                              float max_norm, float lr, float beta1, float beta2, float eps, int32_t step,
                              float* grad_norm_out, void* stream);
int plume_comm_allreduce_small(void* comm, double* values, int32_t count, void* stream);
int plume_comm_publish_codes(void* comm, const uint8_t* flag_code, int64_t bytes, void* stream);
int plume_curriculum_update_peer(void* comm, int32_t horizon, int32_t n_envs, double* state, double* curriculum,
                                 double initial_radius, double min_radius, double radius_decay,
                                 double success_threshold, int32_t window, double decay_factor,
                                 double* window_radius_out, void* stream);

/* The index permutation plume_ppo_grad uses when perm == NULL: out int64[count] = positions
 * [start, start+count) of the bijection of [0,total) keyed by (seed, epoch). */
int plume_permutation(int64_t total, uint64_t seed, int32_t epoch, int64_t start, int64_t count, int64_t* out,
                      void* stream);

/* ---- N3: supervised training of the V2.1 stop head (PPOV2.1/train_lstm.py) ------------------------------------
 * Flat parameter vector in torch named_parameters() order of PeakAndStopPredictor(hidden 32):
 *   lstm.weight_ih_l0[128] lstm.weight_hh_l0[128][32] lstm.bias_ih_l0[128] lstm.bias_hh_l0[128]
 *   fc_peak.weight[32] fc_peak.bias[1] fc_stop.0.weight[32] fc_stop.0.bias[1]                                  */
#define PLUME_LSTM_TRAIN_PARAMS 4546
/* TrajectoryDataset._preprocess (train_lstm.py:28-65) for the episodes listed in episode_ids (int32[n_selected],
 * each with at least `window` logged steps; the host picks them, train_lstm.py:40): from rows of the
 * training_data.nc variables concentration / x / y (float[episodes][max_steps]) and source_x / source_y
 * (float[episodes]) writes two samples per episode -- features float[2 n_selected][window] = conc[:window] / 100
 * (twice: the reference's "positive" window is the same first window), labels float[2 n_selected][2] =
 * {conc[window-1] / 100, 0} and {conc[window-1] / 100, ||pos[window-1] - source|| <= stop_radius}. */
int plume_lstm_dataset(const float* conc, const float* x, const float* y, const float* src_x, const float* src_y,
                       int32_t max_steps, const int32_t* episode_ids, int32_t n_selected, int32_t window,
                       float stop_radius, float* features, float* labels, void* stream);
/* One epoch of train_lstm.py:108-121: for every minibatch of `batch_size` samples taken in `order` (int32
 * [n_samples] sample ids, NULL = identity; the last minibatch may be short) ONE kernel does forward, BPTT,
 * loss = MSE(peak) + BCE(stop), clip_grad_norm_(max_norm) and the AdamW step (optimiser step numbers
 * first_step, first_step+1, ...).  batch_losses float[ceil(n_samples / batch_size)] receives each minibatch's
 * mean loss, grad_norms (may be NULL) the gradient norms before clipping, grad_out (may be NULL,
 * float[PLUME_LSTM_TRAIN_PARAMS]) the unclipped gradient of the last minibatch.  workspace: device memory of
 * plume_lstm_train_workspace_bytes(batch_size) bytes, 256-byte aligned.  hidden must be 32. */
int64_t plume_lstm_train_workspace_bytes(int32_t batch_size);
int plume_lstm_train_epoch(float* params, float* exp_avg, float* exp_avg_sq, int32_t hidden, const float* features,
                           const float* labels, const int32_t* order, int32_t n_samples, int32_t window,
                           int32_t batch_size, float max_norm, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int32_t first_step, void* workspace, int64_t workspace_bytes,
                           float* batch_losses, float* grad_norms, float* grad_out, void* stream);

/* ---- tensor-core GEMM building block ----------------------------------------------------- */
/* C[M][N] = A[M][K] . B[N][K]^T, fp32 in/out, 3xTF32 on tcgen05 (fp32-grade accuracy); N in {128,256},
 * K a multiple of 32.  The GEMM-shaped parts of the PPO update (model.py:23 feature.3, forward and
 * backward, train_ppo2.0.py:54,85) run on this path; exported so it can be tested on its own. */
int plume_tc_gemm(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, void* stream);
/* The same GEMM on tcgen05.mma kind::f16 with the two-term fp16 split (x = hi + lo / s): twice the MAC rate of TF32 and
 * half the operand bytes at fp32-grade accuracy.  scaled_lo != 0: s = 2^11, cross terms in their own TMEM accumulator
 * (no fp16 underflow of lo; the forward GEMM of the update); scaled_lo == 0: s = 1, one accumulator (inputs should be
 * O(1): the backward GEMMs, whose operands the kernel pre-scales); scaled_lo == 2: as 0, with the A operand staged
 * MN-major (the descriptor form the update kernel's G2 reads dz2 with).  K a multiple of 64. */
int plume_tc_gemm_f16(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t scaled_lo,
                      void* stream);

/* ---- P8 curriculum, model.py:188-221 ----------------------------------------------------- */
/* Applies PPOTrainer.update once per finished episode of a [T][N] segment in canonical order
 * (step-major, then env index), on the device.  state double[8 + window]:
 * {trainer_radius, trainer_explore_bonus, env_radius, env_explore_bonus, history_len,
 *  history_successes, episodes_total, successes_total, ring...}; curriculum double[2] receives the
 * values the next resets latch. */
int plume_curriculum_update(const float* dones, const uint8_t* reached, int32_t horizon, int32_t n_envs,
                            double* state, double* curriculum, double initial_radius, double min_radius,
                            double radius_decay, double success_threshold, int32_t window,
                            double decay_factor, void* stream);

/* Same rule on packed flags of ALL ranks: flag_code uint8 [world][T][N] (bit 0 done, bit 1 reached; the layout an
 * all-gather of the ranks' [T][N] arrays produces); canonical order = step-major, then global env id.
 * window_radius_out (may be NULL) double[PLUME_CURRICULUM_MAX_WINDOWS + 2]: [0] = length of the carried partial
 * window at entry, [1] = number of radii that follow, [2 + b] = the trainer's radius in force for the episodes whose
 * ordinal (carried length + canonical position in this segment) falls into window b -- the 'Current_Radius' the
 * reference logs per episode (train_ppo2.0.py:247). */
#define PLUME_CURRICULUM_MAX_WINDOWS 8192
int plume_curriculum_update_packed(const uint8_t* flag_code, int32_t horizon, int32_t n_envs, int32_t world,
                                   double* state, double* curriculum, double initial_radius, double min_radius,
                                   double radius_decay, double success_threshold, int32_t window,
                                   double decay_factor, double* window_radius_out, void* stream);

/* ---- N1: evaluator bookkeeping (evaluate_with_lstm.py:83-113 of PPOV2.1 / PPOV2.0, evaluate_model.py:70-82) ----------
 * Records, for every env that has not finished its evaluation episode yet, the first transition of the segment that ends
 * it (stop_flag or done): steps = base_step + t + 1, early = stopped by the stop test, stop_step, deviation =
 * ||agent_pos - source_pos|| there (buf->src_dist).  pending_threshold (may be NULL): the STOP_THRESHOLD test of the
 * segment's last step, evaluated here against the refreshed threshold on eval_ring. */
int plume_eval_collect(const plume_rollout_buffers* buf, int32_t horizon, int32_t n_envs, int32_t base_step,
                       uint8_t* finished, int32_t* steps, uint8_t* early, int32_t* stop_step, double* deviation,
                       const double* pending_threshold, void* stream);

/* ---- N2: trajectory / per-episode logging (training_data.nc + training_results.csv layouts) ------------------
 * NetCDFWriter.write_episode_data (PPOV2.1/model.py:351-419) and the per-episode statistics of
 * train_ppo2.0.py:128-134,140-180,194-199,236-248, assembled on the device from one [T][N] rollout segment.
 * Episodes are appended in canonical order (step-major, then env id); an episode that spans segments is kept in
 * the env's carry row until it closes.  Tables (DEVICE, caller-owned; x / y / conc NaN-filled once by the caller):
 * E = max_episodes, S = max_steps. */
typedef struct plume_traj_log {
    int32_t max_episodes, max_steps, n_envs, reserved;
    float* x;               /* [E][S] agent x after each step */
    float* y;
    float* conc;            /* [E][S] conc_field[int(x), int(y)] after each step */
    int32_t* steps;         /* [E] episode length */
    float* source;          /* [E][2] */
    uint8_t* success;       /* [E] trajectory[-1]['reached'] */
    double* radius;         /* [E] trainer.current_radius when the episode ended ('Current_Radius') */
    double* sums;           /* [E][6] total reward, then the five info components, summed in time order */
    float* final_conc;      /* [E] 'Final_Conc': conc at the final cell if reached, else 0 */
    float* c_x;             /* [N][S] carry rows of the open episode of every env */
    float* c_y;
    float* c_conc;
    double* c_sums;         /* [N][6] */
    int32_t* c_len;         /* [N] steps of the open episode so far */
    int32_t* count;         /* [1] episodes logged so far (clamped to E) */
} plume_traj_log;
int64_t plume_trajectory_workspace_bytes(int32_t horizon, int32_t n_envs);
/* buf: dones, reached, rewards, info (may be NULL), pos_out, src_out, conc_out of the segment.  comm (may be NULL):
 * with several ranks, the communicator whose flag codes of this segment are published -- the global episode ordinal
 * (for the radius lookup) then counts the other ranks' episodes too.  window / window_radius: the curriculum window
 * length and the array plume_curriculum_update_packed / _peer wrote for THIS segment (NULL: radius_fallback). */
int plume_trajectory_log(const plume_traj_log* log, const plume_rollout_buffers* buf, int32_t horizon, void* comm,
                         int32_t window, const double* window_radius, double radius_fallback, void* workspace,
                         int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLUME_B200_H */
