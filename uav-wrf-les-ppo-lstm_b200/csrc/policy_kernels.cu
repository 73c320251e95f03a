// policy_kernels.cu -- K3: PPOActorCritic.forward (model.py:38-46) and the Categorical
// sample / log_prob of the rollout (train_ppo2.0.py:158-162,185) for a batch of observations.
// One CTA per SM (weights stay in shared memory), grid-stride over 32-sample tiles.
#include "mlp_tc_tile.cuh"

namespace plume {

struct ActArgs {
    const float* uniforms;        // [B] or null
    const int32_t* forced;        // [B] or null
    int32_t* actions;             // [B] or null (forward-only call)
    float* logp;                  // [B] or null
    uint32_t flags;
    bool use_env_rng;             // draw from Philox TAG_ACT keyed by the env state
};

__global__ void __launch_bounds__(kMlpThreads, 1)
policy_kernel(const float* __restrict__ params, const float* __restrict__ x, int batch, float* __restrict__ probs_out,
              float* __restrict__ value_out, int32_t* nan_flag, ActArgs act, Cfg c, plume_env_state st) {
    extern __shared__ __align__(16) float sm[];
    policy_load_weights(sm, params);
    const int tid = threadIdx.x;
    const int tiles = (batch + kTileM - 1) / kTileM;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int base = tile * kTileM;
        __syncthreads();   // previous tile fully consumed before x/out are overwritten
        {
            const int s = tid >> 3, k = tid & 7;
            const int row = base + s;
            sm[PolicySmem::x + tid] = (k < 6 && row < batch) ? x[(size_t)row * 6 + k] : 0.0f;
        }
        policy_forward_tile(sm);
        if (tid < kTileM && base + tid < batch) {
            const int row = base + tid;
            const float* o = sm + PolicySmem::out + tid * 8;
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 5; ++k) bad |= isnan(o[k]);
            if (bad) atomicExch(nan_flag, 1);                         // model.py:41-43
            float p[5];
            softmax5(o, p);
            if (probs_out) {
#pragma unroll
                for (int k = 0; k < 5; ++k) probs_out[(size_t)row * 5 + k] = p[k];
            }
            if (value_out) value_out[row] = o[5];
            if (act.actions) {
                float u = 0.0f;
                if (act.uniforms) u = act.uniforms[row];
                else if (act.use_env_rng) {
                    const EnvRegs e = load_env(st, row);
                    u = action_uniform(c, (uint32_t)(st.env_id_base + row), e);
                }
                float lp;
                const int a = categorical_pick(p, u, (act.flags & PLUME_FLAG_GREEDY) != 0,
                                               act.forced ? act.forced[row] : -1, lp);
                act.actions[row] = a;
                if (act.logp) act.logp[row] = lp;
            }
        }
    }
}

static int launch_policy(const float* params, const float* x, int batch, float* probs, float* value,
                         int32_t* nan_flag, const ActArgs& act, const Cfg& c, const plume_env_state& st,
                         cudaStream_t s) {
    static bool configured = false;
    const int smem = PolicySmem::total * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return fail("policy kernel: cannot reserve %d B of shared memory", smem);
        configured = true;
    }
    const int tiles = (batch + kTileM - 1) / kTileM;
    int grid = sm_count();
    if (grid <= 0) return fail("no CUDA device");
    if (tiles < grid) grid = tiles;
    policy_kernel<<<grid, kMlpThreads, smem, s>>>(params, x, batch, probs, value, nan_flag, act, c, st);
    if (cudaGetLastError() != cudaSuccess) return fail("policy kernel launch failed");
    return 0;
}

}  // namespace plume

using namespace plume;

extern "C" int plume_policy_forward(const float* params, const float* x, int32_t batch, float* probs, float* value,
                                    int32_t* nan_flag, void* stream) {
    PLUME_CHECK_ARG(params && x && nan_flag, "null pointer");
    if (batch <= 0) return 0;
    ActArgs act{};
    Cfg c{};
    plume_env_state st{};
    return launch_policy(params, x, batch, probs, value, nan_flag, act, c, st, as_stream(stream));
}

extern "C" int plume_policy_act(const plume_env_config* cfg, const plume_env_state* st, const float* params,
                                const float* obs, int32_t batch, const float* uniforms,
                                const int32_t* forced_actions, uint32_t flags, int32_t* actions, float* logp,
                                float* value, float* probs, int32_t* nan_flag, void* stream) {
    PLUME_CHECK_ARG(params && obs && nan_flag && actions, "null pointer");
    if (batch <= 0) return 0;
    ActArgs act{};
    act.uniforms = uniforms;
    act.forced = forced_actions;
    act.actions = actions;
    act.logp = logp;
    act.flags = flags;
    Cfg c{};
    plume_env_state s{};
    if (!uniforms && !forced_actions && !(flags & PLUME_FLAG_GREEDY)) {
        PLUME_CHECK_ARG(cfg && st, "Philox action draw needs the env config/state");
        PLUME_CHECK_ARG(batch == st->n_envs, "Philox action draw needs batch == n_envs");
        act.use_env_rng = true;
        c = make_cfg(*cfg);
        s = *st;
    }
    return launch_policy(params, obs, batch, probs, value, nan_flag, act, c, s, as_stream(stream));
}
