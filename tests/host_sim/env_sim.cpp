// env_sim.cpp -- TEST INFRASTRUCTURE: compiles the product's __host__ __device__ core
// (csrc/plume_core.h) with g++ so that the exact per-env logic the CUDA kernels run can be
// checked against the oracle on a machine without a GPU.  Never linked into libplume_b200.
#include <cstring>

#include "../../uav-wrf-les-ppo-lstm_b200/csrc/plume_core.h"

using namespace plume;

// variant of the step the next sim_step_f64 calls run:
//   0  previous-cell concentration passed in, no host tables, sqrt distance
//   1  previous-cell concentration evaluated on demand, host tables, reached test without the sqrt
//   2  as 0 with the generic IEEE divisions (fastdiv off)
static int g_variant = 0;
static const float* g_step_frac = nullptr;
static const float* g_visit_denom = nullptr;

static Cfg sim_cfg(const plume_env_config* cfg) {
    Cfg c = make_cfg(*cfg);
    if (g_variant == 1) {
        c.step_frac_tab = g_step_frac;
        c.visit_denom_tab = g_visit_denom;
    }
    if (g_variant == 2) c.fastdiv = 0;
    return c;
}

extern "C" {

void sim_set_variant(int v, const float* step_frac, const float* visit_denom) {
    g_variant = v;
    g_step_frac = step_frac;
    g_visit_denom = visit_denom;
}

int sim_fastdiv_enabled(const plume_env_config* cfg) { return make_cfg(*cfg).fastdiv; }

// out[i] = a[i] / b through the three-instruction constant division (double and float)
void sim_div_const(const double* a, int n, double b, double* out) {
    const double y = 1.0 / b;
    for (int i = 0; i < n; ++i) out[i] = ddiv_const(a[i], b, y);
}
void sim_div_G(const plume_env_config* cfg, const float* a, int n, float* out) {
    const Cfg c = make_cfg(*cfg);
    for (int i = 0; i < n; ++i) out[i] = div_G(c, a[i]);
}
int sim_reciprocal_ok(double b) { return reciprocal_ok(b) ? 1 : 0; }

void sim_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    const U4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void sim_feistel(uint64_t n, uint64_t seed, uint32_t epoch, int64_t* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = (int64_t)feistel_permute(i, n, seed, epoch);
}

void sim_field_noise(const plume_env_config* cfg, uint32_t gid, uint32_t episode, const int32_t* x, const int32_t* y,
                     int n, float* z, float* u) {
    const Cfg c = make_cfg(*cfg);
    for (int i = 0; i < n; ++i) field_noise(c, gid, episode, x[i], y[i], z[i], u[i]);
}

void sim_source(const plume_env_config* cfg, uint32_t gid, uint32_t episode, double* src) {
    const Cfg c = make_cfg(*cfg);
    EnvRegs e{};
    e.episode = episode - 1;
    alignas(16) uint16_t vis[PLUME_VISIT_STRIDE];
    env_reset(c, gid, e, vis, nullptr, 50.0, 0.6);
    src[0] = e.sx; src[1] = e.sy;
}

// one lockstep step over n envs with float64 materialised fields [n][G][G]
void sim_step_f64(const plume_env_config* cfg, int n, float* pos_x, float* pos_y, const double* src_x,
                  const double* src_y, int32_t* step, const int32_t* episode, uint16_t* visited,
                  const double* radius, const double* ebonus, const double* conc_field, const double* tke_field,
                  const int32_t* actions, const double* noise, float* obs, double* reward, uint8_t* done,
                  uint8_t* reached, double* info) {
    const Cfg c = sim_cfg(cfg);
    MaterialisedField<double> f{conc_field, tke_field};
    for (int i = 0; i < n; ++i) {
        EnvRegs e;
        e.px = pos_x[i]; e.py = pos_y[i]; e.sx = src_x[i]; e.sy = src_y[i];
        e.step = step[i]; e.episode = (uint32_t)episode[i]; e.radius = radius[i]; e.ebonus = ebonus[i];
        e.last_move = 0;
        int px, py;
        cell32_of(c, e, px, py);
        double pc, pt;
        f.eval(c, i, 0, e.episode, e.sx, e.sy, px, py, pc, pt);
        StepResult r;
        if (g_variant == 1)
            env_step<false>(c, f, i, (uint32_t)i, e, visited + (size_t)i * PLUME_VISIT_STRIDE, actions[i], noise[2 * i],
                            noise[2 * i + 1], false, 0.0, f.eval_tke(c, i, 0, e.episode, px, py), r);
        else
            env_step(c, f, i, (uint32_t)i, e, visited + (size_t)i * PLUME_VISIT_STRIDE, actions[i], noise[2 * i],
                     noise[2 * i + 1], true, pc, pt, r);
        pos_x[i] = e.px; pos_y[i] = e.py; step[i] = e.step;
        std::memcpy(obs + 6 * (size_t)i, r.obs, sizeof(r.obs));
        reward[i] = r.reward; done[i] = r.done; reached[i] = r.reached;
        info[5 * (size_t)i + 0] = r.conc_reward; info[5 * (size_t)i + 1] = r.explore_reward;
        info[5 * (size_t)i + 2] = r.move_penalty; info[5 * (size_t)i + 3] = r.tke_penalty;
        info[5 * (size_t)i + 4] = r.boundary_penalty;
    }
}

// procedural evaluation of single cells (float32 Box-Muller on the host differs from the device
// intrinsics in the last bits; the oracle comparison uses a tolerance)
void sim_cell_procedural(const plume_env_config* cfg, uint32_t gid, uint32_t episode, double sx, double sy,
                         const double* sin_tab, const double* cos_tab, const int32_t* x, const int32_t* y, int n,
                         double* conc, double* tke) {
    const Cfg c = make_cfg(*cfg);
    ProceduralField f{sin_tab, cos_tab};
    for (int i = 0; i < n; ++i) f.eval(c, 0, gid, episode, sx, sy, x[i], y[i], conc[i], tke[i]);
}

void sim_observe_f64(const plume_env_config* cfg, int n, const float* pos_x, const float* pos_y, const double* src_x,
                     const double* src_y, const int32_t* step, const uint16_t* visited, const double* conc_field,
                     const double* tke_field, float* obs) {
    const Cfg c = sim_cfg(cfg);
    MaterialisedField<double> f{conc_field, tke_field};
    for (int i = 0; i < n; ++i) {
        EnvRegs e{};
        e.px = pos_x[i]; e.py = pos_y[i]; e.sx = src_x[i]; e.sy = src_y[i]; e.step = step[i];
        observe(c, f, i, 0, e, visited + (size_t)i * PLUME_VISIT_STRIDE, obs + 6 * (size_t)i);
    }
}
}
