// env_kernels.cu -- batched plume environment: reset (P0), field generation (P1/K1),
// observation (P3) and the lockstep step (P2/K2).  Reference: PPOV2.1/environment.py.
//
// Layout: struct-of-arrays env state (one float/double/int array per attribute, index = env),
// so a warp stepping 32 consecutive envs issues fully coalesced 128 B / 256 B requests.
// Materialised fields are [N][G][G] with y fastest (the reference's field[x, y]).
//
// Algorithmic bytes (DESIGN.md "K1"/"K2"):
//   K1 generate : 2 fields x G*G x sizeof(T) written per env-reset (2.0 MB float, 4.0 MB double)
//   K2 step     : per env-step read pos 8 + src 16 + step 4 + episode 4 + radius 8 + bonus 8 +
//                 action 4 + visit 2 + carried cell (tke 8, conc 8, tag 4) = 74 B, write pos 8 + step 4 +
//                 visit 2 + obs 24 + reward 8 + done 1 + reached 1 + carried cell 20 = 68 B  => 142 B
//                 (+20 B info, +16 B injected noise, + field gathers x sizeof(T) in the materialised modes;
//                 source / radius / bonus / episode are written back only by a reset).  bench.py counts
//                 146 B = the procedural step with info and without the carried concentration's 16 B.
#include <cstdarg>

#include "common.cuh"

namespace plume {

std::string& last_error_ref() {
    static thread_local std::string err;
    return err;
}

int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return 1;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    }
    return cached;
}

// ---------------------------------------------------------------------------------------
// K1: field generation
// ---------------------------------------------------------------------------------------
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
};
template <>
struct Vec4<double> {
    static __device__ __forceinline__ void store(double* p, double a, double b, double c, double d) {
        reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
        reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
    }
};

__device__ __forceinline__ float exp2f_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// float instantiation: single-precision evaluation of env:53-63 (the float field is an
// approximation of the float64 reference field by construction)
__device__ __forceinline__ void cell_f32(const Cfg& c, float sx, float sy, int x, int y, float z, float u, float sinx,
                                         float cosy, float& conc, float& tke) {
    const float ddx = (float)x - sx, ddy = (float)y - sy;
    const float d2 = ddx * ddx + ddy * ddy;
    const float base = (float)c.conc_peak * expf(-d2 / (float)c.two_sigma_sq);
    tke = (float)c.ti * ((fabsf(z) + 0.3f * sinx * cosy) + 0.2f * u);
    conc = fminf(fmaxf(base + tke, 0.0f), (float)c.conc_peak);
}

template <typename T>
__global__ void __launch_bounds__(256) generate_fields_kernel(Cfg c, plume_env_state st, const int32_t* env_list,
                                                              int blocks_per_env, float* z_out, float* u_out) {
    const int li = blockIdx.x / blocks_per_env;
    const int quad = (blockIdx.x - li * blocks_per_env) * blockDim.x + threadIdx.x;
    const int cells = c.G * c.G;
    if (quad * 4 >= cells) return;
    const int env = env_list ? env_list[li] : li;
    const uint32_t gid = (uint32_t)(st.env_id_base + env);
    const uint32_t episode = (uint32_t)st.episode_idx[env];
    const double sx = st.src_x[env], sy = st.src_y[env];
    const int cell0 = quad * 4;
    const int x = cell0 / c.G, y = cell0 - x * c.G;    // G % 4 == 0: the 4 cells share the row

    float z[4], u[4];
    field_noise_quad(c, gid, episode, (uint32_t)quad, z, u);

    T conc[4], tke[4];
    const double sinx = st.sin_tab[x];
    const bool dispersion = c.plume_model == PLUME_MODEL_DISPERSION;
    Wind wind{1.0, 0.0, 1.0};
    if (dispersion) wind = wind_of(c, gid, episode);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (sizeof(T) == 8 || dispersion) {
            double cc, tt;
            plume_cell(c, sx, sy, x, y + k, (double)z[k], (double)u[k], sinx, st.cos_tab[y + k], cc, tt,
                       dispersion ? &wind : nullptr);
            conc[k] = (T)cc;
            tke[k] = (T)tt;
        } else {
            float cc, tt;
            cell_f32(c, (float)sx, (float)sy, x, y + k, z[k], u[k], (float)sinx, (float)st.cos_tab[y + k], cc, tt);
            conc[k] = cc;
            tke[k] = tt;
        }
    }
    const size_t off = (size_t)env * cells + cell0;
    Vec4<T>::store(reinterpret_cast<T*>(st.conc_field) + off, conc[0], conc[1], conc[2], conc[3]);
    Vec4<T>::store(reinterpret_cast<T*>(st.tke_field) + off, tke[0], tke[1], tke[2], tke[3]);
    if (z_out) {
        const size_t o2 = (size_t)li * cells + cell0;
        *reinterpret_cast<float4*>(z_out + o2) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(u_out + o2) = make_float4(u[0], u[1], u[2], u[3]);
    }
}

// ---- K1 fast path: float fields, one CTA per (row x, env), no integer divisions ---------------------------
// The kernel is bound by instruction issue, not by its 32 B per thread of stores (ncu r1: 34 % of the HBM
// roofline with the generic kernel), so everything that is not Philox is trimmed: constants pre-converted
// to float on the host, exp as ex2.approx of a pre-scaled argument, sqrt.approx in the Box-Muller radius
// (shared with every other consumer of the field stream through box_muller), rows mapped to blockIdx.x.
struct FieldF32Cfg {
    float peak, ti, exp_scale;     // exp_scale = -log2(e) / (2 sigma^2)
    FieldKeys keys;                // the field stream's Philox round keys (k + r W), bumped on the host
};

// float copies of 0.3*sin(0.05 x) and cos(0.07 y) (double -> float conversions run on the XU pipe, which this
// kernel saturates together with the MUFU ops: ncu/SASS r1: 32 XU-pipe instructions per 4 cells)
__device__ float g_wave_sin_f32[512];
__device__ float g_wave_cos_f32[512];

__global__ void wave_tables_f32_kernel(const double* __restrict__ sin_tab, const double* __restrict__ cos_tab, int G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G) {
        g_wave_sin_f32[i] = 0.3f * (float)sin_tab[i];
        g_wave_cos_f32[i] = (float)cos_tab[i];
    }
}

// One CTA per (kK1Rows rows, env); a thread owns EIGHT consecutive cells of a row (two Philox4x32-7 calls, two 16-byte
// stores per field): the per-thread setup (row terms, env state, addresses) is paid once per eight cells, and a CTA is
// large enough (512 threads) that the grid is not bound by CTA launches (one row per CTA: 512 000 CTAs of two warps
// for 1024 envs ran no faster with 20 % fewer instructions).
constexpr int kK1Rows = 8;
__global__ void __launch_bounds__(64 * kK1Rows) generate_fields_f32_kernel(Cfg c, FieldF32Cfg fc, plume_env_state st,
                                                                           const int32_t* env_list) {
    const int x = blockIdx.x * kK1Rows + threadIdx.y, li = blockIdx.y;
    const int y0 = 8 * threadIdx.x;
    if (y0 >= c.G || x >= c.G) return;
    const int env = env_list ? env_list[li] : li;
    const uint32_t gid = (uint32_t)(st.env_id_base + env);
    const uint32_t episode = (uint32_t)st.episode_idx[env];
    const float sx = (float)st.src_x[env], sy = (float)st.src_y[env];
    const int cell0 = x * c.G + y0;
    const float ddx = (float)x - sx;
    const float ddx2 = ddx * ddx;
    const float s3 = g_wave_sin_f32[x];
    const size_t off = (size_t)env * c.G * c.G + cell0;
    float* const conc_p = reinterpret_cast<float*>(st.conc_field) + off;
    float* const tke_p = reinterpret_cast<float*>(st.tke_field) + off;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (y0 + 4 * h >= c.G) break;                  // G % 8 == 4: the row's last thread owns one quad
        float z[4], u[4];
        field_noise_from_words(philox4x32_field((uint32_t)(cell0 >> 2) + h, episode, gid, kTagField, fc.keys), z, u);
        const float4 cosy4 = *reinterpret_cast<const float4*>(g_wave_cos_f32 + y0 + 4 * h);
        const float cosy[4] = {cosy4.x, cosy4.y, cosy4.z, cosy4.w};
        const float fy0 = (float)(y0 + 4 * h) - sy;       // one conversion, the other three by addition
        float conc[4], tke[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ddy = fy0 + (float)k;
            const float arg = fmaf(ddy, ddy, ddx2) * fc.exp_scale;
            // far from the source the Gaussian underflows: skip the ex2
            const float base = (arg > -126.0f) ? fc.peak * exp2f_approx(arg) : 0.0f;
            tke[k] = fc.ti * (fmaf(0.2f, u[k], fmaf(s3, cosy[k], fabsf(z[k]))));
            conc[k] = fminf(fmaxf(base + tke[k], 0.0f), fc.peak);
        }
        __stcs(reinterpret_cast<float4*>(conc_p + 4 * h), make_float4(conc[0], conc[1], conc[2], conc[3]));
        __stcs(reinterpret_cast<float4*>(tke_p + 4 * h), make_float4(tke[0], tke[1], tke[2], tke[3]));
    }
}

// dump-only variant (no field pointers needed): the draws of the listed envs
__global__ void __launch_bounds__(256) dump_noise_kernel(Cfg c, plume_env_state st, const int32_t* env_list,
                                                         int blocks_per_env, float* z_out, float* u_out) {
    const int li = blockIdx.x / blocks_per_env;
    const int quad = (blockIdx.x - li * blocks_per_env) * blockDim.x + threadIdx.x;
    const int cells = c.G * c.G;
    if (quad * 4 >= cells) return;
    const int env = env_list ? env_list[li] : li;
    const uint32_t gid = (uint32_t)(st.env_id_base + env);
    const uint32_t episode = (uint32_t)st.episode_idx[env];
    const int cell0 = quad * 4;
    float z[4], u[4];
    field_noise_quad(c, gid, episode, (uint32_t)quad, z, u);
    const size_t o2 = (size_t)li * cells + cell0;
    *reinterpret_cast<float4*>(z_out + o2) = make_float4(z[0], z[1], z[2], z[3]);
    *reinterpret_cast<float4*>(u_out + o2) = make_float4(u[0], u[1], u[2], u[3]);
}

// conc_field[x, y] / tke_field[x, y] of every env at one cell per env (the accessor the evaluators use,
// evaluate_with_lstm.py:67-68), for any field mode
template <typename Field>
__global__ void field_at_kernel(Cfg c, plume_env_state st, Field f, const int32_t* x, const int32_t* y, double* conc,
                                double* tke) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n_envs) return;
    const int cx = clip_cell(x[i], c.G), cy = clip_cell(y[i], c.G);
    double cc, tt;
    f.eval(c, i, (uint32_t)(st.env_id_base + i), (uint32_t)st.episode_idx[i], st.src_x[i], st.src_y[i], cx, cy, cc, tt);
    if (conc) conc[i] = cc;
    if (tke) tke[i] = tt;
}

__global__ void noise_at_kernel(Cfg c, plume_env_state st, const int32_t* env_local, const int32_t* x,
                                const int32_t* y, int n, float* z_out, float* u_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int env = env_local[i];
    float z, u;
    field_noise(c, (uint32_t)(st.env_id_base + env), (uint32_t)st.episode_idx[env], x[i], y[i], z, u);
    z_out[i] = z;
    u_out[i] = u;
}

// ---------------------------------------------------------------------------------------
// P0 reset
// ---------------------------------------------------------------------------------------
__global__ void reset_kernel(Cfg c, plume_env_state st, const int32_t* env_list, int n_list, const double* u_src) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    const int env = env_list ? env_list[li] : li;
    EnvRegs e = load_env(st, env);
    env_reset(c, (uint32_t)(st.env_id_base + env), e, st.visited + (size_t)env * PLUME_VISIT_STRIDE,
              u_src ? u_src + 2 * (size_t)li : nullptr, st.curriculum[0], st.curriculum[1]);
    store_env(st, env, e);
    if (st.cell_key) st.cell_key[env] = 0u;      // cached tke belongs to the previous episode
}

// ---------------------------------------------------------------------------------------
// P3 observe
// ---------------------------------------------------------------------------------------
template <typename Field>
__global__ void observe_kernel(Cfg c, plume_env_state st, Field f, float* obs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n_envs) return;
    const EnvRegs e = load_env(st, i);
    float o[6];
    observe(c, f, i, (uint32_t)(st.env_id_base + i), e, st.visited + (size_t)i * PLUME_VISIT_STRIDE, o);
#pragma unroll
    for (int k = 0; k < 6; ++k) obs[(size_t)i * 6 + k] = o[k];
}

// ---------------------------------------------------------------------------------------
// K2: lockstep step, one thread per env
// ---------------------------------------------------------------------------------------
// kSpec selects a specialisation at compile time (smaller code, fewer registers, more resident warps):
//   0  generic: plume model, PLUME_FLAG_FAST_REWARD and the division mode are read at run time
//   1  reference code model, exact float64 reward, constant divisions by reciprocal (make_cfg checked them)
//   2  as 1 with PLUME_FLAG_FAST_REWARD
// Per step the kernel evaluates ONE Philox field cell (the cell after the move); the tke of the cell before
// the move is carried in st.cell_tke (tagged, recomputed when stale), the concentration there is only needed
// inside the boundary band and is evaluated on demand.
#ifndef PLUME_STEP_MIN_BLOCKS
#define PLUME_STEP_MIN_BLOCKS 8
#endif
#ifndef PLUME_STEP_THREADS
#define PLUME_STEP_THREADS 128
#endif
template <typename Field>
struct FieldTraits {
    static constexpr bool kCacheTke = false;
};
template <>
struct FieldTraits<ProceduralField> {
    static constexpr bool kCacheTke = true;
};

template <typename Field, int kSpec>
__global__ void __launch_bounds__(PLUME_STEP_THREADS, kSpec != 0 ? PLUME_STEP_MIN_BLOCKS : 1) step_kernel(Cfg c_in, plume_env_state st, Field f, const int32_t* actions,
                                                   const double* step_noise_in, uint32_t flags, float* obs,
                                                   double* reward, uint8_t* done, uint8_t* reached, float* info,
                                                   float* final_obs, double* noise_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n_envs) return;
    Cfg c = c_in;
    if (kSpec != 0) {               // constant-propagated through the inlined per-env code
        c.plume_model = PLUME_MODEL_ISOTROPIC;
        c.fastdiv = 1;
    }
    const bool fast = kSpec == 0 ? (flags & PLUME_FLAG_FAST_REWARD) != 0 : kSpec == 2;
    const bool dispersion = c.plume_model == PLUME_MODEL_DISPERSION;
    const uint32_t gid = (uint32_t)(st.env_id_base + i);
    EnvRegs e = load_env(st, i);
    uint16_t* vis = st.visited + (size_t)i * PLUME_VISIT_STRIDE;
    const int action = actions[i];
    // every load whose address does not depend on data is issued here, ahead of the first use
    const bool cache = FieldTraits<Field>::kCacheTke && st.cell_tke && st.cell_conc && st.cell_key;
    const uint32_t cached_key = cache ? st.cell_key[i] : 0u;
    const double cached_tke = cache ? st.cell_tke[i] : 0.0;
    const double cached_conc = cache ? st.cell_conc[i] : 0.0;

    double z0, z1;
    if (step_noise_in) {
        const double2 z = reinterpret_cast<const double2*>(step_noise_in)[i];
        z0 = z.x;
        z1 = z.y;
    } else {
        step_noise(c, gid, e, z0, z1);
    }
    if (noise_out) reinterpret_cast<double2*>(noise_out)[i] = make_double2(z0, z1);

    int px, py;
    cell32_of(c, e, px, py);
    double ptke = cached_tke;
    const bool hit = cache && cached_key != 0u && cached_key == cell_key_of(c, px, py, e.episode);
    if (!hit) ptke = f.eval_tke(c, i, gid, e.episode, px, py);
    // the visit-table line this step will most likely touch (cell of position + move, before the turbulence
    // offset): requested now so that its DRAM latency overlaps the Philox / field arithmetic below
    {
        const int ms = (int)c.move_step;
        const int qx = clip_cell(px + (action == 3 ? ms : (action == 4 ? -ms : 0)), c.G) / c.cell_size;
        const int qy = clip_cell(py + (action == 1 ? ms : (action == 2 ? -ms : 0)), c.G) / c.cell_size;
#ifndef PLUME_STEP_NO_PREFETCH
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vis + qx * PLUME_MAX_GRID_DIVISIONS + qy));
#endif
    }

    StepResult r;
    if (fast) env_step_fast<false>(c, f, i, gid, e, vis, action, z0, z1, hit, (float)cached_conc, ptke, r);
    else env_step<false>(c, f, i, gid, e, vis, action, z0, z1, hit, cached_conc, ptke, r);

    reward[i] = r.reward;
    done[i] = r.done ? 1 : 0;
    reached[i] = r.reached ? 1 : 0;
    if (info) {
        const int n = st.n_envs;
        info[0 * (size_t)n + i] = r.conc_reward;
        info[1 * (size_t)n + i] = r.explore_reward;
        info[2 * (size_t)n + i] = (float)r.move_penalty;
        info[3 * (size_t)n + i] = r.tke_penalty;
        info[4 * (size_t)n + i] = (float)r.boundary_penalty;
    }
    bool was_reset = false;
    if ((flags & PLUME_FLAG_AUTO_RESET) && r.done) {
        if (final_obs) {
#pragma unroll
            for (int k = 0; k < 6; ++k) final_obs[(size_t)i * 6 + k] = r.obs[k];
        }
        env_reset(c, gid, e, vis, nullptr, st.curriculum[0], st.curriculum[1]);
        f.eval(c, i, gid, e.episode, e.sx, e.sy, 0, 0, r.cell_conc, r.cell_tke);     // agent_pos = (0,0), env:46
        make_obs(c, e, r.cell_conc, r.cell_tke, 0, r.obs, gid);
        was_reset = true;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) obs[(size_t)i * 6 + k] = r.obs[k];
    // state that a step changes: position, step counter (heading in the README model); the rest only on reset
    st.pos_x[i] = e.px;
    st.pos_y[i] = e.py;
    st.step_count[i] = e.step;
    if (dispersion && st.last_move) st.last_move[i] = (int8_t)e.last_move;
    if (was_reset) store_env(st, i, e);
    if (cache) {
        int ox, oy;
        cell32_of(c, e, ox, oy);
        st.cell_tke[i] = r.cell_tke;
        st.cell_conc[i] = r.cell_conc;
        st.cell_key[i] = cell_key_of(c, ox, oy, e.episode);
    }
}

}  // namespace plume

using namespace plume;

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static int check_cfg(const plume_env_config* cfg, const plume_env_state* st) {
    if (!cfg || !st) return fail("null config/state");
    if (cfg->grid_size <= 0 || cfg->grid_size % 4 != 0) return fail("grid_size must be a positive multiple of 4");
    if (cfg->grid_divisions <= 0 || cfg->grid_divisions > PLUME_MAX_GRID_DIVISIONS)
        return fail("grid_divisions must be in [1,%d]", PLUME_MAX_GRID_DIVISIONS);
    if (cfg->max_steps <= 0 || cfg->max_steps > 65535) return fail("max_steps must be in [1,65535] (uint16 visit counters)");
    if (st->n_envs < 0) return fail("negative n_envs");
    if (cfg->field_mode != PLUME_FIELD_PROCEDURAL && (!st->conc_field || !st->tke_field))
        return fail("materialised field mode needs conc_field/tke_field");
    if (!st->sin_tab || !st->cos_tab || !st->curriculum) return fail("sin_tab/cos_tab/curriculum missing");
    return 0;
}

extern "C" int plume_abi_version(void) { return PLUME_B200_ABI_VERSION; }

extern "C" const char* plume_last_error(void) { return last_error_ref().c_str(); }

extern "C" int plume_device_info(int32_t* sms, int32_t* major, int32_t* minor) {
    int dev = 0;
    PLUME_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    PLUME_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return 0;
}

extern "C" int plume_env_reset(const plume_env_config* cfg, const plume_env_state* st, const int32_t* env_list,
                               int32_t n_list, const double* u_src, void* stream) {
    if (check_cfg(cfg, st)) return 1;
    if (!env_list) n_list = st->n_envs;
    if (n_list <= 0) return 0;
    const Cfg c = make_cfg(*cfg, *st);
    reset_kernel<<<(n_list + 127) / 128, 128, 0, as_stream(stream)>>>(c, *st, env_list, n_list, u_src);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_generate_fields(const plume_env_config* cfg, const plume_env_state* st, const int32_t* env_list,
                                     int32_t n_list, float* z_out, float* u_out, void* stream) {
    if (!cfg || !st) return fail("null config/state");
    if (cfg->grid_size <= 0 || cfg->grid_size % 4 != 0) return fail("grid_size must be a positive multiple of 4");
    if (!env_list) n_list = st->n_envs;
    if (n_list <= 0) return 0;
    PLUME_CHECK_ARG((z_out == nullptr) == (u_out == nullptr), "z_out and u_out must be given together");
    const Cfg c = make_cfg(*cfg, *st);
    const int quads = c.G * c.G / 4;
    const int bpe = (quads + 255) / 256;
    const long long blocks = (long long)bpe * n_list;
    PLUME_CHECK_ARG(blocks < 2147483647LL, "too many envs for one launch");
    cudaStream_t s = as_stream(stream);
    if (cfg->field_mode == PLUME_FIELD_F32) {
        PLUME_CHECK_ARG(st->conc_field && st->tke_field, "field pointers missing");
        if (!z_out && !u_out && c.G % 4 == 0 && c.G <= 512 && n_list <= 65535 &&
            c.plume_model == PLUME_MODEL_ISOTROPIC) {
            FieldF32Cfg fc;
            fc.peak = (float)c.conc_peak;
            fc.ti = (float)c.ti;
            fc.exp_scale = (float)(-1.4426950408889634 / c.two_sigma_sq);
            fc.keys = make_field_keys(c.k0, c.k1);
            wave_tables_f32_kernel<<<(c.G + 255) / 256, 256, 0, s>>>(st->sin_tab, st->cos_tab, c.G);
            generate_fields_f32_kernel<<<dim3((unsigned)((c.G + kK1Rows - 1) / kK1Rows), (unsigned)n_list), dim3(64, kK1Rows), 0,
                                         s>>>(c, fc, *st, env_list);
        } else {
            generate_fields_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(c, *st, env_list, bpe, z_out, u_out);
        }
    } else if (cfg->field_mode == PLUME_FIELD_F64) {
        PLUME_CHECK_ARG(st->conc_field && st->tke_field, "field pointers missing");
        generate_fields_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(c, *st, env_list, bpe, z_out, u_out);
    } else {
        PLUME_CHECK_ARG(z_out != nullptr, "procedural mode: only the noise dump is available");
        dump_noise_kernel<<<(unsigned)blocks, 256, 0, s>>>(c, *st, env_list, bpe, z_out, u_out);
    }
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_field_at(const plume_env_config* cfg, const plume_env_state* st, const int32_t* x, const int32_t* y,
                              double* conc, double* tke, void* stream) {
    if (check_cfg(cfg, st)) return 1;
    PLUME_CHECK_ARG(x && y && (conc || tke), "null pointer");
    if (st->n_envs == 0) return 0;
    const Cfg c = make_cfg(*cfg, *st);
    const int blocks = (st->n_envs + 127) / 128;
    cudaStream_t s = as_stream(stream);
    if (cfg->field_mode == PLUME_FIELD_PROCEDURAL)
        field_at_kernel<<<blocks, 128, 0, s>>>(c, *st, ProceduralField{st->sin_tab, st->cos_tab}, x, y, conc, tke);
    else if (cfg->field_mode == PLUME_FIELD_F32)
        field_at_kernel<<<blocks, 128, 0, s>>>(
            c, *st, MaterialisedField<float>{(const float*)st->conc_field, (const float*)st->tke_field}, x, y, conc, tke);
    else
        field_at_kernel<<<blocks, 128, 0, s>>>(
            c, *st, MaterialisedField<double>{(const double*)st->conc_field, (const double*)st->tke_field}, x, y, conc,
            tke);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_field_noise_at(const plume_env_config* cfg, const plume_env_state* st, const int32_t* env_local,
                                    const int32_t* x, const int32_t* y, int32_t n, float* z_out, float* u_out,
                                    void* stream) {
    if (!cfg || !st) return fail("null config/state");
    if (n <= 0) return 0;
    const Cfg c = make_cfg(*cfg, *st);
    noise_at_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(c, *st, env_local, x, y, n, z_out, u_out);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_env_observe(const plume_env_config* cfg, const plume_env_state* st, float* obs, void* stream) {
    if (check_cfg(cfg, st)) return 1;
    if (st->n_envs == 0) return 0;
    const Cfg c = make_cfg(*cfg, *st);
    const int blocks = (st->n_envs + 127) / 128;
    cudaStream_t s = as_stream(stream);
    if (cfg->field_mode == PLUME_FIELD_PROCEDURAL) {
        observe_kernel<<<blocks, 128, 0, s>>>(c, *st, ProceduralField{st->sin_tab, st->cos_tab}, obs);
    } else if (cfg->field_mode == PLUME_FIELD_F32) {
        observe_kernel<<<blocks, 128, 0, s>>>(
            c, *st, MaterialisedField<float>{(const float*)st->conc_field, (const float*)st->tke_field}, obs);
    } else {
        observe_kernel<<<blocks, 128, 0, s>>>(
            c, *st, MaterialisedField<double>{(const double*)st->conc_field, (const double*)st->tke_field}, obs);
    }
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_env_step(const plume_env_config* cfg, const plume_env_state* st, const int32_t* actions,
                              const double* step_noise_in, uint32_t flags, float* obs, double* reward, uint8_t* done,
                              uint8_t* reached, float* info, float* final_obs, double* noise_out, void* stream) {
    if (check_cfg(cfg, st)) return 1;
    PLUME_CHECK_ARG(actions && obs && reward && done && reached, "null output/input pointer");
    if (st->n_envs == 0) return 0;
    if ((flags & PLUME_FLAG_AUTO_RESET) && cfg->field_mode != PLUME_FIELD_PROCEDURAL)
        return fail("auto-reset inside the step needs the procedural field mode "
                    "(materialised fields are regenerated by plume_generate_fields)");
    const Cfg c = make_cfg(*cfg, *st);
    const int blocks = (st->n_envs + PLUME_STEP_THREADS - 1) / PLUME_STEP_THREADS;
    cudaStream_t s = as_stream(stream);
#define PLUME_STEP_LAUNCH(FIELD, SPEC, ...)                                                                         \
    step_kernel<FIELD, SPEC><<<blocks, PLUME_STEP_THREADS, 0, s>>>(c, *st, FIELD{__VA_ARGS__}, actions, step_noise_in, flags, obs, \
                                                    reward, done, reached, info, final_obs, noise_out)
    if (cfg->field_mode == PLUME_FIELD_PROCEDURAL) {
        const bool spec = c.plume_model == PLUME_MODEL_ISOTROPIC && c.fastdiv;
        if (spec && !(flags & PLUME_FLAG_FAST_REWARD)) PLUME_STEP_LAUNCH(ProceduralField, 1, st->sin_tab, st->cos_tab);
        else if (spec) PLUME_STEP_LAUNCH(ProceduralField, 2, st->sin_tab, st->cos_tab);
        else PLUME_STEP_LAUNCH(ProceduralField, 0, st->sin_tab, st->cos_tab);
    } else if (cfg->field_mode == PLUME_FIELD_F32) {
        PLUME_STEP_LAUNCH(MaterialisedField<float>, 0, (const float*)st->conc_field, (const float*)st->tke_field);
    } else {
        PLUME_STEP_LAUNCH(MaterialisedField<double>, 0, (const double*)st->conc_field, (const double*)st->tke_field);
    }
#undef PLUME_STEP_LAUNCH
    PLUME_LAUNCH_CHECK();
    return 0;
}
