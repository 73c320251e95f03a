// plume_core.h -- per-environment arithmetic shared by every kernel of libplume_b200.
//
// Everything here is a __host__ __device__ inline function so that the exact device logic
// can also be compiled with g++ into the host-side logic tests (tests/host_sim); the
// product only ever runs the device instantiation.
//
// Reference: /root/reference/PPOV2.1/environment.py (cited as env:LINE below).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/plume_b200.h"

#if defined(__CUDACC__)
#define PLUME_HD __host__ __device__ __forceinline__
#else
#define PLUME_HD inline
#endif

namespace plume {

// ---------------------------------------------------------------------------------------
// IEEE double arithmetic that the compiler may not contract into FMAs: the position
// update, the cell truncation and the reached test must be bit-identical to numpy's
// float64 scalar arithmetic (env:107-117,155-156).
// ---------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PLUME_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
PLUME_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
PLUME_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
PLUME_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
PLUME_HD double dsqrt(double a) { return __dsqrt_rn(a); }
PLUME_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
PLUME_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
PLUME_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
PLUME_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
// host build: compiled with -ffp-contract=off
PLUME_HD double dadd(double a, double b) { return a + b; }
PLUME_HD double dsub(double a, double b) { return a - b; }
PLUME_HD double dmul(double a, double b) { return a * b; }
PLUME_HD double ddiv(double a, double b) { return a / b; }
PLUME_HD double dsqrt(double a) { return sqrt(a); }
PLUME_HD float fadd(float a, float b) { return a + b; }
PLUME_HD float fsub(float a, float b) { return a - b; }
PLUME_HD float fmul(float a, float b) { return a * b; }
PLUME_HD float fdiv(float a, float b) { return a / b; }
#endif

// Correctly rounded a / b for a divisor b that is a configuration constant, y = RN(1 / b) precomputed on the
// host: q0 = RN(a y) is within one ulp of a / b, the residual r = a - b q0 is exact in one FMA, and
// RN(q0 + r y) is the correctly rounded quotient (Markstein's theorem; b must not have an all-ones mantissa).
// Three FMA-pipe instructions instead of the ~27 of the generic IEEE division sequence.  Differs from a / b
// only in the sign of a zero result.  make_cfg() enables it per divisor after checking |b y - 1| <= 2^-54;
// tests/test_host_sim.py compares it with a / b on adversarial numerators.
#if defined(__CUDA_ARCH__)
PLUME_HD double ddiv_const(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    return __fma_rn(__fma_rn(-b, q0, a), y, q0);
}
#else
PLUME_HD double ddiv_const(double a, double b, double y) {
    const double q0 = a * y;
    return fma(fma(-b, q0, a), y, q0);
}
#endif

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Stream layout documented in oracle/philox.py.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kTagSrc = 1, kTagField = 2, kTagStep = 3, kTagAct = 4, kTagWind = 5;

struct U4 {
    uint32_t x, y, z, w;
};

template <int kRounds>
PLUME_HD U4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return U4{c0, c1, c2, c3};
}
PLUME_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}

PLUME_HD float uniform24(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }           // [0,1)
PLUME_HD float uniform24_open0(uint32_t r) { return ((float)(r >> 8) + 1.0f) * 5.9604644775390625e-08f; }  // (0,1]
PLUME_HD double uniform53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) / 9007199254740992.0;
}

// Box-Muller in float32; the same function generates the materialised field and the
// procedural lookups, so both see identical draws.
PLUME_HD void box_muller(uint32_t r0, uint32_t r1, float& z0, float& z1) {
    const float u1 = uniform24_open0(r0);
    const float u2 = uniform24(r1);
#if defined(__CUDA_ARCH__)
    // sqrt.approx (MUFU.SQRT, <= 1 ulp) instead of the IEEE sqrtf sequence: K1 is bound by instruction
    // issue; every consumer of the field stream goes through this function, so they all see the same draws
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-2.0f * __logf(u1)));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
#else
    const float rad = sqrtf(-2.0f * logf(u1));
    const float s = sinf(6.283185307179586f * u2), c = cosf(6.283185307179586f * u2);
#endif
    z0 = rad * c;
    z1 = rad * s;
}

// ---------------------------------------------------------------------------------------
// device-side view of the configuration
// ---------------------------------------------------------------------------------------
struct Cfg {
    int32_t G, max_steps, divisions, cell_size, field_mode, plume_model;
    double conc_peak, ti, two_sigma_sq, clip_hi, move_step;
    double conc_coef, tke_factor, bnd_penalty, bnd_start, initial_radius;
    uint32_t k0, k1;
    float peak_f, inv_peak_f, exp2_scale, inv_nine_f, inv_G_f;      // float constants of the fast reward path
    // exact division by the configuration constants (ddiv_const): nine = TI*3, conc_peak, G (float)
    int32_t fastdiv;                 // 1 = the reciprocals below passed the host check
    double nine, inv_nine, inv_peak;
    double G_d, inv_G_d;             // float / G goes through the double quotient (see div_G)
    double ms_eps, inv_ms_eps, inv_eps;   // move_step + 1e-6 and the reciprocals of the two conc_gradient divisors
    // optional host tables (NULL = evaluate): (float)(step / max_steps) and (float)(count**0.75 + 1)
    const float* step_frac_tab;
    const float* visit_denom_tab;
};

inline bool reciprocal_ok(double b) {
    if (!(b > 0.0) || !isfinite(b)) return false;
    const double y = 1.0 / b;
    return fabs(fma(b, y, -1.0)) <= 5.551115123125783e-17;      // 2^-54
}


inline Cfg make_cfg(const plume_env_config& c) {
    Cfg o;
    o.G = c.grid_size;
    o.max_steps = c.max_steps;
    o.divisions = c.grid_divisions;
    o.cell_size = c.grid_size / c.grid_divisions;          // env:37
    o.field_mode = c.field_mode;
    o.plume_model = c.plume_model;
    o.conc_peak = c.conc_peak;
    o.ti = c.turbulence_intensity;
    o.two_sigma_sq = 2 * (c.sigma * c.sigma);              // env:56  2*(GAUSSIAN_RADIUS)**2
    o.clip_hi = c.clip_hi;
    o.move_step = c.grid_size * 0.05;                      // env:98
    o.conc_coef = c.conc_reward_coef;
    o.tke_factor = c.tke_penalty_factor;
    o.bnd_penalty = c.boundary_penalty;
    o.bnd_start = c.boundary_decay_start;
    o.initial_radius = c.initial_radius;
    o.peak_f = (float)c.conc_peak;
    o.inv_peak_f = (float)(1.0 / c.conc_peak);
    o.exp2_scale = (float)(-1.4426950408889634 / o.two_sigma_sq);
    o.inv_nine_f = (float)(1.0 / (c.turbulence_intensity * 3.0));
    o.inv_G_f = (float)(1.0 / (double)c.grid_size);
    o.k0 = (uint32_t)(c.seed & 0xFFFFFFFFu);
    o.k1 = (uint32_t)(c.seed >> 32);
    o.nine = c.turbulence_intensity * 3.0;                 // env:84,107
    o.inv_nine = 1.0 / o.nine;
    o.inv_peak = 1.0 / c.conc_peak;
    o.G_d = (double)c.grid_size;
    o.inv_G_d = 1.0 / o.G_d;
    o.ms_eps = o.move_step + 1e-6;                         // env:119
    o.inv_ms_eps = 1.0 / o.ms_eps;
    o.inv_eps = 1.0 / 1e-6;
    o.fastdiv = (reciprocal_ok(o.nine) && reciprocal_ok(c.conc_peak) && reciprocal_ok(o.G_d) &&
                 reciprocal_ok(o.ms_eps) && reciprocal_ok(1e-6)) ? 1 : 0;
    o.step_frac_tab = nullptr;
    o.visit_denom_tab = nullptr;
    return o;
}

inline Cfg make_cfg(const plume_env_config& c, const plume_env_state& st) {
    Cfg o = make_cfg(c);
    o.step_frac_tab = st.step_frac_tab;
    o.visit_denom_tab = st.visit_denom_tab;
    return o;
}

// tag of a cached per-cell value: valid bit | 13 bits of the episode | linear cell index (G <= 512); 0 = none
PLUME_HD uint32_t cell_key_of(const Cfg& c, int x, int y, uint32_t episode) {
    if (c.G > 512) return 0u;
    return 0x80000000u | ((episode & 0x1FFFu) << 18) | ((uint32_t)x * (uint32_t)c.G + (uint32_t)y);
}

// a / nine, a / conc_peak, a / G(float) -- exact either way, three instructions when the reciprocals qualify
PLUME_HD double div_nine(const Cfg& c, double a) { return c.fastdiv ? ddiv_const(a, c.nine, c.inv_nine) : ddiv(a, c.nine); }
PLUME_HD double div_peak(const Cfg& c, double a) {
    return c.fastdiv ? ddiv_const(a, c.conc_peak, c.inv_peak) : ddiv(a, c.conc_peak);
}
// float32 a / float32 G (env:81): the float64 quotient of two float32 values rounded to float32 equals the
// float32 quotient (double rounding is innocuous for division when the wide format has >= 2p+2 bits)
PLUME_HD float div_G(const Cfg& c, float a) {
    return c.fastdiv ? (float)ddiv_const((double)a, c.G_d, c.inv_G_d) : fdiv(a, (float)c.G);
}

// ---------------------------------------------------------------------------------------
// P1: one cell of the plume, env:53-63.  `z`,`u` are the cell's randn/rand draws.
// ---------------------------------------------------------------------------------------
// README plume (README.md:50,97): wind of (env, episode) -- direction uniform on the circle, speed in [1,5) m/s
struct Wind {
    double c, s, speed;
};
constexpr double kWindMaxSpeed = 5.0;
constexpr double kDispersionRef = 10.0;     // centre line saturates at the peak within 10 px of the source

PLUME_HD Wind wind_of(const Cfg& c, uint32_t env_gid, uint32_t episode) {
    const U4 r = philox4x32_10(0u, episode, env_gid, kTagWind, c.k0, c.k1);
    const double phi = 6.283185307179586 * (((double)(r.x >> 5) * 67108864.0 + (double)(r.y >> 6)) / 9007199254740992.0);
    const double u = ((double)(r.z >> 5) * 67108864.0 + (double)(r.w >> 6)) / 9007199254740992.0;
    return Wind{cos(phi), sin(phi), 1.0 + (kWindMaxSpeed - 1.0) * u};
}

// Gaussian dispersion in the wind-rotated frame: sigma_y = 0.3 xd^0.71 (README.md:50), amplitude
// peak * min(1, sigma_y(10) / sigma_y(xd)), zero upwind of the source.
PLUME_HD double dispersion_base(const Cfg& c, double ddx, double ddy, const Wind& w) {
    const double xd = ddx * w.c + ddy * w.s;          // downwind distance
    if (!(xd > 0.0)) return 0.0;
    const double yc = ddy * w.c - ddx * w.s;          // crosswind offset
    const double sig = 0.3 * pow(xd, 0.71);
    const double sig0 = 0.3 * pow(kDispersionRef, 0.71);
    const double amp = sig0 < sig ? sig0 / sig : 1.0;
    return c.conc_peak * amp * exp(-(yc * yc) / (2.0 * sig * sig));
}

PLUME_HD void plume_cell(const Cfg& c, double sx, double sy, int x, int y, double z, double u, double sinx,
                         double cosy, double& conc, double& tke, const Wind* wind = nullptr) {
    const double ddx = dsub((double)x, sx), ddy = dsub((double)y, sy);
    double base;
    if (wind) {
        base = dispersion_base(c, ddx, ddy, *wind);
    } else {
        const double dist = dsqrt(dadd(dmul(ddx, ddx), dmul(ddy, ddy)));      // env:54
        base = dmul(c.conc_peak, exp(ddiv(-dmul(dist, dist), c.two_sigma_sq)));   // env:56
    }
    const double wave = dmul(dmul(0.3, sinx), cosy);                          // env:59
    tke = dmul(c.ti, dadd(dadd(fabs(z), wave), dmul(0.2, u)));                // env:57-61,63
    const double v = dadd(base, tke);
    conc = v < 0.0 ? 0.0 : (v > c.conc_peak ? c.conc_peak : v);               // env:62
}

// ---- the field stream: ONE Philox4x32-7 call per FOUR consecutive cells ------------------------------------------
// (the plume of an episode is 2 x 250 000 draws per env; K1 was bound by instruction issue with 40 % of its
// instructions in two Philox4x32-10 calls per four cells -- profiles/r1d_k1_ncu_summary.txt.  Seven rounds are the
// reduced-round variant Random123 documents as Crush-resistant (Salmon et al. SC'11, Table 2); the other streams keep
// ten.)  Counter (cell >> 2, episode, env, TAG_FIELD); the 128 bits of a call are cut into
//   word 0: [31..12] radius uniform of cells 4q, 4q+1 (20 bits, (m + 1) 2^-20 in (0, 1])   [11..0] u of cell 4q
//   word 1: [31..12] radius uniform of cells 4q+2, 4q+3                                    [11..0] u of cell 4q+1
//   word 2: [15..0]  angle uniform of the first pair (16 bits, m 2^-16)   [31..16] angle uniform of the second pair
//   word 3: [11..0]  u of cell 4q+2   [23..12] u of cell 4q+3   (u = m 2^-12 in [0, 1))
// Box-Muller per pair: z_even = r cos(2 pi a), z_odd = r sin(2 pi a), r = sqrt(-2 ln(radius uniform)).
// The conversions are exact in float32 (the bit patterns below equal m * 2^-b), so oracle/philox.py states them
// arithmetically.
constexpr int kFieldRounds = 7;
PLUME_HD float field_unit_bits(uint32_t m, int bits) {      // m * 2^-bits for m < 2^bits <= 2^23, without int->float
#if defined(__CUDA_ARCH__)
    return __uint_as_float(0x3F800000u | (m << (23 - bits))) - 1.0f;
#else
    return (float)m * (1.0f / (float)(1u << bits));
#endif
}
PLUME_HD void box_muller_pair(uint32_t radius20, uint32_t angle16, float& z0, float& z1) {
    const float u1 = field_unit_bits(radius20, 20) + 9.5367431640625e-07f;      // (m + 1) 2^-20, exact
    const float u2 = field_unit_bits(angle16, 16);
#if defined(__CUDA_ARCH__)
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-2.0f * __logf(u1)));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
#else
    const float rad = sqrtf(-2.0f * logf(u1));
    const float s = sinf(6.283185307179586f * u2), c = cosf(6.283185307179586f * u2);
#endif
    z0 = rad * c;
    z1 = rad * s;
}
// Philox with the round keys k + r * W precomputed (K1: the key bumps are 2 of the ~10 instructions of a round)
struct FieldKeys {
    uint32_t k[2 * kFieldRounds];
};
PLUME_HD FieldKeys make_field_keys(uint32_t k0, uint32_t k1) {
    FieldKeys f;
    for (int r = 0; r < kFieldRounds; ++r) {
        f.k[2 * r] = k0 + (uint32_t)r * kPhiloxW0;
        f.k[2 * r + 1] = k1 + (uint32_t)r * kPhiloxW1;
    }
    return f;
}
PLUME_HD U4 philox4x32_field(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const FieldKeys& fk) {
#pragma unroll
    for (int r = 0; r < kFieldRounds; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ fk.k[2 * r];
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ fk.k[2 * r + 1];
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
    return U4{c0, c1, c2, c3};
}
PLUME_HD void field_noise_from_words(const U4& r, float* z, float* u) {
    box_muller_pair(r.x >> 12, r.z & 0xFFFFu, z[0], z[1]);
    box_muller_pair(r.y >> 12, r.z >> 16, z[2], z[3]);
    u[0] = field_unit_bits(r.x & 0xFFFu, 12);
    u[1] = field_unit_bits(r.y & 0xFFFu, 12);
    u[2] = field_unit_bits(r.w & 0xFFFu, 12);
    u[3] = field_unit_bits((r.w >> 12) & 0xFFFu, 12);
}
// the draws of the four cells 4q .. 4q+3 (cell = x*G + y) of (env, episode)
PLUME_HD void field_noise_quad(const Cfg& c, uint32_t env_gid, uint32_t episode, uint32_t quad, float* z, float* u) {
    const U4 r = philox4x32<kFieldRounds>(quad, episode, env_gid, kTagField, c.k0, c.k1);
    box_muller_pair(r.x >> 12, r.z & 0xFFFFu, z[0], z[1]);
    box_muller_pair(r.y >> 12, r.z >> 16, z[2], z[3]);
    u[0] = field_unit_bits(r.x & 0xFFFu, 12);
    u[1] = field_unit_bits(r.y & 0xFFFu, 12);
    u[2] = field_unit_bits(r.w & 0xFFFu, 12);
    u[3] = field_unit_bits((r.w >> 12) & 0xFFFu, 12);
}
// the two draws of cell (x,y) of (env, episode): only this cell's pair is evaluated
PLUME_HD void field_noise(const Cfg& c, uint32_t env_gid, uint32_t episode, int x, int y, float& z, float& u) {
    const uint32_t cell = (uint32_t)x * (uint32_t)c.G + (uint32_t)y;
    const U4 r = philox4x32<kFieldRounds>(cell >> 2, episode, env_gid, kTagField, c.k0, c.k1);
    const uint32_t k = cell & 3u;
    float z0, z1;
    if (k < 2u) box_muller_pair(r.x >> 12, r.z & 0xFFFFu, z0, z1);
    else box_muller_pair(r.y >> 12, r.z >> 16, z0, z1);
    z = (k & 1u) ? z1 : z0;
    const uint32_t m = k == 0u ? r.x : (k == 1u ? r.y : (k == 2u ? r.w : r.w >> 12));
    u = field_unit_bits(m & 0xFFFu, 12);
}

// Field access policies -------------------------------------------------------------------
struct ProceduralField {
    const double* sin_tab;
    const double* cos_tab;
    PLUME_HD void eval(const Cfg& c, int env_local, uint32_t env_gid, uint32_t episode, double sx, double sy, int x,
                       int y, double& conc, double& tke) const {
        (void)env_local;
        float z, u;
        field_noise(c, env_gid, episode, x, y, z, u);
        if (c.plume_model == PLUME_MODEL_DISPERSION) {
            const Wind w = wind_of(c, env_gid, episode);
            plume_cell(c, sx, sy, x, y, (double)z, (double)u, sin_tab[x], cos_tab[y], conc, tke, &w);
        } else {
            plume_cell(c, sx, sy, x, y, (double)z, (double)u, sin_tab[x], cos_tab[y], conc, tke);
        }
    }
    // tke alone (env:57-61,63): the position update of the next step needs only this, not the Gaussian
    PLUME_HD double eval_tke(const Cfg& c, int env_local, uint32_t env_gid, uint32_t episode, int x, int y) const {
        (void)env_local;
        float z, u;
        field_noise(c, env_gid, episode, x, y, z, u);
        const double wave = dmul(dmul(0.3, sin_tab[x]), cos_tab[y]);
        return dmul(c.ti, dadd(dadd(fabs((double)z), wave), dmul(0.2, (double)u)));
    }
    // PLUME_FLAG_FAST_REWARD: tke exactly as above (it drives the float64 position update, i.e. the flags),
    // the concentration in float32 (it only feeds the observation and the reward: fp32 rel 1e-5 bar)
    PLUME_HD void eval_fast(const Cfg& c, int env_local, uint32_t env_gid, uint32_t episode, double sx, double sy, int x,
                            int y, float& conc, double& tke) const {
        if (c.plume_model == PLUME_MODEL_DISPERSION) {
            double cc;
            eval(c, env_local, env_gid, episode, sx, sy, x, y, cc, tke);
            conc = (float)cc;
            return;
        }
        float z, u;
        field_noise(c, env_gid, episode, x, y, z, u);
        const double wave = dmul(dmul(0.3, sin_tab[x]), cos_tab[y]);
        tke = dmul(c.ti, dadd(dadd(fabs((double)z), wave), dmul(0.2, (double)u)));      // env:57-61,63 (exact)
        const float ddx = (float)x - (float)sx, ddy = (float)y - (float)sy;
        const float arg = (ddx * ddx + ddy * ddy) * c.exp2_scale;
#if defined(__CUDA_ARCH__)
        float base;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(base) : "f"(arg));
#else
        const float base = exp2f(arg);
#endif
        const float v = c.peak_f * base + (float)tke;
        conc = v < 0.0f ? 0.0f : (v > c.peak_f ? c.peak_f : v);                          // env:62
    }
};

template <typename T>
struct MaterialisedField {
    const T* conc_field;
    const T* tke_field;
    PLUME_HD void eval(const Cfg& c, int env_local, uint32_t, uint32_t, double, double, int x, int y, double& conc,
                       double& tke) const {
        const size_t off = ((size_t)env_local * c.G + x) * c.G + y;      // field[x, y], x = first axis
        conc = (double)conc_field[off];
        tke = (double)tke_field[off];
    }
    PLUME_HD void eval_fast(const Cfg& c, int env_local, uint32_t, uint32_t, double, double, int x, int y, float& conc,
                            double& tke) const {
        const size_t off = ((size_t)env_local * c.G + x) * c.G + y;
        conc = (float)conc_field[off];
        tke = (double)tke_field[off];
    }
    PLUME_HD double eval_tke(const Cfg& c, int env_local, uint32_t, uint32_t, int x, int y) const {
        return (double)tke_field[((size_t)env_local * c.G + x) * c.G + y];
    }
};

// ---------------------------------------------------------------------------------------
// per-env registers
// ---------------------------------------------------------------------------------------
struct EnvRegs {
    float px, py;        // agent_pos (float32)
    double sx, sy;       // source_pos
    int32_t step;        // step_count
    uint32_t episode;    // episode_idx
    double radius;       // current_radius (latched)
    double ebonus;       // explore_bonus (latched)
    int32_t last_move;   // last non-zero action (README reward: heading change), 0 = none yet
};

struct StepResult {
    float obs[6];
    double reward;
    bool done, reached;
    float conc_reward, explore_reward, tke_penalty;
    double move_penalty, boundary_penalty;
    double cur_conc;       // conc at the float64 cell / peak (env:118); evaluated only where the reward needs it
    double cell_conc, cell_tke;   // field at the float32 cell after the move (what obs[2], obs[3] hold)
    double distance;       // ||agent_pos - source_pos|| after the move (env:155)
};

PLUME_HD int clip_cell(int v, int G) { return v < 0 ? 0 : (v > G - 1 ? G - 1 : v); }

// vc**0.75 + 1 (env:140): vc^3 is exact in double, two correctly rounded sqrt's.
PLUME_HD double visit_denominator(int vc) {
    const double v = (double)vc;
    return dadd(dsqrt(dsqrt(dmul(dmul(v, v), v))), 1.0);
}

// (float)(vc**0.75 + 1) and (float)(step / max_steps): host tables when the caller provides them.  The tables hold
// max_steps + 2 and max_steps + 1 entries; an env stepped past `done` (the reference allows it: evaluators run a
// fixed number of steps, environment.py has no guard) leaves them and takes the arithmetic form.
PLUME_HD float visit_denominator_f(const Cfg& c, int vc) {
    if (c.visit_denom_tab && vc <= c.max_steps + 1) {
#if defined(__CUDA_ARCH__)
        return __ldg(c.visit_denom_tab + vc);
#else
        return c.visit_denom_tab[vc];
#endif
    }
    return (float)visit_denominator(vc);
}
PLUME_HD float step_fraction_f(const Cfg& c, int step) {
    if (c.step_frac_tab && step <= c.max_steps) {
#if defined(__CUDA_ARCH__)
        return __ldg(c.step_frac_tab + step);
#else
        return c.step_frac_tab[step];
#endif
    }
    return (float)ddiv((double)step, (double)c.max_steps);
}

// P3 _get_obs, env:71-87, given the field values at the float32 cell.
PLUME_HD void make_obs(const Cfg& c, const EnvRegs& e, double cell_conc, double cell_tke, int visit_here,
                       float* obs, uint32_t env_gid = 0) {
    obs[0] = div_G(c, e.px);                                                  // env:81
    obs[1] = div_G(c, e.py);
    obs[2] = (float)div_peak(c, cell_conc);                                   // env:83
    obs[3] = (float)div_nine(c, cell_tke);                                    // env:84
    obs[4] = step_fraction_f(c, e.step);                                      // env:85
    // env:78  min(visit / 5.0, 1.0): six possible values, folded at compile time (no float64 division)
    obs[5] = visit_here >= 5 ? 1.0f
             : (visit_here == 4 ? (float)(4.0 / 5.0)
                : (visit_here == 3 ? (float)(3.0 / 5.0)
                   : (visit_here == 2 ? (float)(2.0 / 5.0) : (visit_here == 1 ? (float)(1.0 / 5.0) : 0.0f))));
    if (c.plume_model == PLUME_MODEL_DISPERSION) {      // README state: [CH4], wind vector, UAV position
        const Wind w = wind_of(c, env_gid, e.episode);
        obs[3] = (float)(w.c * w.speed / kWindMaxSpeed);
        obs[5] = (float)(w.s * w.speed / kWindMaxSpeed);
    }
}

PLUME_HD void cell32_of(const Cfg& c, const EnvRegs& e, int& x, int& y) {
    x = clip_cell((int)e.px, c.G);                                            // env:72-73
    y = clip_cell((int)e.py, c.G);
}

template <typename Field>
PLUME_HD void observe(const Cfg& c, const Field& f, int env_local, uint32_t env_gid, const EnvRegs& e,
                      const uint16_t* visited, float* obs) {
    int x, y;
    cell32_of(c, e, x, y);
    double conc, tke;
    f.eval(c, env_local, env_gid, e.episode, e.sx, e.sy, x, y, conc, tke);
    const int vis = visited[(x / c.cell_size) * PLUME_MAX_GRID_DIVISIONS + (y / c.cell_size)];
    make_obs(c, e, conc, tke, vis, obs, env_gid);
}

// reached = ||agent_pos - source_pos|| <= radius (env:155-156) without the square root when the squared
// distance is clearly on one side of radius^2 (relative margin 1e-12 >> the few ulp of the two roundings);
// inside the margin the correctly rounded sqrt decides, so the flag is bit-identical.
PLUME_HD bool reached_test(double s2, double radius, double& distance, bool need_distance) {
    if (!need_distance) {
        const double r2 = dmul(radius, radius);
        if (s2 < dmul(r2, 1.0 - 1e-12)) return true;
        if (s2 > dmul(r2, 1.0 + 1e-12)) return false;
    }
    distance = dsqrt(s2);
    return distance <= radius;
}

// P2 MethaneEnv.step, env:89-178.  `visited` points at this env's PLUME_VISIT_STRIDE counters.
// prev_tke: tke at the float32 cell before the move (env:105-108).  prev_conc_known / prev_cell_conc: the
// concentration there (env:93-95) if the caller has it; otherwise it is evaluated on demand -- the code model
// needs it only inside the boundary band.  kDistance = false skips the sqrt of env:155 wherever the reached
// test is decided without it (out.distance is then undefined).
template <bool kDistance = true, typename Field>
PLUME_HD void env_step(const Cfg& c, const Field& f, int env_local, uint32_t env_gid, EnvRegs& e, uint16_t* visited,
                       int action, double z0, double z1, bool prev_conc_known, double prev_cell_conc,
                       double prev_cell_tke, StepResult& out) {
    int ppx, ppy;
    cell32_of(c, e, ppx, ppy);                                                // env:93 (before the move)
    e.step += 1;                                                              // env:90
    // env:98-102
    const double ms = c.move_step;
    double dx = 0.0, dy = 0.0;
    if (action == 1) dy = ms;
    else if (action == 2) dy = -ms;
    else if (action == 3) dx = ms;
    else if (action == 4) dx = -ms;
    const double dnorm = (action == 0) ? 0.0 : ms;                            // ||(dx,dy)||, exact
    // -0.15 * (1 - ||d|| / move_step): ||d|| is exactly 0 or move_step, so the product is -0.15 or -0.0
    const double move_penalty = (action == 0) ? -0.15 : -0.0;
    // env:105-108   move_step*0.2*(randn(2)*tke/(TI*3))
    const double gain = dmul(ms, 0.2);
    const double tx = dmul(gain, div_nine(c, dmul(z0, prev_cell_tke)));
    const double ty = dmul(gain, div_nine(c, dmul(z1, prev_cell_tke)));
    // env:111-113
    double nx = dadd(dadd((double)e.px, dx), tx);
    double ny = dadd(dadd((double)e.py, dy), ty);
    nx = nx < 0.0 ? 0.0 : (nx > c.clip_hi ? c.clip_hi : nx);
    ny = ny < 0.0 ? 0.0 : (ny > c.clip_hi ? c.clip_hi : ny);
    e.px = (float)nx;
    e.py = (float)ny;
    // env:116-119
    const int cx = clip_cell((int)nx, c.G), cy = clip_cell((int)ny, c.G);
    int ox, oy;
    cell32_of(c, e, ox, oy);
    double conc64, tke64;
    f.eval(c, env_local, env_gid, e.episode, e.sx, e.sy, cx, cy, conc64, tke64);
    double conc32 = conc64, tke32 = tke64;
    if (ox != cx || oy != cy)   // float32 rounding of the position crossed a cell edge (rare)
        f.eval(c, env_local, env_gid, e.episode, e.sx, e.sy, ox, oy, conc32, tke32);
    // env:118-131.  boundary_dist = min over the four quotients v/G = (min v)/G (correctly rounded division is
    // monotonic), and the penalty -- the only consumer of conc_gradient in the code model -- is zero unless
    // boundary_dist < boundary_decay_start: the float64 divisions are evaluated only near the boundary (and
    // always in the README model, whose reward needs the concentrations).  Bit-identical results.
    const double G = (double)c.G;
    const double vmin = fmin(fmin(nx, dsub(G, nx)), fmin(ny, dsub(G, ny)));
    double cur_conc = 0.0, bpen = 0.0;
    const bool need_conc = c.plume_model == PLUME_MODEL_DISPERSION;
    double prev_conc = 0.0;
    if (need_conc || vmin < dmul(dadd(c.bnd_start, 1e-3), G)) {
        if (!prev_conc_known) {
            double unused;
            f.eval(c, env_local, env_gid, e.episode, e.sx, e.sy, ppx, ppy, prev_cell_conc, unused);
        }
        prev_conc = div_peak(c, prev_cell_conc);                              // env:95
        cur_conc = div_peak(c, conc64);
        // conc_gradient (env:119): the divisor ||d|| + 1e-6 is one of two constants
        const double dc = dsub(cur_conc, prev_conc);
        const double grad = !c.fastdiv ? ddiv(dc, dadd(dnorm, 1e-6))
                            : (action == 0 ? ddiv_const(dc, 1e-6, c.inv_eps) : ddiv_const(dc, c.ms_eps, c.inv_ms_eps));
        const double bd = c.fastdiv ? ddiv_const(vmin, c.G_d, c.inv_G_d) : ddiv(vmin, G);
        if (bd < c.bnd_start && grad < -0.01) {
            const double gap = dsub(c.bnd_start, bd);
            bpen = dmul(-c.bnd_penalty, dmul(gap, gap));
        }
    }
    // env:134-137: floor(nx / cell_size) == int(nx) / cell_size for 0 <= nx < grid (integer divisor)
    const int gx = (int)nx / c.cell_size, gy = (int)ny / c.cell_size;
    const int slot = gx * PLUME_MAX_GRID_DIVISIONS + gy;
    const int vc = (int)visited[slot] + 1;
    visited[slot] = (uint16_t)vc;
    // env:140-143
    const int slot32 = (ox / c.cell_size) * PLUME_MAX_GRID_DIVISIONS + (oy / c.cell_size);
    const int vis32 = slot32 == slot ? vc : (int)visited[slot32];
    make_obs(c, e, conc32, tke32, vis32, out.obs, env_gid);
    const float explore = fdiv(fmul((float)e.ebonus, fsub(1.0f, out.obs[5])), visit_denominator_f(c, vc));
    // env:146-152 (numpy>=2 promotion: float32 terms, float64 from move_penalty on)
    const float conc_reward = fmul((float)c.conc_coef, out.obs[2]);
    const float tke_term = fmul((float)c.tke_factor, out.obs[3]);
    double total = dadd((double)fadd(conc_reward, explore), move_penalty);
    total = dsub(total, (double)tke_term);
    total = dadd(total, bpen);
    float info_conc = conc_reward, info_explore = explore, info_tke = -tke_term;
    double info_move = move_penalty, info_bnd = bpen;
    if (c.plume_model == PLUME_MODEL_DISPERSION) {
        // README reward R = d[CH4] - 0.2 |d theta| (README.md:52,99): concentrations normalised by the peak,
        // theta = heading of the move (axis aligned: the change is 0, pi/2 or pi); staying keeps the heading
        const double dconc = dsub(cur_conc, prev_conc);
        double dtheta = 0.0;
        if (action != 0 && e.last_move != 0 && action != e.last_move)
            dtheta = ((action <= 2) == (e.last_move <= 2)) ? 3.141592653589793 : 1.5707963267948966;
        if (action != 0) e.last_move = action;
        total = dsub(dconc, dmul(0.2, dtheta));
        info_conc = (float)dconc;
        info_explore = 0.0f;
        info_tke = 0.0f;
        info_move = -dmul(0.2, dtheta);
        info_bnd = 0.0;
    }
    // env:155-158
    const double ex = dsub((double)e.px, e.sx), ey = dsub((double)e.py, e.sy);
    double distance = 0.0;
    const bool reached = reached_test(dadd(dmul(ex, ex), dmul(ey, ey)), e.radius, distance, kDistance);
    if (reached) total = dadd(total, fmin(500.0, dmul(150.0, ddiv(c.initial_radius, e.radius))));
    out.reward = total;
    out.reached = reached;
    out.done = (e.step >= c.max_steps) || reached;                            // env:161
    out.conc_reward = info_conc;
    out.explore_reward = info_explore;
    out.tke_penalty = info_tke;
    out.move_penalty = info_move;
    out.boundary_penalty = info_bnd;
    out.cur_conc = cur_conc;
    out.cell_conc = conc32;
    out.cell_tke = tke32;
    out.distance = distance;
}

// P2 with PLUME_FLAG_FAST_REWARD: everything that decides a FLAG or an INDEX (position update, float32 rounding
// of the position, cells, visit counters, distance <= radius, step limit) is the float64 arithmetic of env_step,
// line for line; the observation's concentration/tke entries and the reward terms are float32 with reciprocal
// multiplies (north_star bar: fp32 rel 1e-5; measured ~1e-7).  prev_cell_conc is the float32 field value.
template <bool kDistance = true, typename Field>
PLUME_HD void env_step_fast(const Cfg& c, const Field& f, int env_local, uint32_t env_gid, EnvRegs& e,
                            uint16_t* visited, int action, double z0, double z1, bool prev_conc_known,
                            float prev_cell_conc, double prev_cell_tke, StepResult& out) {
    int ppx, ppy;
    cell32_of(c, e, ppx, ppy);
    e.step += 1;                                                              // env:90
    const double ms = c.move_step;
    double dx = 0.0, dy = 0.0;
    if (action == 1) dy = ms;
    else if (action == 2) dy = -ms;
    else if (action == 3) dx = ms;
    else if (action == 4) dx = -ms;
    const float dnorm = (action == 0) ? 0.0f : (float)ms;
    const float move_penalty = (action == 0) ? -0.15f : 0.0f;                 // env:101-102
    // env:105-113, exact
    const double gain = dmul(ms, 0.2);
    const double tx = dmul(gain, div_nine(c, dmul(z0, prev_cell_tke)));
    const double ty = dmul(gain, div_nine(c, dmul(z1, prev_cell_tke)));
    double nx = dadd(dadd((double)e.px, dx), tx);
    double ny = dadd(dadd((double)e.py, dy), ty);
    nx = nx < 0.0 ? 0.0 : (nx > c.clip_hi ? c.clip_hi : nx);
    ny = ny < 0.0 ? 0.0 : (ny > c.clip_hi ? c.clip_hi : ny);
    e.px = (float)nx;
    e.py = (float)ny;
    // env:116-119
    const int cx = clip_cell((int)nx, c.G), cy = clip_cell((int)ny, c.G);
    int ox, oy;
    cell32_of(c, e, ox, oy);
    float conc64;
    double tke64;
    f.eval_fast(c, env_local, env_gid, e.episode, e.sx, e.sy, cx, cy, conc64, tke64);
    float conc32 = conc64;
    double tke32 = tke64;
    if (ox != cx || oy != cy) f.eval_fast(c, env_local, env_gid, e.episode, e.sx, e.sy, ox, oy, conc32, tke32);
    // env:118-131 (the gradient only matters inside the boundary band, or for the README reward)
    const float fx = (float)nx, fy = (float)ny, G = (float)c.G;
    const float bd = fminf(fminf(fx, G - fx), fminf(fy, G - fy)) * c.inv_G_f;
    float bpen = 0.0f, prev_conc = 0.0f, cur_conc = 0.0f;
    if (c.plume_model == PLUME_MODEL_DISPERSION || bd < (float)c.bnd_start) {
        if (!prev_conc_known) {
            double unused;
            f.eval_fast(c, env_local, env_gid, e.episode, e.sx, e.sy, ppx, ppy, prev_cell_conc, unused);
        }
        prev_conc = prev_cell_conc * c.inv_peak_f;
        cur_conc = conc64 * c.inv_peak_f;
        const float grad = (cur_conc - prev_conc) / (dnorm + 1e-6f);
        if (bd < (float)c.bnd_start && grad < -0.01f) {
            const float gap = (float)c.bnd_start - bd;
            bpen = -(float)c.bnd_penalty * gap * gap;
        }
    }
    // env:134-137: floor(nx / cell_size) == int(nx) / cell_size for 0 <= nx < G (integer divisor)
    const int gx = (int)nx / c.cell_size, gy = (int)ny / c.cell_size;
    const int slot = gx * PLUME_MAX_GRID_DIVISIONS + gy;
    const int vc = (int)visited[slot] + 1;
    visited[slot] = (uint16_t)vc;
    const int slot32 = (ox / c.cell_size) * PLUME_MAX_GRID_DIVISIONS + (oy / c.cell_size);
    const int vis32 = slot32 == slot ? vc : (int)visited[slot32];
    // env:71-87
    out.obs[0] = div_G(c, e.px);
    out.obs[1] = div_G(c, e.py);
    out.obs[2] = conc32 * c.inv_peak_f;
    out.obs[3] = (float)tke32 * c.inv_nine_f;
    out.obs[4] = step_fraction_f(c, e.step);
    out.obs[5] = fminf((float)vis32 * 0.2f, 1.0f);
    if (c.plume_model == PLUME_MODEL_DISPERSION) make_obs(c, e, (double)conc32, tke32, vis32, out.obs, env_gid);
    // env:140: vc ** 0.75 + 1
    float denom;
    if (c.visit_denom_tab) {
        denom = visit_denominator_f(c, vc);
    } else {
        const float v3 = (float)vc * (float)vc * (float)vc;
        denom = sqrtf(sqrtf(v3)) + 1.0f;
    }
    const float explore = (float)e.ebonus * (1.0f - out.obs[5]) / denom;
    const float conc_reward = (float)c.conc_coef * out.obs[2];
    const float tke_term = (float)c.tke_factor * out.obs[3];
    double total = (double)((((conc_reward + explore) + move_penalty) - tke_term) + bpen);    // env:146-152
    float info_conc = conc_reward, info_explore = explore, info_tke = -tke_term, info_move = move_penalty,
          info_bnd = bpen;
    if (c.plume_model == PLUME_MODEL_DISPERSION) {
        const float dconc = cur_conc - prev_conc;
        float dtheta = 0.0f;
        if (action != 0 && e.last_move != 0 && action != e.last_move)
            dtheta = ((action <= 2) == (e.last_move <= 2)) ? 3.14159265f : 1.57079633f;
        if (action != 0) e.last_move = action;
        total = (double)(dconc - 0.2f * dtheta);
        info_conc = dconc;
        info_explore = 0.0f;
        info_tke = 0.0f;
        info_move = -0.2f * dtheta;
        info_bnd = 0.0f;
    }
    // env:155-161, exact
    const double ex = dsub((double)e.px, e.sx), ey = dsub((double)e.py, e.sy);
    double distance = 0.0;
    const bool reached = reached_test(dadd(dmul(ex, ex), dmul(ey, ey)), e.radius, distance, kDistance);
    if (reached) total = dadd(total, fmin(500.0, dmul(150.0, ddiv(c.initial_radius, e.radius))));
    out.reward = total;
    out.reached = reached;
    out.done = (e.step >= c.max_steps) || reached;
    out.conc_reward = info_conc;
    out.explore_reward = info_explore;
    out.tke_penalty = info_tke;
    out.move_penalty = (double)info_move;
    out.boundary_penalty = (double)info_bnd;
    out.cur_conc = (double)cur_conc;
    out.cell_conc = (double)conc32;
    out.cell_tke = tke32;
    out.distance = distance;
}

// P0 reset, env:42-50: source draw, zero position/step, clear visit table, latch curriculum.
PLUME_HD void env_reset(const Cfg& c, uint32_t env_gid, EnvRegs& e, uint16_t* visited, const double* u_src,
                        double radius, double ebonus) {
    e.episode += 1;
    double ux, uy;
    if (u_src) {
        ux = u_src[0];
        uy = u_src[1];
    } else {
        const U4 r = philox4x32_10(0u, e.episode, env_gid, kTagSrc, c.k0, c.k1);
        ux = uniform53(r.x, r.y);
        uy = uniform53(r.z, r.w);
    }
    const double span = (double)(c.G - 100);                                  // env:43-44
    e.sx = dadd(dmul(ux, span), 50.0);
    e.sy = dadd(dmul(uy, span), 50.0);
    e.px = 0.0f;                                                              // env:46
    e.py = 0.0f;
    e.step = 0;                                                               // env:47
    e.last_move = 0;
    e.radius = radius;
    e.ebonus = ebonus;
    // env:49 (the per-env table is PLUME_VISIT_STRIDE * 2 = 208 B = 13 x 16 B, 16-byte aligned)
    static_assert(PLUME_VISIT_STRIDE % 8 == 0, "visit table rows are cleared 16 bytes at a time");
    struct alignas(16) Zero16 {
        uint64_t a, b;
    };
    Zero16* v16 = reinterpret_cast<Zero16*>(visited);
    for (int i = 0; i < PLUME_VISIT_STRIDE / 8; ++i) v16[i] = Zero16{0ull, 0ull};
}

PLUME_HD void step_noise(const Cfg& c, uint32_t env_gid, const EnvRegs& e, double& z0, double& z1) {
    const U4 r = philox4x32_10((uint32_t)e.step, e.episode, env_gid, kTagStep, c.k0, c.k1);
    float a, b;
    box_muller(r.x, r.y, a, b);
    z0 = (double)a;
    z1 = (double)b;
}

PLUME_HD float action_uniform(const Cfg& c, uint32_t env_gid, const EnvRegs& e) {
    const U4 r = philox4x32_10((uint32_t)e.step, e.episode, env_gid, kTagAct, c.k0, c.k1);
    return uniform24(r.x);
}

// ---------------------------------------------------------------------------------------
// stateless permutation of [0,n): 4-round Feistel over the enclosing power of 4, cycle-walked.
// Replaces torch.randperm (train_ppo2.0.py:43) without materialising the permutation.
// ---------------------------------------------------------------------------------------
PLUME_HD uint64_t feistel_permute(uint64_t idx, uint64_t n, uint64_t seed, uint32_t epoch) {
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < n) ++half_bits;
    const uint64_t half_mask = (1ull << half_bits) - 1;
    uint64_t v = idx;
    do {
        uint32_t l = (uint32_t)(v >> half_bits), r = (uint32_t)(v & half_mask);
#pragma unroll
        for (uint32_t round = 0; round < 4; ++round) {
            const U4 h = philox4x32_10(r, round, epoch, 0x5045524Du, (uint32_t)seed, (uint32_t)(seed >> 32));
            const uint32_t nl = r;
            r = (l ^ h.x) & (uint32_t)half_mask;
            l = nl;
        }
        v = ((uint64_t)l << half_bits) | r;
    } while (v >= n);
    return v;
}

}  // namespace plume
