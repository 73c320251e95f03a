"""The product's per-environment device code (csrc/plume_core.h, __host__ __device__) compiled
with g++ and checked bit-for-bit against the oracle: catches logic errors without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import philox as ph
from oracle import plume_oracle as po
from tests.helpers import golden_oracle_env, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim():
    src = os.path.join(ROOT, "tests", "host_sim", "env_sim.cpp")
    out = os.path.join(ROOT, "tests", "host_sim", "libenv_sim.so")
    deps = [src, os.path.join(ROOT, "uav-wrf-les-ppo-lstm_b200", "csrc", "plume_core.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", out])
    return C.CDLL(out)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def _ccfg(cfg, mode, seed):
    import uav_wrf_les_ppo_lstm_b200 as pb
    return pb._lib.make_env_config(cfg, mode, seed)


def test_philox_device_function(sim):
    out = (C.c_uint32 * 4)()
    sim.sim_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_field_noise_and_source_streams(sim):
    cfg = po.config_for("2.1")
    seed = 0xDEADBEEF12345
    ec = _ccfg(cfg, 0, seed)
    x = np.arange(0, 500, 7, dtype=np.int32)
    y = (x * 3 % 500).astype(np.int32)
    z = np.zeros(len(x), np.float32)
    u = np.zeros(len(x), np.float32)
    sim.sim_field_noise(C.byref(ec), 17, 3, P(x), P(y), len(x), P(z), P(u))
    z64, u64 = ph.field_noise64(seed, 17, 3, x.astype(np.int64) * 500 + y)
    assert np.abs(z - z64).max() < 5e-6 and np.array_equal(u, u64)
    src = np.zeros(2)
    sim.sim_source(C.byref(ec), 17, 3, P(src))
    ux, uy = ph.source_uniforms(seed, 17, 3)
    assert src[0] == ux * 400 + 50 and src[1] == uy * 400 + 50


@pytest.mark.parametrize("n", [1, 2, 5, 256, 1000, 4096, 5000])
def test_feistel_is_a_permutation(sim, n):
    out = np.zeros(n, np.int64)
    sim.sim_feistel(n, 99, 2, P(out))
    assert sorted(out.tolist()) == list(range(n))
    if n > 100:
        other = np.zeros(n, np.int64)
        sim.sim_feistel(n, 99, 3, P(other))
        assert (other != out).mean() > 0.9


def _set_variant(sim, variant, cfg):
    """0: previous-cell concentration passed in, no tables; 1: evaluated on demand + host tables + reached test
    without the sqrt (what the K2 kernel runs); 2: generic IEEE divisions.  Returns objects to keep alive."""
    import uav_wrf_les_ppo_lstm_b200 as pb
    sf, vd = pb.env.host_tables(cfg)
    sim.sim_set_variant(variant, P(sf), P(vd))
    return sf, vd


def test_constant_division_is_correctly_rounded(sim):
    """ddiv_const / fdiv_const (three FMA-pipe instructions) == IEEE division, on random numerators, on
    numerators constructed next to rounding boundaries of the quotient, and on float positions."""
    rng = np.random.default_rng(3)
    sim.sim_reciprocal_ok.argtypes = [C.c_double]
    sim.sim_div_const.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    for b in (9.0, 100.0, 500.0, 1000.0, 2.7):
        assert sim.sim_reciprocal_ok(b) == 1
        n = 200000
        a = np.ldexp(rng.random(n) + 1.0, rng.integers(-30, 30, n)) * rng.choice([-1.0, 1.0], n)
        q = np.ldexp(rng.random(n) + 1.0, rng.integers(-8, 8, n))
        mid = (q.astype(np.longdouble) + np.nextafter(q, np.inf).astype(np.longdouble)) / 2
        near = (mid * np.longdouble(b)).astype(np.float64)
        near = np.concatenate([near, np.nextafter(near, np.inf), np.nextafter(near, -np.inf)])
        z = rng.standard_normal(n).astype(np.float32).astype(np.float64) * (rng.random(n) * 14.0 - 0.9)
        for arr in (a, near, z):
            arr = np.ascontiguousarray(arr)
            out = np.zeros_like(arr)
            sim.sim_div_const(P(arr), len(arr), b, P(out))
            assert np.array_equal(out, arr / b)
    assert sim.sim_reciprocal_ok(0.0) == 0
    cfg = po.config_for("2.1")
    ec = _ccfg(cfg, 2, 0)
    assert sim.sim_fastdiv_enabled(C.byref(ec)) == 1
    pos = np.concatenate([(rng.random(2_000_000) * 500.0).astype(np.float32),
                          np.arange(0, 500, dtype=np.float32), np.float32(499.0) + np.zeros(1, np.float32)])
    outf = np.zeros_like(pos)
    sim.sim_div_G(C.byref(ec), P(pos), len(pos), P(outf))
    assert np.array_equal(outf, pos / np.float32(500.0))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("name", ["env_v21_s11.npz", "env_v21_s12.npz", "env_v20_s21.npz", "env_v11_s31.npz"])
def test_device_step_logic_against_reference_golden(sim, name, variant):
    """csrc env_step (host build) vs the trace the REAL reference produced."""
    g = load_golden(name)
    cfg, ora, z_steps = golden_oracle_env(g)
    ec = _ccfg(cfg, 2, 0)
    keep = _set_variant(sim, variant, cfg)
    px = np.zeros(1, np.float32); py = np.zeros(1, np.float32)
    step = np.zeros(1, np.int32); ep = np.ones(1, np.int32)
    vis = np.zeros((1, 104), np.uint16)
    rad = np.full(1, float(g["radius"])); eb = np.full(1, cfg.explore_bonus)
    sx = ora.src[:, 0].copy(); sy = ora.src[:, 1].copy()
    conc = np.ascontiguousarray(ora.fields.conc); tke = np.ascontiguousarray(ora.fields.tke)
    obs = np.zeros((1, 6), np.float32); rew = np.zeros(1); done = np.zeros(1, np.uint8)
    reached = np.zeros(1, np.uint8); info = np.zeros((1, 5))
    sim.sim_observe_f64(C.byref(ec), 1, P(px), P(py), P(sx), P(sy), P(step), P(vis), P(conc), P(tke), P(obs))
    assert np.array_equal(obs[0], g["obs0"])
    for t, a in enumerate(g["actions"]):
        act = np.array([a], np.int32)
        z = np.ascontiguousarray(z_steps[t][None])
        sim.sim_step_f64(C.byref(ec), 1, P(px), P(py), P(sx), P(sy), P(step), P(ep), P(vis), P(rad), P(eb), P(conc),
                         P(tke), P(act), P(z), P(obs), P(rew), P(done), P(reached), P(info))
        assert np.array_equal(obs[0], g["obs"][t]), t
        assert rew[0] == g["reward"][t]
        assert bool(done[0]) == g["done"][t] and bool(reached[0]) == g["reached"][t]
        assert np.array_equal(info[0], g["info"][t])
        assert px[0] == g["pos"][t, 0] and py[0] == g["pos"][t, 1]
    assert np.array_equal(vis[0, :100].reshape(10, 10).astype(np.int64), g["visited"])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("version", ["2.1", "1.1"])
def test_device_step_logic_random_walk(sim, version, variant):
    cfg = po.config_for(version)
    ec = _ccfg(cfg, 2, 1234)
    keep = _set_variant(sim, variant, cfg)
    n, T, G = 6, 300, cfg.grid_size
    rng = np.random.default_rng(5)
    ora = po.OracleVecEnv(cfg, n)
    usrc = rng.random((n, 2))
    zf = rng.standard_normal((n, G, G)).astype(np.float32)
    uf = rng.random((n, G, G)).astype(np.float32)
    for i in range(n):
        ora.reset_env(i, usrc[i], zf[i], uf[i])
    ora.current_radius[:] = 30.0
    px = np.zeros(n, np.float32); py = np.zeros(n, np.float32)
    step = np.zeros(n, np.int32); ep = np.ones(n, np.int32)
    vis = np.zeros((n, 104), np.uint16); rad = np.full(n, 30.0); eb = np.full(n, 0.6)
    sx = ora.src[:, 0].copy(); sy = ora.src[:, 1].copy()
    conc = np.ascontiguousarray(ora.fields.conc); tke = np.ascontiguousarray(ora.fields.tke)
    obs = np.zeros((n, 6), np.float32); rew = np.zeros(n); done = np.zeros(n, np.uint8)
    reached = np.zeros(n, np.uint8); info = np.zeros((n, 5))
    for t in range(T):
        a = rng.integers(0, 5, n).astype(np.int32)
        z = rng.standard_normal((n, 2)).astype(np.float32).astype(np.float64)
        o, r, d, inf = ora.step(a, z)
        sim.sim_step_f64(C.byref(ec), n, P(px), P(py), P(sx), P(sy), P(step), P(ep), P(vis), P(rad), P(eb), P(conc),
                         P(tke), P(a), P(z), P(obs), P(rew), P(done), P(reached), P(info))
        assert np.array_equal(o, obs) and np.array_equal(r, rew), t
        assert np.array_equal(d, done.astype(bool)) and np.array_equal(inf["reached"], reached.astype(bool))
        assert np.array_equal(ora.visited.reshape(n, 100), vis[:, :100].astype(np.int64))
