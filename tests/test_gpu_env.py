"""GPU parity of the plume environment kernels (P0-P3, K1/K2) through the C ABI, against the
oracle and against the golden traces the real reference produced."""
import numpy as np
import pytest
import torch

from oracle import philox as ph
from oracle import plume_oracle as po
from tests.helpers import INFO_KEYS, golden_env_stream, load_golden

pytestmark = pytest.mark.gpu

ENV_FIXTURES = ["env_v21_s11.npz", "env_v21_s12.npz", "env_v20_s21.npz", "env_v11_s31.npz"]


def pb():
    import uav_wrf_les_ppo_lstm_b200 as m
    return m


@pytest.mark.parametrize("name", ENV_FIXTURES)
def test_step_kernel_reproduces_reference_trace(name):
    """K2 with the reference's own float64 fields, actions and randn draws: bit-exact."""
    g = load_golden(name)
    cfg, u_src, z_field, u_field, z_steps = golden_env_stream(g)
    conc, tke = po.plume_fields(cfg, u_src * (cfg.grid_size - 100) + 50, z_field, u_field)
    env = pb().VecMethaneEnv(1, version=str(g["version"]), field_mode="f64")
    obs0 = env.reset(u_src=u_src[None], fields=(conc[None], tke[None]))
    env.current_radius = float(g["radius"])
    assert np.array_equal(env.source_pos.cpu().numpy()[0], g["source_pos"])
    assert np.array_equal(obs0.cpu().numpy()[0], g["obs0"])
    for t, a in enumerate(g["actions"]):
        o, r, d, info = env.step(torch.tensor([int(a)]), step_noise=z_steps[t][None])
        assert np.array_equal(o.cpu().numpy()[0], g["obs"][t]), t
        assert r.cpu().numpy()[0] == g["reward"][t], t
        assert bool(d[0]) == g["done"][t] and bool(info["reached"][0]) == g["reached"][t]
        got = np.array([float(info[k][0]) for k in INFO_KEYS])
        assert np.allclose(got, g["info"][t].astype(np.float32), rtol=0, atol=0), t
        assert np.array_equal(env.agent_pos.cpu().numpy()[0], g["pos"][t])
    assert np.array_equal(env.visited.cpu().numpy()[0].astype(np.int64), g["visited"])


@pytest.mark.parametrize("version", ["2.1", "2.0", "1.1"])
def test_step_kernel_batch_vs_oracle(version):
    cfg = po.config_for(version)
    n, T, G = 24, 300, cfg.grid_size
    rng = np.random.default_rng(17)
    ora = po.OracleVecEnv(cfg, n)
    usrc = rng.random((n, 2))
    zf = rng.standard_normal((n, G, G)).astype(np.float32)
    uf = rng.random((n, G, G)).astype(np.float32)
    for i in range(n):
        ora.reset_env(i, usrc[i], zf[i], uf[i])
    env = pb().VecMethaneEnv(n, version=version, field_mode="f64")
    obs0 = env.reset(u_src=usrc, fields=(ora.fields.conc, ora.fields.tke))
    ora.current_radius[:] = 25.0
    env.current_radius = 25.0
    assert np.array_equal(obs0.cpu().numpy(), ora.observe())
    reached_any = False
    for t in range(T):
        d_src = ora.src - ora.pos32
        greedy = np.where(np.abs(d_src[:, 0]) > np.abs(d_src[:, 1]), np.where(d_src[:, 0] > 0, 3, 4),
                          np.where(d_src[:, 1] > 0, 1, 2))
        a = np.where(rng.random(n) < 0.6, rng.integers(0, 5, n), greedy).astype(np.int32)
        z = rng.standard_normal((n, 2)).astype(np.float32).astype(np.float64)
        o, r, d, info = ora.step(a, z)
        go, gr, gd, ginfo = env.step(torch.from_numpy(a), step_noise=z)
        assert np.array_equal(go.cpu().numpy(), o), t
        assert np.array_equal(gr.cpu().numpy(), r), t
        assert np.array_equal(gd.cpu().numpy(), d) and np.array_equal(ginfo["reached"].cpu().numpy(), info["reached"])
        for k in INFO_KEYS:
            assert np.array_equal(ginfo[k].cpu().numpy(), np.asarray(info[k], dtype=np.float32)), (k, t)
        assert np.array_equal(env.visited.cpu().numpy().astype(np.int64), ora.visited)
        reached_any |= bool(info["reached"].any())
    assert reached_any


def test_philox_streams_on_device():
    """Source draw, field noise and step noise of the device generator vs the numpy Philox."""
    seed, base = 0x1234ABCD5678, 1000
    env = pb().VecMethaneEnv(40, version="2.1", seed=seed, env_id_base=base, field_mode="procedural")
    ids = np.arange(40) + base
    ux, uy = ph.source_uniforms(seed, ids, np.ones(40, dtype=np.int64))
    src = env.source_pos.cpu().numpy()
    assert np.array_equal(src[:, 0], ux * 400 + 50) and np.array_equal(src[:, 1], uy * 400 + 50)
    z, u = env.field_noise(env_ids=[0, 7, 39])
    z, u = z.cpu().numpy(), u.cpu().numpy()
    cells = np.arange(500 * 500)
    for j, e in enumerate([0, 7, 39]):
        z64, u64 = ph.field_noise64(seed, base + e, 1, cells)
        assert np.array_equal(u[j].reshape(-1), u64.astype(np.float32))
        assert np.abs(z[j].reshape(-1) - z64).max() < 5e-5      # float32 fast intrinsics vs float64 Box-Muller
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    # lookups of single cells agree with the dump
    xs = np.array([0, 3, 499, 250], dtype=np.int32)
    ys = np.array([0, 498, 499, 1], dtype=np.int32)
    zz, uu = env.field_noise_at(np.array([7, 7, 7, 7], dtype=np.int32), xs, ys)
    assert np.array_equal(zz.cpu().numpy(), z[1][xs, ys]) and np.array_equal(uu.cpu().numpy(), u[1][xs, ys])
    # step noise
    noise = torch.zeros(40, 2, dtype=torch.float64, device=env.device)
    env.step(torch.zeros(40, dtype=torch.int32), noise_out=noise)
    z0, z1 = ph.step_noise64(seed, ids, np.ones(40, dtype=np.int64), np.zeros(40, dtype=np.int64))   # counter = step_count before the step
    got = noise.cpu().numpy()
    assert np.abs(got[:, 0] - z0).max() < 2e-5 and np.abs(got[:, 1] - z1).max() < 2e-5


@pytest.mark.parametrize("mode,rtol", [("f64", 1e-13), ("f32", 2e-6)])
def test_field_generation_vs_oracle(mode, rtol):
    """K1: materialised fields from the Philox stream == oracle fields built from the dumped draws."""
    cfg = po.config_for("2.1")
    env = pb().VecMethaneEnv(6, version="2.1", seed=99, field_mode=mode)
    z, u = env.field_noise()
    src = env.source_pos.cpu().numpy()
    conc = env.conc_field_t.cpu().numpy().astype(np.float64)
    tke = env.tke_field_t.cpu().numpy().astype(np.float64)
    for i in range(6):
        c_ref, t_ref = po.plume_fields(cfg, src[i], z[i].cpu().numpy(), u[i].cpu().numpy())
        assert np.allclose(tke[i], t_ref, rtol=rtol, atol=rtol)
        assert np.allclose(conc[i], c_ref, rtol=rtol, atol=rtol * 100)
    # a partial regeneration only touches the listed envs
    before = env.tke_field_t.clone()
    env.reset(env_ids=[2, 4])
    after = env.tke_field_t
    for i in range(6):
        same = torch.equal(before[i], after[i])
        assert same == (i not in (2, 4))
    assert env.episode_idx.cpu().tolist() == [1, 1, 2, 1, 2, 1]


def test_procedural_mode_vs_oracle():
    """Procedural lookups (no field in memory) against the oracle fed with the dumped draws:
    flags/cells bit-exact, float64 reward to 1e-12 (device exp vs numpy exp), obs bit-exact
    except the concentration entry (1 float32 ulp)."""
    cfg = po.config_for("2.1")
    n, T = 12, 400
    env = pb().VecMethaneEnv(n, version="2.1", seed=5, field_mode="procedural")
    z, u = env.field_noise()
    ora = po.OracleVecEnv(cfg, n)
    src = env.source_pos.cpu().numpy()
    for i in range(n):
        ora.set_source(i, src[i], z[i].cpu().numpy(), u[i].cpu().numpy())
    env.current_radius = 20.0
    ora.current_radius[:] = 20.0
    rng = np.random.default_rng(3)
    noise = torch.zeros(n, 2, dtype=torch.float64, device=env.device)
    o0 = env.observe().cpu().numpy()
    assert np.allclose(o0, ora.observe(), rtol=2e-7, atol=1e-9)
    done_seen = np.zeros(n, dtype=bool)
    for t in range(T):
        d_src = ora.src - ora.pos32
        greedy = np.where(np.abs(d_src[:, 0]) > np.abs(d_src[:, 1]), np.where(d_src[:, 0] > 0, 3, 4),
                          np.where(d_src[:, 1] > 0, 1, 2))
        a = np.where(rng.random(n) < 0.5, rng.integers(0, 5, n), greedy).astype(np.int32)
        go, gr, gd, ginfo = env.step(torch.from_numpy(a), noise_out=noise)
        o, r, d, info = ora.step(a, noise.cpu().numpy())
        live = ~done_seen
        go = go.cpu().numpy()
        assert np.array_equal(gd.cpu().numpy()[live], d[live]), t
        assert np.array_equal(ginfo["reached"].cpu().numpy()[live], info["reached"][live])
        assert np.array_equal(go[live][:, [0, 1, 4, 5]], o[live][:, [0, 1, 4, 5]]), t
        assert np.allclose(go[live], o[live], rtol=2e-7, atol=1e-9)
        assert np.allclose(gr.cpu().numpy()[live], r[live], rtol=1e-9, atol=1e-9)
        assert np.array_equal(env.visited.cpu().numpy().astype(np.int64)[live], ora.visited[live])
        done_seen |= d
    assert done_seen.any()


def test_auto_reset_step():
    n = 512
    env = pb().VecMethaneEnv(n, version="2.1", seed=1, field_mode="procedural", auto_reset=True)
    env.curriculum[0] = 200.0          # huge radius: many envs finish immediately after the next reset
    env.reset()
    rng = np.random.default_rng(0)
    total_done = np.zeros(n, dtype=np.int64)
    for t in range(30):
        o, r, d, info = env.step(torch.from_numpy(rng.integers(1, 5, n).astype(np.int32)))
        d = d.cpu().numpy()
        total_done += d
        # reset envs show the reset observation (position 0, step 0, no visits)
        oo = o.cpu().numpy()
        assert np.all(oo[d][:, [0, 1, 4, 5]] == 0)
        assert np.array_equal(oo, env.observe().cpu().numpy())
    ep = env.episode_idx.cpu().numpy()
    assert np.array_equal(ep, 2 + total_done)      # constructor reset + explicit reset + one per done
    assert total_done.sum() > 0
    assert np.all(env.radius_t.cpu().numpy() == 200.0)


@pytest.mark.parametrize("fast", [False, True])
def test_carried_cell_values_are_transparent(fast):
    """K2 carries tke / concentration of the current cell from one step to the next (cell_tke, cell_conc, tagged by
    cell_key).  An env whose tags are wiped before every step (everything recomputed from the Philox stream) must
    produce bit-identical results, also across auto-resets, host edits of the position and of the source."""
    n, T = 2048, 160
    mk = lambda: pb().VecMethaneEnv(n, version="2.1", seed=5, field_mode="procedural", auto_reset=True,
                                    fast_reward=fast)
    a, b = mk(), mk()
    for env in (a, b):
        env.curriculum[0] = 60.0
        env.reset()
    rng = np.random.default_rng(1)
    hits = 0
    for t in range(T):
        act = torch.from_numpy(rng.integers(0, 5, n).astype(np.int32))
        if t == 50:                     # host moves half of the agents: the tags of `a` no longer match
            for env in (a, b):
                env.pos_x[: n // 2] = 123.25
                env.pos_y[: n // 2] = 77.5
        if t == 90:                     # host moves the sources: set_source drops the carried concentration
            src = np.stack([np.full(n, 300.0), np.full(n, 40.0)], axis=1)
            a.set_source(src)
            b.set_source(src)
        b.cell_key_t.zero_()
        hits += int((a.cell_key_t != 0).sum().item())
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), t
        for k in ia:
            assert torch.equal(ia[k], ib[k]), (t, k)
        assert torch.equal(a.visited, b.visited) and torch.equal(a.agent_pos, b.agent_pos)
    assert hits >= n * (T - 2)          # the carried values were in use in `a` (not after reset / set_source)
    assert int(a.episode_idx.max().item()) > 2


def test_errors_are_loud():
    m = pb()
    with pytest.raises(ValueError):
        m.VecMethaneEnv(4, field_mode="f64", auto_reset=True)
    env = m.VecMethaneEnv(4, field_mode="procedural")
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(5, dtype=torch.int32))       # wrong number of actions
    import ctypes as C
    lib = m._lib.load()
    rc = lib.plume_env_step(C.byref(env.c_config), C.byref(env.c_state), None, None, 0, None, None, None, None,
                            None, None, None, None)
    assert rc != 0 and b"null" in lib.plume_last_error()


def test_readme_dispersion_model_vs_oracle():
    """P1' (README.md:48-53,95-100): Gaussian dispersion sigma_y = 0.3 x^0.71 in the wind-rotated frame,
    observation [x, y, CH4, wind_x, step, wind_y], reward dCH4 - 0.2 |dtheta| (+ arrival bonus).  There is no
    reference code behind the README: the oracle is our own CPU restatement (parity UNPINNED); device
    pow/exp/cos differ from numpy's in the last bits, so values are compared to 1e-9 and flags exactly."""
    from dataclasses import replace
    from oracle import philox as ph
    cfg = replace(po.config_for("2.1"), plume_model="dispersion")
    n, T, seed = 16, 300, 8
    env = pb().VecMethaneEnv(n, version="2.1", seed=seed, field_mode="procedural", plume_model="dispersion")
    z, u = env.field_noise()
    ora = po.OracleVecEnv(cfg, n, fields=po.CellNoiseFields(
        cfg, n, lambda idx, x, y: (z.cpu().numpy()[idx, x, y], u.cpu().numpy()[idx, x, y])))
    src = env.source_pos.cpu().numpy()
    ep = env.episode_idx.cpu().numpy()
    for i in range(n):
        ora.set_source(i, src[i])
        ora.set_wind(i, ph.wind(seed, i, int(ep[i])))
    assert len({round(float(w), 6) for w in ora.wind[:, 0]}) > 8        # every env has its own wind
    env.current_radius = 25.0
    ora.current_radius[:] = 25.0
    rng = np.random.default_rng(4)
    noise = torch.zeros(n, 2, dtype=torch.float64, device=env.device)
    o0 = env.observe().cpu().numpy()
    assert np.allclose(o0, ora.observe(), rtol=1e-6, atol=1e-7)
    assert np.allclose(o0[:, 3] ** 2 + o0[:, 5] ** 2, (ora.wind[:, 2] / 5.0) ** 2, rtol=1e-5)
    done_seen = np.zeros(n, dtype=bool)
    turn_pen = 0
    for t in range(T):
        d_src = ora.src - ora.pos32
        greedy = np.where(np.abs(d_src[:, 0]) > np.abs(d_src[:, 1]), np.where(d_src[:, 0] > 0, 3, 4),
                          np.where(d_src[:, 1] > 0, 1, 2))
        a = np.where(rng.random(n) < 0.5, rng.integers(0, 5, n), greedy).astype(np.int32)
        go, gr, gd, ginfo = env.step(torch.from_numpy(a), noise_out=noise)
        o, r, d, info = ora.step(a, noise.cpu().numpy())
        live = ~done_seen
        assert np.array_equal(gd.cpu().numpy()[live], d[live]), t
        assert np.array_equal(ginfo["reached"].cpu().numpy()[live], info["reached"][live])
        assert np.allclose(go.cpu().numpy()[live], o[live], rtol=1e-6, atol=1e-7)
        assert np.allclose(gr.cpu().numpy()[live], r[live], rtol=1e-9, atol=1e-9), t
        assert np.allclose(ginfo["move_penalty"].cpu().numpy()[live], info["move_penalty"][live], atol=1e-6)
        turn_pen += int((info["move_penalty"][live] < 0).sum())
        done_seen |= d
    assert done_seen.any() and turn_pen > 0
    # materialised float64 fields of the same model agree with the procedural lookups
    env2 = pb().VecMethaneEnv(2, version="2.1", seed=seed, field_mode="f64", plume_model="dispersion")
    z2, u2 = env2.field_noise()
    src2 = env2.source_pos.cpu().numpy()
    conc = env2.conc_field_t.cpu().numpy()
    for i in range(2):
        wc, ws, _ = ph.wind(seed, i, 1)
        xs, ys = np.mgrid[:500, :500]
        ref, _ = po.plume_cells(cfg, src2[i, 0], src2[i, 1], xs, ys, z2[i].cpu().numpy(), u2[i].cpu().numpy(),
                                (wc, ws))
        assert np.allclose(conc[i], ref, rtol=1e-9, atol=1e-9)
        assert (ref > 50).sum() > 50 and (ref > 50).sum() < 20000      # a narrow plume, not a disc


@pytest.mark.parametrize("version,mode", [("2.1", "procedural"), ("1.1", "procedural"), ("2.0", "f64")])
def test_fast_reward_mode_keeps_flags_and_indices_exact(version, mode):
    """PLUME_FLAG_FAST_REWARD: positions, cells, visit tables, reached/done flags are bit-identical to the exact
    kernel (float64 arithmetic of environment.py:105-117,134-137,155-161); observations and rewards agree to the
    north-star tolerance (fp32 rel 1e-5)."""
    m = pb()
    n, T = 256, 300
    kw = dict(version=version, seed=13, field_mode=mode)
    ex = m.VecMethaneEnv(n, **kw)
    fa = m.VecMethaneEnv(n, fast_reward=True, **kw)
    for env in (ex, fa):
        env.current_radius = 30.0
    rng = np.random.default_rng(1)
    noise = torch.zeros(n, 2, dtype=torch.float64, device=ex.device)
    done_seen = torch.zeros(n, dtype=torch.bool, device=ex.device)
    worst = 0.0
    for t in range(T):
        src = ex.source_pos - ex.agent_pos.double()
        greedy = torch.where(src[:, 0].abs() > src[:, 1].abs(), torch.where(src[:, 0] > 0, 3, 4),
                             torch.where(src[:, 1] > 0, 1, 2))
        rnd = torch.from_numpy(rng.integers(0, 5, n)).to(ex.device)
        a = torch.where(torch.from_numpy(rng.random(n) < 0.4).to(ex.device), rnd, greedy).int()
        o1, r1, d1, i1 = ex.step(a, noise_out=noise)
        o1, r1, d1 = o1.clone(), r1.clone(), d1.clone()
        o2, r2, d2, i2 = fa.step(a, step_noise=noise)
        live = ~done_seen
        assert torch.equal(d1[live], d2[live]) and torch.equal(i1["reached"][live], i2["reached"][live]), t
        assert torch.equal(ex.pos_x[live], fa.pos_x[live]) and torch.equal(ex.pos_y[live], fa.pos_y[live]), t
        assert torch.equal(ex.visited_t[live], fa.visited_t[live])
        assert torch.equal(o1[live][:, [0, 1]], o2[live][:, [0, 1]])
        assert torch.allclose(o1[live], o2[live], rtol=1e-5, atol=1e-6)
        assert torch.allclose(r1[live], r2[live], rtol=1e-5, atol=1e-5), t
        worst = max(worst, float((r1[live] - r2[live]).abs().max())) if bool(live.any()) else worst
        done_seen |= d1
    assert bool(done_seen.any()) and worst < 1e-5
