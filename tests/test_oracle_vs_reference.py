"""Pins oracle/ against the UNMODIFIED reference imported from /root/reference (container
only: skipped where the reference is absent, e.g. on the GPU box)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import plume_oracle as po
from oracle import ppo_oracle as pp
from oracle.ref_harness import (QueueFeed, load_reference, make_reference_env, reference_available,
                                run_reference_lines)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")
INFO_KEYS = ("concentration_reward", "explore_reward", "move_penalty", "tke_penalty", "boundary_penalty")


@pytest.mark.parametrize("version", ["1.1", "2.0", "2.1"])
def test_config_matches_reference(version):
    ref = load_reference(version).config
    cfg = po.config_for(version)
    assert (cfg.grid_size, cfg.max_steps, cfg.conc_peak, cfg.turbulence_intensity) == (
        ref.GRID_SIZE, ref.MAX_STEPS, ref.CONC_PEAK, ref.TURBULENCE_INTENSITY)
    assert (cfg.gamma, cfg.lam, cfg.clip_epsilon, cfg.entropy_beta, cfg.learning_rate, cfg.batch_size, cfg.epochs) == (
        ref.GAMMA, ref.LAMBDA, ref.CLIP_EPSILON, ref.ENTROPY_BETA, ref.LEARNING_RATE, ref.BATCH_SIZE, ref.EPOCHS)
    assert (cfg.explore_bonus, cfg.decay_factor, cfg.grid_divisions, cfg.initial_radius, cfg.min_radius,
            cfg.radius_decay, cfg.success_threshold, cfg.window_size) == (
        ref.EXPLORE_BONUS, ref.DECAY_FACTOR, ref.GRID_DIVISIONS, ref.INITIAL_RADIUS, ref.MIN_RADIUS,
        ref.RADIUS_DECAY, ref.SUCCESS_THRESHOLD, ref.WINDOW_SIZE)
    assert (cfg.conc_reward_coef, cfg.tke_penalty_factor, cfg.boundary_penalty, cfg.boundary_decay_start) == (
        ref.CONC_REWARD_COEF, ref.TKE_PENALTY_FACTOR, ref.BOUNDARY_PENALTY, ref.BOUNDARY_DECAY_START)
    if version == "2.1":
        assert cfg.sigma == ref.GAUSSIAN_RADIUS
    # the product's own config must agree with the oracle's
    import uav_wrf_les_ppo_lstm_b200 as pb
    pc = pb.config_for(version)
    for f in ("grid_size", "max_steps", "sigma", "clip_hi", "gamma", "lam", "clip_epsilon", "entropy_beta",
              "learning_rate", "batch_size", "epochs", "explore_bonus", "initial_radius", "min_radius",
              "radius_decay", "success_threshold", "window_size", "decay_factor", "conc_reward_coef",
              "tke_penalty_factor", "boundary_penalty", "boundary_decay_start"):
        assert getattr(pc, f) == getattr(cfg, f), f


@pytest.mark.parametrize("version,radius", [("2.1", 50.0), ("2.1", 9.0), ("2.0", 30.0), ("1.1", 12.0)])
def test_env_step_bit_exact(version, radius):
    cfg = po.config_for(version)
    n_env, n_steps, G = 3, 250, cfg.grid_size
    rng = np.random.default_rng(hash((version, radius)) % 2**32)
    ora = po.OracleVecEnv(cfg, n_env)
    usrc = rng.random((n_env, 2))
    zf = rng.standard_normal((n_env, G, G)).astype(np.float32)
    uf = rng.random((n_env, G, G)).astype(np.float32)
    for i in range(n_env):
        ora.reset_env(i, usrc[i], zf[i], uf[i])
    ora.current_radius[:] = radius
    zs = rng.standard_normal((n_steps, n_env, 2)).astype(np.float32)
    obs0 = ora.observe()
    # actions: biased walk towards the source so that `reached` and the arrival bonus are covered
    refs, feeds = [], []
    for i in range(n_env):
        feed = QueueFeed()
        feed.push_reset(usrc[i], zf[i], uf[i])
        env = make_reference_env(version, feed)
        env.current_radius = radius
        assert np.array_equal(env._get_obs(), obs0[i])
        # run the reference env i completely, recording its outputs
        rec = []
        for t in range(n_steps):
            d = env.source_pos - env.agent_pos
            if rng.random() < 0.5:
                a = int(rng.integers(0, 5))
            elif abs(d[0]) > abs(d[1]):
                a = 3 if d[0] > 0 else 4
            else:
                a = 1 if d[1] > 0 else 2
            feed.push_step(zs[t, i])
            o, r, dn, info = env.step(a)
            rec.append((a, o, r, dn, info, env.trajectory[-1]["reached"], dict(env.visited), env.agent_pos.copy()))
            if dn:
                break
        refs.append(rec)
    horizon = max(len(r) for r in refs)
    reached_any = False
    for t in range(horizon):
        acts = np.array([refs[i][t][0] if t < len(refs[i]) else 0 for i in range(n_env)])
        o, r, d, info = ora.step(acts, zs[t])
        for i in range(n_env):
            if t >= len(refs[i]):
                continue
            a, ro, rr, rd, rinfo, rreached, rvis, rpos = refs[i][t]
            assert np.array_equal(ro, o[i]), (version, i, t)
            assert rr == r[i] and type(rr) is np.float64
            assert rd == d[i] and rreached == info["reached"][i]
            for k in INFO_KEYS:
                assert rinfo[k] == info[k][i], (k, t)
            assert np.array_equal(rpos, ora.pos32[i])
            for (gx, gy), v in rvis.items():
                assert ora.visited[i, int(gx), int(gy)] == v
            reached_any |= bool(rreached)
    assert reached_any or radius > 40


def test_scalar_port_matches_vector_oracle():
    cfg = po.config_for("2.1")

    class Feed(np.random.Generator):
        pass
    rng = np.random.default_rng(3)
    sc = po.OracleScalarEnv(cfg, np.random.default_rng(42))
    rng2 = np.random.default_rng(42)
    usrc = rng2.random(2)
    z = rng2.standard_normal((cfg.grid_size, cfg.grid_size))
    u = rng2.random((cfg.grid_size, cfg.grid_size))
    vec = po.OracleVecEnv(cfg, 1)
    vec.reset_env(0, usrc, z, u)
    for t in range(100):
        a = int(rng.integers(0, 5))
        zz = rng2.standard_normal(2)
        o, r, d, info = vec.step(np.array([a]), zz[None])
        so, sr, sd, sinfo = sc.step(a)
        assert np.array_equal(so, o[0]) and sr == r[0] and sd == d[0]


def test_model_forward_and_keys():
    ref = load_reference("2.1")
    torch.manual_seed(0)
    m_ref = ref.model.PPOActorCritic(6, 5)
    m_or = pp.OracleActorCritic(6, 5)
    assert list(m_ref.state_dict().keys()) == list(m_or.state_dict().keys())
    m_or.load_state_dict(m_ref.state_dict())
    x = torch.rand(64, 6)
    p1, v1 = m_ref(x)
    p2, v2 = m_or(x)
    assert torch.equal(p1, p2) and torch.equal(v1, v2)
    a = torch.randint(0, 5, (64,))
    assert torch.equal(torch.distributions.Categorical(p1).log_prob(a), pp.categorical_log_prob(p2, a))
    # the product's module exposes the same keys/shapes (state_dict interchange)
    import uav_wrf_les_ppo_lstm_b200 as pb
    m_pb = pb.PPOActorCritic(device="cpu")
    assert {k: tuple(v.shape) for k, v in m_pb.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in m_ref.state_dict().items()}
    m_pb.load_state_dict(m_ref.state_dict())
    for k, (off, shape) in pb._lib.MLP_OFFSETS.items():
        n = int(np.prod(shape))
        assert torch.equal(m_pb.flat[off:off + n].view(shape), m_ref.state_dict()[k])
    m_ref.load_state_dict(m_pb.state_dict())


def test_update_model_bit_exact():
    ref = load_reference("2.1")
    cfg = po.config_for("2.1")
    torch.manual_seed(1)
    m_ref = ref.model.PPOActorCritic(6, 5)
    m_or = pp.OracleActorCritic(6, 5)
    m_or.load_state_dict(m_ref.state_dict())
    rng = np.random.default_rng(0)
    M = 300     # not a multiple of the minibatch: exercises the short last split
    S = rng.random((M, 6)).astype(np.float32)
    A = rng.integers(0, 5, M)
    R = rng.normal(size=M)
    with torch.no_grad():
        P, V = m_ref(torch.FloatTensor(S))
    LP = pp.categorical_log_prob(P, torch.LongTensor(A)).numpy() + 0.1 * rng.normal(size=M).astype(np.float32)
    D = rng.random(M) < 0.05
    buf = ref.model.PPOBuffer()
    for i in range(M):
        buf.store(S[i], A[i], R[i], V[i].item(), LP[i], D[i])
    o_ref = torch.optim.Adam(m_ref.parameters(), lr=cfg.learning_rate)
    o_or = torch.optim.Adam(m_or.parameters(), lr=cfg.learning_rate)
    torch.manual_seed(123)
    ref.train._update_model(buf, m_ref, o_ref)
    torch.manual_seed(123)
    perms = [torch.randperm(M) for _ in range(cfg.epochs)]
    st, ac, rw, va, lp, dn = buf.get()
    pp.ppo_update(m_or, o_or, st, ac, rw, va, lp, dn, cfg, perms=perms)
    for a, b in zip(m_ref.state_dict().values(), m_or.state_dict().values()):
        assert torch.equal(a, b)


def test_curriculum_bit_exact():
    ref = load_reference("2.1")
    cfg = po.config_for("2.1")

    class E:
        current_radius = 50.0
        explore_bonus = 0.6
    e1, e2 = E(), E()
    t1, t2 = ref.model.PPOTrainer(e1, None, None), pp.OracleCurriculum(e2, cfg)
    rng = np.random.default_rng(0)
    with contextlib.redirect_stdout(io.StringIO()):
        for k in range(1500):
            s = bool(rng.random() < min(0.9, 0.2 + k / 1000))
            t1.update(s)
            t2.update(s)
            assert t1.current_radius == t2.current_radius and t1.explore_bonus == t2.explore_bonus
            assert e1.current_radius == e2.current_radius and e1.explore_bonus == e2.explore_bonus
    assert t1.current_radius < 50.0
    # the product's host-side PPOTrainer.update is the same rule
    import uav_wrf_les_ppo_lstm_b200 as pb
    e3 = E()
    t3 = pb.PPOTrainer(e3, cfg=pb.config_for("2.1"))
    e4 = E()
    t4 = ref.model.PPOTrainer(e4, None, None)
    rng = np.random.default_rng(5)
    with contextlib.redirect_stdout(io.StringIO()):
        for k in range(1000):
            s = bool(rng.random() < 0.7)
            t3.update(s)
            t4.update(s)
            assert abs(t3.current_radius - t4.current_radius) < 1e-12 and abs(t3.explore_bonus - t4.explore_bonus) < 1e-12


def test_lstm_heads():
    ref = load_reference("2.1")
    torch.manual_seed(2)
    l_ref = ref.evaluate_with_lstm.PeakAndStopPredictor(input_dim=1)
    l_or = pp.OraclePeakAndStop()
    assert list(l_ref.state_dict().keys()) == list(l_or.state_dict().keys())
    l_or.load_state_dict(l_ref.state_dict())
    x = torch.rand(8, 20, 1)
    with torch.no_grad():
        for a, b in zip(l_ref(x), l_or(x)):
            assert torch.equal(a, b)
    ref20 = load_reference("2.0")
    t_ref = ref20.model.ConcentrationThresholdPredictor().eval()
    t_or = pp.OracleThresholdPredictor().eval()
    t_or.load_state_dict(t_ref.state_dict())
    x = torch.rand(1, 10, 1)
    with torch.no_grad():
        assert torch.equal(t_ref(x, [10]), t_or(x, [10]))        # the hot path is batch 1
        xb = torch.rand(4, 10, 1)
        assert torch.equal(t_ref(xb, [10] * 4), t_or(xb, [10] * 4))


def test_trend_label():
    ref = load_reference("2.1")
    rng = np.random.default_rng(1)
    for _ in range(20):
        conc = rng.random(20) * 100
        pos = rng.random((20, 2)) * 499
        src = rng.random(2) * 400 + 50
        want = ref.model.calculate_dynamic_label({"concentrations": conc, "positions": pos, "source_pos": src})
        assert pp.trend_label(conc, pos[-1], src)[0] == want


def _gae_case(seed, T=257):
    rng = np.random.RandomState(seed)
    rewards = torch.from_numpy(rng.randn(T).astype(np.float32) * 3)
    values = torch.from_numpy(rng.randn(T).astype(np.float32))
    dones = torch.from_numpy((rng.rand(T) < 0.08).astype(np.float32))
    dones[-1] = float(seed % 2)
    next_value = torch.from_numpy(rng.randn(1).astype(np.float32))
    return rewards, values, dones, next_value


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("script", ["PPOV1.1/train_ppo1.0.py", "PPOV1.0/ppo0.0.py", "PPOV1.1/train_ppo_gail.py"])
def test_older_gae_loop_bootstrap_bit_exact(script, seed):
    """P5' bootstrap variant: the reference's own inline loop (its source lines, executed) == oracle, bit for bit."""
    import os
    from oracle.ref_harness import REFERENCE_ROOT
    if not os.path.exists(os.path.join(REFERENCE_ROOT, script)):
        pytest.skip("script not in this reference tree")
    rewards, values, dones, next_value = _gae_case(seed)
    ns = dict(torch=torch, rewards=rewards.clone(), values=values.clone(), dones=dones.clone(),
              next_value=next_value.clone(), GAMMA=0.99, LAMBDA=0.95)
    try:
        run_reference_lines(script, "advantages = torch.zeros_like(rewards)", "advantages.std() + 1e-8", ns)
    except StopIteration:
        pytest.skip("this script has no such loop")
    adv, ret = pp.gae_bootstrap_v10(rewards, values, dones, next_value, 0.99, 0.95)
    assert torch.equal(adv, ns["advantages"]) and torch.equal(ret, ns["returns"])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_older_gae_loop_v12_bit_exact(seed):
    """P5' PPOV1.2 variant, same method."""
    rewards, values, dones, _ = _gae_case(seed)
    ns = dict(torch=torch, rewards=rewards.clone(), values=values.clone(), dones=dones.clone(), GAMMA=0.99,
              LAMBDA=0.95)
    run_reference_lines("PPOV1.2/ppo注释版.py", "advantages = torch.zeros_like(rewards)", "returns = advantages + values", ns)
    adv, ret = pp.gae_v12(rewards, values, dones, 0.99, 0.95)
    assert torch.equal(adv, ns["advantages"]) and torch.equal(ret, ns["returns"])
