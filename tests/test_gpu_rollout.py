"""GPU parity of the fused persistent rollout kernel: the transitions it produced (with its own
Philox draws and sampled actions) are replayed through the oracle; the fused kernel must also
agree bit-for-bit with the discrete step/act kernels it is built from."""
import numpy as np
import pytest
import torch

from oracle import plume_oracle as po
from oracle import ppo_oracle as pp

pytestmark = pytest.mark.gpu


def pb():
    import uav_wrf_les_ppo_lstm_b200 as m
    return m


def _setup(N, T, seed, with_head=True, radius=None, actor_gain=40.0, plume_model="isotropic"):
    m = pb()
    torch.manual_seed(seed)
    env = m.VecMethaneEnv(N, version="2.1", seed=seed, field_mode="procedural", auto_reset=True,
                          plume_model=plume_model)
    if radius is not None:
        env.curriculum[0] = radius
        env.reset()
    model = m.PPOActorCritic(device="cuda")
    with torch.no_grad():
        model.actor.weight.mul_(actor_gain)
    head = None
    if with_head:
        head = m.PeakAndStopPredictor(device="cuda")
        with torch.no_grad():
            head.fc_stop[0].bias.fill_(1.386)       # sigmoid(1.386) = 0.8: flags straddle the threshold
            head.fc_stop[0].weight.mul_(40.0)
            head.lstm.weight_ih_l0.mul_(6.0)
    eng = m.RolloutEngine(env, model, head, horizon=T, with_info=True, with_trend=True)
    return m, env, model, head, eng


def test_rollout_replay_through_oracle():
    N, T = 96, 80
    m, env, model, head, eng = _setup(N, T, seed=11, radius=30.0)
    cfg = po.config_for("2.1")
    src0 = env.source_pos.cpu().numpy().copy()
    ep0 = env.episode_idx.cpu().numpy().copy()
    noise = torch.zeros(T, N, 2, dtype=torch.float64, device="cuda")
    buf = eng.collect(noise_out=noise)
    eng.check_nan()
    obs = buf.obs.cpu().numpy()
    acts = buf.actions.cpu().numpy()
    rew = buf.rewards.cpu().numpy()
    dones = buf.dones.cpu().numpy() != 0
    reached = buf.reached.cpu().numpy() != 0
    epi = buf.episode_idx.cpu().numpy()
    info = buf.info.cpu().numpy()
    zs = noise.cpu().numpy()

    # (1) env transitions of every env's first episode == oracle on the same actions/noise
    frozen = m.VecMethaneEnv(N, version="2.1", seed=11, field_mode="procedural")
    frozen.curriculum[0] = 30.0
    frozen.reset()                                  # same episode index (2) as the rollout env at start
    assert np.array_equal(frozen.episode_idx.cpu().numpy(), ep0)

    def noise_cb(idx, x, y):
        z, u = frozen.field_noise_at(np.asarray(idx, dtype=np.int32), np.asarray(x, dtype=np.int32),
                                     np.asarray(y, dtype=np.int32))
        return z.cpu().numpy(), u.cpu().numpy()

    ora = po.OracleVecEnv(cfg, N, fields=po.CellNoiseFields(cfg, N, noise_cb))
    for i in range(N):
        ora.set_source(i, src0[i])
    ora.current_radius[:] = 30.0
    alive = np.ones(N, dtype=bool)
    assert np.allclose(ora.observe(), obs[0], rtol=2e-7, atol=1e-9)
    for t in range(T):
        o, r, d, inf = ora.step(acts[t], zs[t])
        assert np.array_equal(d[alive], dones[t][alive]), t
        assert np.array_equal(inf["reached"][alive], reached[t][alive])
        assert np.allclose(r[alive], rew[t][alive], rtol=1e-5, atol=1e-6)
        assert np.allclose(np.asarray(inf["explore_reward"], dtype=np.float32)[alive], info[t, 1][alive], rtol=1e-6)
        assert np.all(epi[t][alive] == ep0[alive])
        if t + 1 < T:
            cont = alive & ~d
            assert np.array_equal(o[cont][:, [0, 1, 4, 5]], obs[t + 1][cont][:, [0, 1, 4, 5]]), t
            assert np.allclose(o[cont], obs[t + 1][cont], rtol=2e-7, atol=1e-9)
            fresh = alive & d                       # auto-reset: next obs is a reset observation
            assert np.all(obs[t + 1][fresh][:, [0, 1, 4, 5]] == 0)
            assert np.all(epi[t + 1][fresh] == ep0[fresh] + 1)
        alive &= ~d
    assert (~alive).sum() > 5                       # several envs reached the source

    # (2) policy outputs on the recorded observations == oracle MLP
    ora_m = pp.OracleActorCritic()
    ora_m.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    with torch.no_grad():
        p, v = ora_m(torch.from_numpy(obs.reshape(-1, 6)))
    lp = pp.categorical_log_prob(p, torch.from_numpy(acts.reshape(-1)).long())
    assert torch.allclose(buf.values.cpu().reshape(-1), v.squeeze(-1), rtol=1e-5, atol=2e-6)
    assert torch.allclose(buf.log_probs.cpu().reshape(-1), lp, rtol=1e-5, atol=2e-6)
    assert len(set(acts.reshape(-1).tolist())) == 5

    # (3) stop head == oracle LSTM on the per-episode sliding window of obs[2] after each step
    ora_l = pp.OraclePeakAndStop()
    ora_l.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    sp = buf.stop_prob.cpu().numpy()
    sf = buf.stop_flag.cpu().numpy() != 0
    last = eng.last_obs.cpu().numpy()
    hist = [[] for _ in range(N)]
    wins, where = [], []
    for t in range(T):
        nxt = obs[t + 1] if t + 1 < T else last
        for i in range(N):
            if dones[t, i]:
                # the post-step observation of a finished episode is not in obs[t+1]; use the info term
                hist[i].append(info[t, 0, i] / 2.0)
            else:
                hist[i].append(nxt[i, 2])
            if len(hist[i]) >= 20:
                wins.append(hist[i][-20:])
                where.append((t, i))
            else:
                assert sp[t, i] == 0 and not sf[t, i]
            if dones[t, i]:
                hist[i] = []
    assert len(wins) > 1000
    with torch.no_grad():
        _, s_ref = ora_l(torch.tensor(np.array(wins, dtype=np.float32)).unsqueeze(-1))
    got = np.array([sp[t, i] for t, i in where])
    assert np.allclose(got, s_ref.numpy(), rtol=1e-5, atol=2e-6)
    gotf = np.array([sf[t, i] for t, i in where])
    clear = np.abs(s_ref.numpy() - 0.8) > 1e-5
    assert np.array_equal(gotf[clear], (s_ref.numpy() > 0.8)[clear])
    assert 0 < gotf.sum() < len(gotf)

    # (4) trend features over the same window
    tr = buf.trend.cpu().numpy()
    t_i = where[len(where) // 2]
    k = where.index(t_i)
    w = np.array(wins[k], dtype=np.float64) * 100.0
    t, i = t_i
    if not dones[t, i]:
        pos_next = (obs[t + 1][i, :2] if t + 1 < T else last[i, :2]).astype(np.float64) * 500.0
        lab = pp.trend_label(w, pos_next, src0[i] if epi[t, i] == ep0[i] else frozen.source_pos.cpu().numpy()[i])
        if epi[t, i] == ep0[i]:
            assert np.isclose(tr[t, i, 0], lab[0], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("plume_model", ["isotropic", "dispersion"])
def test_fused_rollout_equals_discrete_kernels(plume_model):
    """Same seed, same Philox counters: the persistent kernel and the step-by-step kernels (policy
    act -> env step with auto-reset) produce identical transitions (code model and README model)."""
    N, T = 200, 60
    m, env_f, model, head, eng = _setup(N, T, seed=3, with_head=False, radius=40.0, plume_model=plume_model)
    buf = eng.collect()
    env_d = m.VecMethaneEnv(N, version="2.1", seed=3, field_mode="procedural", auto_reset=True,
                            plume_model=plume_model)
    env_d.curriculum[0] = 40.0
    obs = env_d.reset().clone()
    for t in range(T):
        assert torch.equal(buf.obs[t], obs), t
        a, lp, v, _ = model.act(obs, env=env_d)
        o, r, d, info = env_d.step(a)
        assert torch.equal(buf.actions[t], a), t
        assert torch.equal(buf.log_probs[t], lp) and torch.equal(buf.values[t], v)
        assert torch.equal(buf.rewards[t], r.float())
        assert torch.equal(buf.dones[t] != 0, d)
        obs = o.clone()
    assert torch.equal(eng.last_obs, obs)
    assert torch.equal(env_f.episode_idx, env_d.episode_idx)
    assert torch.equal(env_f.visited_t, env_d.visited_t)
    assert torch.equal(env_f.pos_x, env_d.pos_x) and torch.equal(env_f.step_count_t, env_d.step_count_t)


def test_rollout_segments_are_continuous():
    """Two collect() calls of T/2 == one of T (state, windows and counters persist)."""
    N, T = 64, 64
    m, env_a, model, head, eng_a = _setup(N, T, seed=9, radius=30.0)
    full = eng_a.collect()
    ra, sa = full.rewards.clone(), full.stop_prob.clone()
    m2, env_b, model_b, head_b, eng_b = _setup(N, T, seed=9, radius=30.0)
    first = eng_b.collect(horizon=T // 2)
    r1, s1 = first.rewards[: T // 2].clone(), first.stop_prob[: T // 2].clone()
    second = eng_b.collect(horizon=T // 2)
    assert torch.equal(ra[: T // 2], r1) and torch.equal(ra[T // 2:], second.rewards[: T // 2])
    assert torch.equal(sa[: T // 2], s1) and torch.equal(sa[T // 2:], second.stop_prob[: T // 2])


def test_forced_actions_noise_and_stop_termination():
    N, T = 64, 50
    m, env, model, head, eng = _setup(N, T, seed=21, radius=10.0)
    forced = torch.randint(0, 5, (T, N), dtype=torch.int32)
    noise = torch.randn(T, N, 2, dtype=torch.float64)
    used = torch.zeros(T, N, 2, dtype=torch.float64, device="cuda")
    probe = eng.collect(forced_actions=forced, step_noise=noise, noise_out=used)
    assert torch.equal(probe.actions.cpu(), forced) and torch.equal(used.cpu(), noise)
    sp = probe.stop_prob.cpu()
    med = float(sp[sp > 0].median())
    # shift the stop-head bias so that the median stop probability sits on the 0.8 threshold
    shift = float(np.log(0.8 / 0.2) - np.log(med / (1 - med)))
    m, env, model, head, eng = _setup(N, T, seed=21, radius=10.0)
    with torch.no_grad():
        head.fc_stop[0].bias.add_(shift)
    buf = eng.collect(forced_actions=forced, step_noise=noise, stop_terminates=True)
    stop = buf.stop_flag.cpu() != 0
    done = buf.dones.cpu() != 0
    assert stop.any() and not stop.all() and torch.all(done[stop])      # a stop decision ends the episode
    assert (done & ~stop & ~(buf.reached.cpu() != 0)).sum() == 0        # nothing else ends one here
    # greedy rollouts are deterministic given the state
    m2, env2, model2, head2, eng2 = _setup(N, T, seed=21, radius=10.0)
    g = eng2.collect(greedy=True)
    p, _ = model2(g.obs.reshape(-1, 6))
    assert torch.equal(g.actions.reshape(-1).long(), p.argmax(-1))


def test_baseline_shape_rollout_sample_through_oracle():
    """BASELINE shape (4096 envs x 256 steps, radius 50, unscaled actor): 64 envs sampled across the tiles are
    replayed through the oracle on the kernel's own actions and draws -- flags and indices bit-exact, rewards /
    observations within fp32 rel 1e-5 -- and the policy outputs of ALL 1 M transitions are checked against the
    oracle MLP."""
    N, T = 4096, 256
    m, env, model, head, eng = _setup(N, T, seed=5, radius=None, actor_gain=1.0)
    cfg = po.config_for("2.1")
    rng = np.random.default_rng(0)
    ids = np.sort(rng.choice(N, size=64, replace=False)).astype(np.int64)
    ids[0], ids[-1] = 0, N - 1                      # first and last env of the grid
    src0 = env.source_pos.cpu().numpy().copy()
    ep0 = env.episode_idx.cpu().numpy().copy()
    noise = torch.zeros(T, N, 2, dtype=torch.float64, device="cuda")
    buf = eng.collect(noise_out=noise)
    eng.check_nan()
    obs = buf.obs.cpu().numpy()
    acts = buf.actions.cpu().numpy()
    rew = buf.rewards.cpu().numpy()
    dones = buf.dones.cpu().numpy() != 0
    reached = buf.reached.cpu().numpy() != 0
    epi = buf.episode_idx.cpu().numpy()
    info = buf.info.cpu().numpy()
    zs = noise.cpu().numpy()

    frozen = m.VecMethaneEnv(N, version="2.1", seed=5, field_mode="procedural")
    assert np.array_equal(frozen.episode_idx.cpu().numpy(), ep0)

    def noise_cb(idx, x, y):
        z, u = frozen.field_noise_at(ids[np.asarray(idx)].astype(np.int32), np.asarray(x, dtype=np.int32),
                                     np.asarray(y, dtype=np.int32))
        return z.cpu().numpy(), u.cpu().numpy()

    n = len(ids)
    ora = po.OracleVecEnv(cfg, n, fields=po.CellNoiseFields(cfg, n, noise_cb))
    for i in range(n):
        ora.set_source(i, src0[ids[i]])
    alive = np.ones(n, dtype=bool)
    assert np.allclose(ora.observe(), obs[0][ids], rtol=2e-7, atol=1e-9)
    checked = 0
    for t in range(T):
        o, r, d, inf = ora.step(acts[t][ids], zs[t][ids])
        assert np.array_equal(d[alive], dones[t][ids][alive]), t
        assert np.array_equal(inf["reached"][alive], reached[t][ids][alive]), t
        assert np.allclose(r[alive], rew[t][ids][alive], rtol=1e-5, atol=1e-6), t
        for k, key in enumerate(("concentration_reward", "explore_reward", "move_penalty", "tke_penalty",
                                 "boundary_penalty")):
            assert np.allclose(np.asarray(inf[key], dtype=np.float32)[alive], info[t, k][ids][alive], rtol=1e-5,
                               atol=1e-7), (t, key)
        assert np.all(epi[t][ids][alive] == ep0[ids][alive])
        checked += int(alive.sum())
        if t + 1 < T:
            cont = alive & ~d
            assert np.array_equal(o[cont][:, [0, 1, 4, 5]], obs[t + 1][ids][cont][:, [0, 1, 4, 5]]), t
            assert np.allclose(o[cont], obs[t + 1][ids][cont], rtol=2e-7, atol=1e-9), t
        alive &= ~d
    assert checked > 0.6 * n * T and 0 < (~alive).sum() < n        # most transitions replayed; some episodes ended

    ora_m = pp.OracleActorCritic()
    ora_m.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    with torch.no_grad():
        p, v = ora_m(torch.from_numpy(obs.reshape(-1, 6)))
    lp = pp.categorical_log_prob(p, torch.from_numpy(acts.reshape(-1)).long())
    assert torch.allclose(buf.values.cpu().reshape(-1), v.squeeze(-1), rtol=1e-5, atol=2e-6)
    assert torch.allclose(buf.log_probs.cpu().reshape(-1), lp, rtol=1e-5, atol=2e-6)


def test_full_size_invariants():
    """BASELINE size (4096 envs): size-independent properties of the rollout."""
    N, T = 4096, 64
    m, env, model, head, eng = _setup(N, T, seed=2, radius=50.0, actor_gain=1.0)
    ep0 = env.episode_idx.clone()
    buf = eng.collect()
    eng.check_nan()
    done = buf.dones != 0
    epi = buf.episode_idx
    assert torch.equal(env.episode_idx - ep0, done.sum(0).int())              # one reset per done
    assert torch.equal(epi[1:] - epi[:-1], done[:-1].int())                    # episode index steps at dones
    step_obs = buf.obs[:, :, 4] * 1000
    nxt = torch.where(done[:-1], torch.zeros_like(step_obs[1:]), step_obs[:-1] + 1)
    assert torch.allclose(step_obs[1:], nxt, atol=1e-3)                        # step counter in the observation
    assert torch.all(buf.reached[done] == 1)                                  # nothing times out in 64 steps
    assert torch.isfinite(buf.rewards).all() and torch.isfinite(buf.values).all()
    assert torch.all((buf.actions >= 0) & (buf.actions < 5))
    assert torch.all(buf.obs[:, :, 0] >= 0) and torch.all(buf.obs[:, :, 0] <= 1)
    assert float(buf.log_probs.max()) <= 0


@pytest.mark.parametrize("path", ["cuda", "tc"])
@pytest.mark.parametrize("splits", [(64,), (7, 13, 44), (32, 32)])
def test_deferred_stop_head_equals_in_loop_head(splits, path):
    """plume_stop_head_segment (the head taken out of the lockstep loop) reproduces the in-loop head:
    stop probability, flag, peak, trend features, and the window ring carried from one segment to the
    next (also for segments shorter than the window).  The CUDA-core kernel is bit-identical to the
    in-loop head; the tensor-core kernel (3xTF32 gate GEMM, ex2/rcp activations) agrees to fp32 rel 1e-5,
    flags equal wherever the probability is not within 1e-5 of the threshold."""
    N, T = 80, 64
    m, env_a, model_a, head_a, eng_a = _setup(N, T, seed=17, radius=25.0)
    m, env_b, model_b, head_b, eng_b = _setup(N, T, seed=17, radius=25.0)
    eng_b.stop_head_path = m._lib.KERNEL_PATHS[path]
    ref = eng_a.collect(defer_stop_head=False)
    sp, sf, pk, tr = ref.stop_prob.clone(), ref.stop_flag.clone(), ref.peak_pred.clone(), ref.trend.clone()
    assert (sp > 0).any() and sf.any() and not sf.all()
    t0 = 0
    for h in splits:
        seg = eng_b.collect(horizon=h)                    # default: deferred
        assert torch.equal(seg.rewards[:h], ref.rewards[t0:t0 + h])
        if path == "cuda":
            assert torch.equal(seg.stop_prob[:h], sp[t0:t0 + h])
            assert torch.equal(seg.stop_flag[:h], sf[t0:t0 + h])
            assert torch.equal(seg.peak_pred[:h], pk[t0:t0 + h])
        else:
            assert torch.equal(seg.stop_prob[:h] == 0, sp[t0:t0 + h] == 0)          # same windows evaluated
            assert torch.allclose(seg.stop_prob[:h], sp[t0:t0 + h], rtol=1e-5, atol=1e-6)
            assert torch.allclose(seg.peak_pred[:h], pk[t0:t0 + h], rtol=1e-5, atol=1e-6)
            clear = (sp[t0:t0 + h] - 0.8).abs() > 1e-5
            assert torch.equal(seg.stop_flag[:h][clear], sf[t0:t0 + h][clear])
        assert torch.equal(seg.trend[:h], tr[t0:t0 + h])
        t0 += h
    assert torch.equal(eng_a.window_fill, eng_b.window_fill)
    full = eng_a.window_fill >= eng_a.window
    assert full.any() and torch.equal(eng_a.conc_window[full], eng_b.conc_window[full])
    # mixing the modes keeps the carried state consistent
    nxt_a = eng_a.collect(horizon=8, defer_stop_head=False).stop_prob[:8].clone()
    nxt_b = eng_b.collect(horizon=8, defer_stop_head=False).stop_prob[:8].clone()
    assert torch.equal(nxt_a, nxt_b)


@pytest.mark.parametrize("hidden,path", [(48, "auto"), (64, "tensor"), (64, "simt"), (128, "auto"), (256, "auto")])
def test_deferred_stop_head_other_hidden_sizes(hidden, path):
    """BASELINE configs[4] sweeps the stop-head hidden size up to 256: 64 has a resident-weight tcgen05 kernel (and the
    CUDA-core one), the others run on the kernels the library has for them.  Checked against the torch-CPU LSTM
    (the oracle, bit-pinned to the reference's PeakAndStopPredictor) on the windows rebuilt from the segment's samples
    (also across a segment boundary: the carried ring), trend features against the 32-unit run.  N = 300 envs: three
    128-window tiles per step, the last one ragged."""
    import uav_wrf_les_ppo_lstm_b200 as m
    from oracle import ppo_oracle as pp
    N, splits, W = 300, (30, 34), 20
    torch.manual_seed(hidden)
    env = m.VecMethaneEnv(N, version="2.1", seed=17, field_mode="procedural", auto_reset=True)
    env.curriculum[0] = 25.0
    env.reset()
    model = m.PPOActorCritic(device="cuda")
    head = m.PeakAndStopPredictor(hidden_dim=hidden, device="cuda")
    ora = pp.OraclePeakAndStop(hidden_dim=hidden)
    ora.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    eng = m.RolloutEngine(env, model, head, horizon=max(splits), with_trend=True, stop_head_path=path)
    # the same rollout with the 32-unit head gives the reference trend features (they do not depend on the head)
    env2 = m.VecMethaneEnv(N, version="2.1", seed=17, field_mode="procedural", auto_reset=True)
    env2.curriculum[0] = 25.0
    env2.reset()
    eng2 = m.RolloutEngine(env2, model, m.PeakAndStopPredictor(device="cuda"), horizon=max(splits), with_trend=True)
    samples, fills, probs, peaks, flags = [], [], [], [], []
    for h in splits:
        seg, seg2 = eng.collect(horizon=h), eng2.collect(horizon=h)
        assert torch.equal(seg.rewards[:h], seg2.rewards[:h]) and torch.equal(seg.trend[:h], seg2.trend[:h])
        samples.append(seg.conc_sample[:h].clone()); fills.append(seg.fill_t[:h].clone())
        probs.append(seg.stop_prob[:h].clone()); peaks.append(seg.peak_pred[:h].clone()); flags.append(seg.stop_flag[:h].clone())
    cs, fill = torch.cat(samples).cpu(), torch.cat(fills).cpu()
    sp, pk, sf = torch.cat(probs).cpu(), torch.cat(peaks).cpu(), torch.cat(flags).cpu()
    full = fill >= W
    assert full.any() and (~full).any()
    assert torch.all(sp[~full] == 0) and torch.all(pk[~full] == 0) and not sf[~full].any()
    tt, nn = full.nonzero(as_tuple=True)
    windows = torch.stack([cs[t - W + 1:t + 1, n] for t, n in zip(tt.tolist(), nn.tolist())])
    assert (tt >= splits[0]).any()                       # some windows straddle the segment boundary
    with torch.no_grad():
        rp, rs = ora(windows)
    assert torch.allclose(pk[full], rp, rtol=1e-5, atol=2e-6)
    assert torch.allclose(sp[full], rs, rtol=1e-5, atol=2e-6)
    clear = (rs - 0.8).abs() > 1e-5
    assert torch.equal(sf[full][clear].bool(), (rs > 0.8)[clear])


@pytest.mark.parametrize("chunks", [(16, 48), (8, 8, 48), (5, 59)])
def test_overlapped_collection_changes_nothing(chunks):
    """RolloutEngine.overlap_chunks: the segment launched as shorter lockstep segments on one stream with the stop head of
    each on another (the head of the first rows under the lockstep kernel of the next) writes the same buffer and
    carries the same state as the single launch -- every array bit for bit, over two consecutive segments."""
    N, T = 80, 64
    m, env_a, model_a, head_a, eng_a = _setup(N, T, seed=23, radius=25.0)
    m, env_b, model_b, head_b, eng_b = _setup(N, T, seed=23, radius=25.0)
    eng_b.overlap_chunks = chunks
    for it in range(2):
        a, b = eng_a.collect(), eng_b.collect()
        torch.cuda.synchronize()
        assert eng_b.launches == (it + 1) * 2 * len(chunks) and eng_a.launches == (it + 1) * 2
        for name in ("obs", "actions", "rewards", "values", "log_probs", "dones", "reached", "flag_code", "stop_prob",
                     "stop_flag", "peak_pred", "trend", "info", "episode_idx", "conc_sample", "fill_t", "src_dist"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (it, name)
        assert torch.equal(eng_a.window_fill, eng_b.window_fill) and torch.equal(eng_a.last_obs, eng_b.last_obs)
        assert torch.equal(eng_a.conc_window, eng_b.conc_window)
    assert (a.stop_prob > 0).any() and a.stop_flag.any()
    # a horizon the chunks do not sum to falls back to the single launch
    eng_b.collect(horizon=40)
    assert eng_b.launches == 4 * len(chunks) + 2


def test_trainer_overlaps_the_first_rows_of_the_stop_head():
    """PlumeTrainer enables the two-launch split where the lockstep kernel leaves SMs idle (rollout.overlap_split); the
    rollout of its first iteration equals the single-launch trainer's."""
    import uav_wrf_les_ppo_lstm_b200 as m
    ta = m.PlumeTrainer(num_envs=1024, horizon=256, seed=5, overlap_stop_head=False)
    tb = m.PlumeTrainer(num_envs=1024, horizon=256, seed=5)
    assert ta.engine.overlap_chunks is None and tb.engine.overlap_chunks is not None
    assert sum(tb.engine.overlap_chunks) == 256 and len(tb.engine.overlap_chunks) == 2
    ta.train_iteration(); tb.train_iteration()
    torch.cuda.synchronize()
    a, b = ta.engine.buffer, tb.engine.buffer
    for name in ("obs", "actions", "rewards", "values", "log_probs", "dones", "stop_prob", "stop_flag", "peak_pred", "trend"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert tb.launches_per_iteration == ta.launches_per_iteration + 2
    # a collection may be left pending behind the lockstep kernels and joined later (off in the trainer: no gain)
    assert not tb.defer_head_join and not tb.engine._head_pending
    seg = tb.engine.collect(join=False)
    assert tb.engine._head_pending
    tb.engine.join_stop_head()
    assert not tb.engine._head_pending
    ref = ta.engine.collect()
    torch.cuda.synchronize()
    assert torch.allclose(ta.model.flat, tb.model.flat, atol=1e-5)


def test_deferred_stop_head_rejects_terminating_stop():
    N, T = 32, 8
    m, env, model, head, eng = _setup(N, T, seed=1)
    with pytest.raises(ValueError):
        eng.collect(stop_terminates=True, defer_stop_head=True)
