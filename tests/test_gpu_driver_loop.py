"""The N = 1 drop-in contract (north_star: "a drop-in for the train_ppo*.py drivers"): a
transcription of the reference driver's loop body (PPOV2.1/train_ppo2.0.py:137-199,222-251) runs on
``pb.MethaneEnv`` / ``pb.PPOActorCritic`` / ``pb.PPOBuffer`` / ``pb.update_model`` / ``pb.PPOTrainer`` /
``pb.TrajectoryLogger`` and must reproduce what the UNMODIFIED reference produced for the same seed
(tests/golden/driver_loop_s45.npz, written by oracle/make_golden.py::driver_loop_trace from the reference's own
classes incl. ``NetCDFWriter`` and the CSV rows).  The env draws are rebuilt from ``RandomState(seed)`` in the
reference's order; the sampled actions and the ``torch.randperm`` draws are replayed from the fixture (they come
from torch's global generator, which the product does not share)."""
import numpy as np
import pytest
import torch

from oracle import plume_oracle as po
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def test_reference_driver_loop_runs_on_the_dropins():
    import uav_wrf_les_ppo_lstm_b200 as pb
    g = load_golden("driver_loop_s45.npz")
    cfg = po.config_for("2.1")
    G, BATCH_SIZE = cfg.grid_size, 256
    rs = np.random.RandomState(int(g["seed"]))

    def reset_draws():
        return rs.rand(2), rs.randn(G, G), rs.rand(G, G)

    reset_draws()                                        # MethaneEnv() itself resets once (environment.py:40)
    env = pb.MethaneEnv(version="2.1")
    model = pb.PPOActorCritic(6, 5)
    model.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init.")})
    optimizer = pb.FusedAdam(model, lr=3e-5)
    buffer = pb.PPOBuffer(horizon=BATCH_SIZE, num_envs=1)
    trainer = pb.PPOTrainer(env, model, optimizer)
    # what the reference keeps in episode_data / hands to NetCDFWriter: one logging row per step
    logbuf = pb.PPOBuffer(horizon=cfg.max_steps, num_envs=1, with_info=True, with_trajectory=True, with_stop=False)
    log = pb.TrajectoryLogger(env, max_episodes=int(g["episodes"]))
    i, n_updates = 0, 0
    perms = g["perms"].astype(np.int64)
    for episode in range(int(g["episodes"])):
        u, z, uf = reset_draws()
        conc, tke = po.plume_fields(cfg, u * (G - 100) + 50, z, uf)
        state = env.reset(u_src=u[None], fields=(conc[None], tke[None]))
        assert np.array_equal(env.source_pos, g["sources"][episode])
        done = False
        success = False
        while not done:
            assert np.array_equal(state, g["state"][i]), (episode, i)              # bit-exact observations
            state_t = torch.FloatTensor(state).unsqueeze(0).cuda()
            with torch.no_grad():
                probs, value = model(state_t)
            action = int(g["action"][i])                 # the reference's Categorical.sample() draw, replayed
            logp = torch.log(probs[0, action] / probs[0].sum()).item()
            # fp32 rel 1e-5 while both sides hold the same parameters; after an _update_model the two parameter sets
            # differ by what Adam makes of fp32 summation-order noise (a fraction of lr per step on entries whose
            # gradient nearly cancels: the yardstick of tests/test_gpu_learner.py), which moves the outputs by ~1e-5
            tol = 1e-5 if n_updates == 0 else 2e-4
            assert abs(value.item() - g["value"][i]) <= tol * max(1.0, abs(g["value"][i])), (i, value.item())
            assert abs(logp - g["logp"][i]) <= tol * max(1.0, abs(g["logp"][i])), (i, logp)
            next_state, reward, done, info = env.step(action, step_noise=rs.randn(2))
            assert reward == g["reward"][i] and done == bool(g["done"][i]), (i, reward)   # float64 reward, bit-exact
            x, y = env.agent_pos
            current_conc = env.conc_field[np.clip(int(x), 0, G - 1), np.clip(int(y), 0, G - 1)]
            assert np.array_equal(env.agent_pos, g["pos"][i]) and current_conc == g["conc"][i]
            buffer.store(state, action, reward, value.item(), logp, done)
            logbuf.store(state, action, reward, value.item(), logp, done, reached=env.trajectory[-1]["reached"],
                         info=info, pos=env.agent_pos, conc=current_conc, src=env.source_pos)
            if len(buffer.states) >= BATCH_SIZE:
                assert bool(g["update_after"][i])
                pb._update_model(buffer, model, optimizer, perms=list(perms[5 * n_updates:5 * n_updates + 5]))
                buffer.clear()
                n_updates += 1
            else:
                assert not bool(g["update_after"][i])
            state = next_state
            i += 1
        if env.trajectory[-1]["reached"]:
            success = True
        # the reference logs trainer.current_radius BEFORE this episode's curriculum update (:247,251)
        assert log.consume(logbuf, radius=trainer.current_radius) == 1
        logbuf.clear()
        trainer.update(success)
        cur = g["curriculum"][episode]
        assert (trainer.current_radius, trainer.explore_bonus) == (cur[0], cur[1])
        assert (env.current_radius, env.explore_bonus) == (cur[2], cur[3])
    assert i == len(g["action"]) and n_updates == int(g["n_updates"])

    # --- training_results CSV (train_ppo2.0.py:236-248) and training_data.nc (model.py:405-419) -------------------
    rows, want = log.csv_rows(), g["csv"]
    assert rows.shape == want.shape
    assert np.array_equal(rows[:, [0, 2, 8]], want[:, [0, 2, 8]])                  # episode, success, steps
    assert np.array_equal(rows[:, 10], want[:, 10])                                # Current_Radius
    assert np.array_equal(rows[:, 9].astype(np.float32), want[:, 9].astype(np.float32))   # Final_Conc: 0.0 / conc at the final cell
    assert (want[:, 9] > 0).any() and (want[:, 9] == 0).any()
    assert np.allclose(rows[:, [1, 3, 4, 5, 6, 7]], want[:, [1, 3, 4, 5, 6, 7]], rtol=1e-5, atol=1e-6)
    nc = log.nc_variables()
    for name in ("x", "y", "concentration"):
        assert np.array_equal(nc[name], g["nc_" + name], equal_nan=True), name
    # (the reference never writes the coordinate variables 'episode' / 'step': model.py:365-369 only creates them)
    for name in ("is_source", "source_x", "source_y", "source_concentration", "gaussian_sigma", "peak_concentration"):
        assert np.array_equal(nc[name], g["nc_" + name]), name

    # --- the four _update_model calls moved the parameters like the reference's did -------------------------------
    num = den = 0.0
    for k, v in model.state_dict().items():
        init, final = torch.from_numpy(g["init." + k]), torch.from_numpy(g["final." + k])
        assert (v.cpu() - final).abs().max().item() <= 10 * 3e-5, k      # 20 Adam steps of at most lr each
        d_ref, d_gpu = (final - init).flatten().double(), (v.cpu() - init).flatten().double()
        num += float((d_ref * d_gpu).sum())
        den += float(d_ref.norm() * d_gpu.norm())
        assert float(d_ref.norm()) > 0
    assert num / den > 0.98, num / den


def test_methane_env_caches_the_field_per_episode():
    """``env.conc_field`` is read every step by the reference driver (train_ppo2.0.py:167-170): one download per
    episode, refreshed by reset()."""
    import uav_wrf_les_ppo_lstm_b200 as pb
    env = pb.MethaneEnv(version="2.1", seed=3)
    a = env.conc_field
    assert env.conc_field is a and a.shape == (500, 500) and a.dtype == np.float64
    launches = env.launches
    for _ in range(5):
        env.step(1)
        assert env.conc_field is a
    assert env.launches == launches + 5                  # only the step kernels
    env.reset()
    b = env.conc_field
    assert b is not a and not np.array_equal(a, b)
    sx, sy = env.source_pos
    assert b[int(sx), int(sy)] > 90.0                    # the plume sits on the new source


def test_update_model_rejects_foreign_optimisers():
    import uav_wrf_les_ppo_lstm_b200 as pb
    model = pb.PPOActorCritic(device="cuda")
    buf = pb.PPOBuffer(8, 1, "cuda")
    for _ in range(8):
        buf.store(np.zeros((1, 6), np.float32), [1], [0.5], [0.1], [-1.6], [0.0])
    with pytest.raises(TypeError):
        pb.update_model(buf, model, torch.optim.Adam(model.parameters(), lr=3e-5))
    other = pb.PPOActorCritic(device="cuda")
    with pytest.raises(TypeError):
        pb.update_model(buf, model, pb.FusedAdam(other))
