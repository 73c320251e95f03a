"""N2: trajectory / per-episode logging in the training_data.nc and training_results.csv layouts
(PPOV2.1/model.py:351-419, nc_info.txt, train_ppo2.0.py:128-134,166-248), assembled on the device from the
[T, N] rollout buffers by plume_trajectory_log; checked against a plain per-step Python replay of the same buffers
(the reference driver's bookkeeping, train_ppo2.0.py:140-199,236-251, applied per env in canonical episode order).
The writer itself is pinned against the reference's NetCDFWriter / CSV in tests/test_gpu_driver_loop.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trajectory_logger_matches_per_step_replay(tmp_path):
    import uav_wrf_les_ppo_lstm_b200 as m
    torch.manual_seed(0)
    N, T, segs = 64, 40, 4
    import dataclasses
    # a 16-episode curriculum window (reference: 120) so that the radius changes several times inside the test
    env = m.VecMethaneEnv(N, version="2.1", seed=4, field_mode="procedural", auto_reset=True,
                          config=dataclasses.replace(m.config_for("2.1"), window_size=16, initial_radius=140.0))
    env.curriculum[0] = 140.0
    env.reset()
    model = m.PPOActorCritic(device="cuda")
    with torch.no_grad():
        model.actor.weight.mul_(40.0)
    head = m.PeakAndStopPredictor(device="cuda")
    eng = m.RolloutEngine(env, model, head, horizon=T, with_info=True, with_trajectory=True)
    log = m.TrajectoryLogger(env, max_episodes=500)
    trainer = m.PPOTrainer(env, model, None)
    trainer.device_state()[0] = 140.0            # trainer and env start from the same radius
    trainer.device_state()[2] = 140.0
    from oracle import ppo_oracle as pp
    from oracle import plume_oracle as po

    class _E:
        current_radius = 140.0
        explore_bonus = 0.6
    ora = pp.OracleCurriculum(_E(), dataclasses.replace(po.config_for("2.1"), window_size=16, initial_radius=140.0))
    ora.current_radius = 140.0
    # replay state (what the reference driver keeps per episode, train_ppo2.0.py:140-180)
    cur = [dict(x=[], y=[], conc=[], sums=np.zeros(6)) for _ in range(N)]
    episodes = []
    for _ in range(segs):
        buf = eng.collect()
        trainer.update_from_rollout(buf)             # the logger takes every episode's radius from this replay
        got = log.consume(buf, trainer=trainer)
        pos = buf.pos_out.cpu().numpy()
        conc = buf.conc_out.cpu().numpy()
        assert np.allclose(conc, buf.conc_sample.cpu().numpy() * 100.0, rtol=1e-6, atol=1e-6)
        rew = buf.rewards.cpu().numpy().astype(np.float64)
        info = buf.info.cpu().numpy().astype(np.float64)          # [T, 5, N]
        done = buf.dones.cpu().numpy() != 0
        reached = buf.reached.cpu().numpy() != 0
        src = buf.src_out.cpu().numpy()
        n_before = len(episodes)
        for t in range(T):
            for n in range(N):
                c = cur[n]
                c["x"].append(pos[t, n, 0]); c["y"].append(pos[t, n, 1]); c["conc"].append(np.float32(conc[t, n]))
                c["sums"] += np.concatenate([[rew[t, n]], info[t, :, n]])
                if done[t, n]:
                    # train_ppo2.0.py:194-196,246-251: final concentration only on success; the trainer's radius is
                    # logged BEFORE this episode's curriculum update
                    episodes.append(dict(c, src=src[t, n].copy(), success=bool(reached[t, n]),
                                         final_conc=np.float32(conc[t, n]) if reached[t, n] else np.float32(0.0),
                                         radius=ora.current_radius))
                    ora.update(bool(reached[t, n]))
                    cur[n] = dict(x=[], y=[], conc=[], sums=np.zeros(6))
        assert got == len(episodes) - n_before
    assert log.count == len(episodes) and log.count > 20
    assert any(len(e["x"]) > T for e in episodes), "no episode spanned two segments: the carry path is untested"
    nc = log.nc_variables()
    rows = log.csv_rows()
    assert set(nc) == {"episode", "step", "x", "y", "concentration", "is_source", "source_concentration", "source_x",
                       "source_y", "gaussian_sigma", "peak_concentration"}                     # nc_info.txt
    assert nc["x"].shape == (log.count, 1000) and nc["x"].dtype == np.float32 and nc["is_source"].dtype == np.int8
    for i, e in enumerate(episodes):
        L = len(e["x"])
        assert rows[i, 8] == L and rows[i, 0] == i + 1 and rows[i, 2] == int(e["success"])
        x = np.array(e["x"], dtype=np.float32)
        x[-1] = e["src"][0]                                       # model.py:410-412: last step <- source
        y = np.array(e["y"], dtype=np.float32)
        y[-1] = e["src"][1]
        assert np.array_equal(nc["x"][i, :L], x) and np.isnan(nc["x"][i, L:]).all()
        assert np.array_equal(nc["y"][i, :L], y)
        assert np.array_equal(nc["concentration"][i, :L], np.array(e["conc"], dtype=np.float32))
        assert nc["is_source"][i].sum() == 1 and nc["is_source"][i, L - 1] == 1
        assert np.allclose(rows[i, [1, 3, 4, 5, 6, 7]], e["sums"], rtol=1e-9, atol=1e-9)
        assert np.float32(rows[i, 9]) == e["final_conc"]                        # Final_Conc
        assert np.isclose(rows[i, 10], e["radius"], rtol=1e-12), (i, rows[i, 10], e["radius"])   # Current_Radius
        assert nc["source_x"][i] == e["src"][0] and 50 <= nc["source_x"][i] <= 450
    assert np.all(nc["gaussian_sigma"] == 15.0) and np.all(nc["peak_concentration"] == 100.0)
    radii = rows[:, 10]
    assert len(np.unique(radii)) > 1, "the radius never changed: the per-episode radius lookup is untested"
    assert (rows[:, 9] > 0).any()          # (Final_Conc = 0.0 of a failed episode: tests/test_gpu_driver_loop.py)
    log.save(str(tmp_path / "training_data_like.npz"), str(tmp_path / "training_results.csv"))
    back = np.load(tmp_path / "training_data_like.npz")
    assert np.array_equal(back["is_source"], nc["is_source"])
    hdr = open(tmp_path / "training_results.csv").readline().strip().split(",")
    assert hdr == ["Episode", "Total_Reward", "Success", "Conc_Reward", "Explore_Reward", "Move_Penalty", "TKE_Penalty",
                   "Boundary_Penalty", "Steps", "Final_Conc", "Current_Radius"]              # train_ppo2.0.py:128-134
