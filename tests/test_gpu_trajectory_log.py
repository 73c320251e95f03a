"""N2: trajectory / per-episode logging in the training_data.nc and training_results.csv layouts
(PPOV2.1/model.py:351-419, nc_info.txt, train_ppo2.0.py:128-134,166-248), assembled on the device from the
[T, N] rollout buffers; checked against a plain per-step Python replay of the same buffers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trajectory_logger_matches_per_step_replay(tmp_path):
    import uav_wrf_les_ppo_lstm_b200 as m
    torch.manual_seed(0)
    N, T, segs = 64, 40, 4
    env = m.VecMethaneEnv(N, version="2.1", seed=4, field_mode="procedural", auto_reset=True)
    env.curriculum[0] = 140.0
    env.reset()
    model = m.PPOActorCritic(device="cuda")
    with torch.no_grad():
        model.actor.weight.mul_(40.0)
    head = m.PeakAndStopPredictor(device="cuda")
    eng = m.RolloutEngine(env, model, head, horizon=T, with_info=True, with_trajectory=True)
    log = m.TrajectoryLogger(env, max_episodes=500)
    # replay state (what the reference driver keeps per episode, train_ppo2.0.py:140-180)
    cur = [dict(x=[], y=[], conc=[], sums=np.zeros(6)) for _ in range(N)]
    episodes = []
    for _ in range(segs):
        buf = eng.collect()
        got = log.consume(buf)
        pos = buf.pos_out.cpu().numpy()
        conc = buf.conc_sample.cpu().numpy() * 100.0
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
                    episodes.append(dict(c, src=src[t, n].copy(), success=bool(reached[t, n])))
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
        assert nc["source_x"][i] == e["src"][0] and 50 <= nc["source_x"][i] <= 450
    assert np.all(nc["gaussian_sigma"] == 15.0) and np.all(nc["peak_concentration"] == 100.0)
    log.save(str(tmp_path / "training_data_like.npz"), str(tmp_path / "training_results.csv"))
    back = np.load(tmp_path / "training_data_like.npz")
    assert np.array_equal(back["is_source"], nc["is_source"])
    hdr = open(tmp_path / "training_results.csv").readline().strip().split(",")
    assert hdr == ["Episode", "Total_Reward", "Success", "Conc_Reward", "Explore_Reward", "Move_Penalty", "TKE_Penalty",
                   "Boundary_Penalty", "Steps", "Final_Conc", "Current_Radius"]              # train_ppo2.0.py:128-134
