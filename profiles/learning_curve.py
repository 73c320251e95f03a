"""Does the batched trainer learn?  Trains PPOV2.1 with the REFERENCE schedule (an update of 5 epochs x one 256-sample
minibatch every 256 transitions, lr 3e-5, config.py) for `--episodes` episodes and prints what the reference's
training_results2_0.csv records: success rate and the curriculum radius over the run (reference, 2000 episodes:
63.6 % success, radius 50 -> 8.28).
    python profiles/learning_curve.py --envs 1 --horizon 256 --episodes 2000"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402


def run(envs: int, horizon: int, episodes: int, seed: int = 0, minibatch: int = 256):
    tr = pb.PlumeTrainer(num_envs=envs, horizon=horizon, version="2.1", seed=seed, minibatch_size=minibatch,
                         stop_head=False)
    hist = []
    it = 0
    while True:
        tr.train_iteration()
        it += 1
        if it % 200 == 0 or it == 1:
            st = tr.curriculum.sync_from_device()
            hist.append({"iteration": it, **st})
            if st["episodes"] >= episodes:
                break
    st = tr.curriculum.sync_from_device()
    return {"envs": envs, "horizon": horizon, "minibatch": minibatch, "iterations": it, "episodes": st["episodes"],
            "successes": st["successes"], "success_rate": st["successes"] / max(st["episodes"], 1),
            "final_radius": st["radius"], "final_explore_bonus": st["explore_bonus"],
            "env_steps": it * envs * horizon, "history": hist}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--horizon", type=int, default=256)
    ap.add_argument("--episodes", type=int, default=2000)
    ap.add_argument("--seeds", type=int, nargs="+", default=[0])
    a = ap.parse_args()
    for s in a.seeds:
        import time
        t0 = time.time()
        r = run(a.envs, a.horizon, a.episodes, s)
        r["seconds"] = time.time() - t0
        r["seed"] = s
        h = r.pop("history")
        r["radius_curve"] = [round(x["radius"], 2) for x in h][:: max(1, len(h) // 12)]
        print(json.dumps(r), flush=True)
