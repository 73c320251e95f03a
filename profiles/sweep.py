"""BASELINE.json configs[4]: scaling sweep over envs/GPU (256 ... 65536) and the stop-head hidden size, one GPU.
Prints one JSON line per point: rollout (lockstep loop + deferred stop head) and whole-iteration env-steps/s.
    python profiles/sweep.py [--hidden 32 64] [--envs 256 1024 4096 16384 65536] [--horizon 256]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hidden", type=int, nargs="+", default=[32, 64, 256])
    ap.add_argument("--max-envs-generic", type=int, default=4096,
                    help="largest envs/GPU for hidden sizes that run on the generic (L2-weights) LSTM kernel")
    ap.add_argument("--envs", type=int, nargs="+", default=[256, 1024, 4096, 16384, 65536])
    ap.add_argument("--horizon", type=int, default=256)
    args = ap.parse_args()
    T = args.horizon
    for H in args.hidden:
        for N in args.envs:
            if H not in (32, 64, 128, 256) and N > args.max_envs_generic:
                continue
            tr = pb.PlumeTrainer(num_envs=N, horizon=T, minibatch_size=max(N * T // 4, 256), seed=1)
            if H != 32:
                tr.head = pb.PeakAndStopPredictor(hidden_dim=H, device=tr.device)
                tr.engine.stop_head = tr.head
            for _ in range(2):
                tr.train_iteration()
            torch.cuda.synchronize()
            roll = timed(lambda: tr.engine.collect())
            full = timed(lambda: tr.train_iteration())
            print(json.dumps({"envs_per_gpu": N, "horizon": T, "lstm_hidden": H,
                              "stop_head": ("tcgen05, resident weights" if H in (32, 64) else
                                            ("tcgen05, weights streamed from L2 (TMA)" if H in (128, 256) else
                                             "cuda-core, L2 weights")),
                              "rollout_ms": round(roll, 3), "rollout_env_steps_per_s": N * T / roll * 1e3,
                              "us_per_lockstep_iteration": round(1e3 * roll / T, 2),
                              "iteration_ms": round(full, 3), "ppo_env_steps_per_s": N * T / full * 1e3}), flush=True)
            del tr
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
