"""One PPO iteration at the bench configuration (4096 envs x 256 steps, V2.1 + LSTM stop head),
for ncu: `python profiles/profile_step.py` (plain) then the same command under ncu.
Also launches the field-generation kernel K1 (1024 envs, float fields) once."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    tr = pb.PlumeTrainer(num_envs=n, horizon=256, minibatch_size=(n * 256) // 4)
    tr.train_iteration()          # warm-up (launches 1..64)
    torch.cuda.synchronize()
    tr.train_iteration()          # profiled iteration
    torch.cuda.synchronize()
    env = pb.VecMethaneEnv(1024, field_mode="f32", seed=1)       # constructor launches K1 once
    lib = pb._lib.load()
    pb._lib.check(lib.plume_generate_fields(C.byref(env.c_config), C.byref(env.c_state), None, 1024, None, None,
                                            torch.cuda.current_stream().cuda_stream), "generate")
    acts = torch.zeros(1024, dtype=torch.int32, device="cuda")
    env.step(acts)
    torch.cuda.synchronize()
    print("profile_step ok", float(tr.last_losses[0, 0]))


if __name__ == "__main__":
    main()
