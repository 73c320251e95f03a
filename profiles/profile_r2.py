"""Round-2 profiling driver: launches ONE kernel family per invocation so that `ncu -k regex:<kernel> -c 1` (after the
script's own warm-up, skipped with -s) captures it.
    python profiles/profile_r2.py k1|k2|iteration|head64|head128|head256|eval_fixed"""
import ctypes as C
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "iteration"
dev = torch.device("cuda:0")
lib = pb._lib.load()
if what == "k1":
    env = pb.VecMethaneEnv(1024, device=dev, field_mode="f32", seed=1)
    for _ in range(3):
        pb._lib.check(lib.plume_generate_fields(C.byref(env.c_config), C.byref(env.c_state), None, 1024, None, None,
                                                torch.cuda.current_stream().cuda_stream), "generate")
elif what == "k2":
    n = 1 << 20
    env = pb.VecMethaneEnv(n, device=dev, field_mode="procedural", auto_reset=True, seed=2)
    acts = torch.randint(0, 5, (n,), dtype=torch.int32, device=dev)
    for _ in range(12):
        env.step(acts)
elif what == "iteration":
    tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=4096 * 256 // 4)
    for _ in range(3):
        tr.train_iteration()
elif what.startswith("head"):
    H = int(what[4:])
    tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=4096 * 256 // 4, lstm_hidden=H)
    for _ in range(3):
        tr.rollout_only()
elif what == "eval_fixed":
    torch.manual_seed(0)
    model = pb.PPOActorCritic(device=dev)
    pb.evaluate_policy(model, stop="fixed", num_envs=4096, seed=1)
torch.cuda.synchronize()
print("done", what)
