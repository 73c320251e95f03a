"""Times one PPO iteration (4096 envs x 256 steps, 5 epochs x 4 minibatches) and the update part alone with a
variant build of the library: python profiles/debug/variant_bench.py [profiles/debug/libplume_b200_vA.so | -] [chunks]"""
import sys
import torch
sys.path.insert(0, '.')
import uav_wrf_les_ppo_lstm_b200 as pb
if len(sys.argv) > 1 and sys.argv[1] != "-":
    pb._lib.LIB_PATH = sys.argv[1]
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=4096 * 256 // 4)
if len(sys.argv) > 2:                       # rollout collected in overlapped chunks: e.g. 32,32,32,160; "0" = single launch
    tr.engine.overlap_chunks = tuple(int(v) for v in sys.argv[2].split(",")) if sys.argv[2] != "0" else None
if len(sys.argv) > 3:                       # "0": the stream waits for the stop head right after the rollout
    tr.defer_head_join = sys.argv[3] != "0"
for _ in range(3):
    tr.train_iteration()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 8
e0.record()
for _ in range(n):
    tr.train_iteration()
e1.record()
torch.cuda.synchronize()
r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rr = []
for _ in range(12):
    tr.train_iteration(rollout_events=(r0, r1))
    torch.cuda.synchronize()
    rr.append(r0.elapsed_time(r1))
print(f"{' '.join(sys.argv[1:]) or 'product'}: {e0.elapsed_time(e1) / n:.3f} ms / iteration, rollout + stop head "
      + " ".join(f"{v:.3f}" for v in rr) + " ms")
