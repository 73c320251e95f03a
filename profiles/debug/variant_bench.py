"""Times one PPO iteration (4096 envs x 256 steps, 5 epochs x 4 minibatches) and the update part alone with a
variant build of the library: python profiles/debug/variant_bench.py profiles/debug/libplume_b200_vA.so"""
import sys
import torch
sys.path.insert(0, '.')
import uav_wrf_les_ppo_lstm_b200 as pb
if len(sys.argv) > 1:
    pb._lib.LIB_PATH = sys.argv[1]
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, minibatch_size=4096 * 256 // 4)
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
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'product'}: {e0.elapsed_time(e1) / n:.3f} ms / iteration")
