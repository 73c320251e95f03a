"""torchrun: device time of the optimiser-step kernel alone (fused all-reduce + clip + Adam over peer memory, or plain
clip + Adam at one rank), back to back and after a 0.4 ms dummy kernel on every rank (the situation inside the update)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import uav_wrf_les_ppo_lstm_b200 as pb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, device=dev, seed=0, rank=rank, world_size=world,
                     process_group=dist.group.WORLD, minibatch_size=4096 * 256 // 4)
tr.train_iteration()
opt = tr.optimizer
torch.cuda.synchronize(); dist.barrier()
def ev(): return torch.cuda.Event(enable_timing=True)
filler = torch.empty(1 << 24, device=dev)
for label, gap in (("back to back", False), ("after ~0.4 ms of other work", True)):
    ts = []
    for _ in range(60):
        if gap:
            filler.normal_()
        a, b = ev(), ev()
        a.record(); opt.step(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[10:])
    out = [None] * world
    dist.all_gather_object(out, (ts[len(ts) // 2], ts[0], ts[-1]))
    if rank == 0:
        print(f"{label:32s} median/min/max us per rank:", " ".join(f"{m:.1f}/{lo:.1f}/{hi:.1f}" for m, lo, hi in out))
dist.barrier(); dist.destroy_process_group()
