"""Debug helper (torchrun): per-phase device time of one PPO iteration on every rank."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import uav_wrf_les_ppo_lstm_b200 as pb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pg = dist.group.WORLD
tr = pb.PlumeTrainer(num_envs=4096, horizon=256, device=dev, seed=0, rank=rank, world_size=world, process_group=pg,
                     minibatch_size=4096 * 256 // 4)
for _ in range(3):
    tr.train_iteration()
torch.cuda.synchronize(); dist.barrier()
def ev(): return torch.cuda.Event(enable_timing=True)
names = ["rollout", "curriculum(+flag exchange)", "advantages(+stats exchange)", "20 x (grad + exchange + adam)"]
acc = [0.0] * 4
K = 5
for _ in range(K):
    e = [ev() for _ in range(5)]
    e[0].record(); buf = tr.engine.collect()
    e[1].record()
    if tr.comm is not None:
        torch.cuda.current_stream().wait_event(tr._codes_done)
        tr.curriculum.update_from_rollout(buf, pg, comm=tr.comm, codes_published=True)
    else:
        tr.curriculum.update_from_rollout(buf, pg)
    e[2].record()
    ws = tr.workspace
    pb.compute_advantages(buf, tr.cfg, ws, pg, comm=tr.comm)
    e[3].record()
    # the update without recomputing advantages is not exposed; time the whole update and subtract
    pb.update_model(buf, tr.model, tr.optimizer, cfg=tr.cfg, minibatch_size=tr.minibatch_size, workspace=ws,
                    process_group=pg, perm_seed=tr.iteration, check_nan=False)
    e[4].record(); torch.cuda.synchronize()
    t = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
    t[3] -= t[2]
    acc = [a + x for a, x in zip(acc, t)]
    tr.iteration += 1
out = [None] * world
dist.all_gather_object(out, [a / K for a in acc])
if rank == 0:
    for i, n in enumerate(names):
        print(f"{n:36s}", " ".join(f"{o[i]:7.3f}" for o in out), "ms per rank")
dist.barrier(); dist.destroy_process_group()
