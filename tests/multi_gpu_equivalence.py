"""N ranks x M envs == 1 rank x N*M envs (run under torchrun on >= 2 GPUs; also launched by
tests/test_gpu_multi.py when the box has them).

Every rank runs ``PlumeTrainer`` on its env shard (Philox counters use the global env id, gradients and
advantage statistics are all-reduced over NCCL, the curriculum replays the all-gathered global episode stream);
rank 0 then runs the same job alone with all envs and compares: its shard of the rollout must be bit-identical,
the curriculum state identical, and -- with one minibatch per epoch, so that both jobs average over the same
sample set -- the updated parameters equal up to fp32 summation order."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uav_wrf_les_ppo_lstm_b200 as pb  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
    M, T, iters = 256, 48, 2
    tr = pb.PlumeTrainer(num_envs=M, horizon=T, device=dev, seed=3, rank=rank, world_size=world, process_group=pg,
                         minibatch_size=M * T)
    tr.env.curriculum[0] = 35.0          # episodes finish inside the test
    tr.env.reset()
    tr.curriculum.device_state()[0] = 35.0
    tr.curriculum.device_state()[2] = 35.0
    snaps = []
    for _ in range(iters):
        tr.train_iteration()
        b = tr.engine.buffer
        snaps.append((b.obs.clone(), b.actions.clone(), b.rewards.clone(), b.dones.clone(), b.stop_prob.clone()))
    torch.cuda.synchronize()
    if tr.comm is not None:
        tr.comm.check()
    params = tr.model.flat.clone()
    cur = tr.curriculum.device_state().clone()
    gathered = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    for g in gathered:
        assert torch.equal(g, params), "ranks diverged"
    ok = True
    if rank == 0:
        ref = pb.PlumeTrainer(num_envs=M * world, horizon=T, device=dev, seed=3, minibatch_size=M * world * T)
        ref.env.curriculum[0] = 35.0
        ref.env.reset()
        ref.curriculum.device_state()[0] = 35.0
        ref.curriculum.device_state()[2] = 35.0
        for it in range(iters):
            ref.train_iteration()
            b = ref.engine.buffer
            mine = snaps[it]
            if it == 0:      # the first rollout runs on identical parameters: bit-identical shard
                assert torch.equal(b.obs[:, :M], mine[0]) and torch.equal(b.actions[:, :M], mine[1])
                assert torch.equal(b.rewards[:, :M], mine[2]) and torch.equal(b.dones[:, :M], mine[3])
                assert torch.equal(b.stop_prob[:, :M], mine[4])
        torch.cuda.synchronize()
        rc = ref.curriculum.device_state()
        assert torch.equal(rc[4:8], cur[4:8]), (rc, cur)                 # same global episode stream
        assert float(rc[6]) > 0, "no episode finished: the test would be vacuous"
        moved = (ref.model.flat - ref.model.flat.new_tensor(0)).abs().max()
        err = (ref.model.flat - params).abs().max().item()
        step = ref.cfg.learning_rate * ref.cfg.epochs * iters
        print(f"exchange: {'peer-memory fused all-reduce+clip+Adam' if tr.comm is not None else 'NCCL all-reduce'}")
        print(f"multi-gpu equivalence: {world} ranks x {M} envs vs 1 x {M * world}: max |dparam| = {err:.3e} "
              f"(total Adam movement {step:.1e}), episodes {int(rc[6])}, successes {int(rc[7])}")
        ok = err < 0.05 * step
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_EQUIVALENCE_OK")


if __name__ == "__main__":
    main()
