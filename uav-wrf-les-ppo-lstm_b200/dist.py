"""Multi-GPU plumbing: one process per GPU, envs sharded by rank, no rollout traffic.

The only exchange steps of the path are (1) three doubles per update -- {sum, sum of squares,
count} of the raw advantages, so that every rank normalises with the *global* statistics
(train_ppo2.0.py:34-38 over the whole batch) -- and (2) one all-reduce (sum) of the flat gradient
per minibatch; each rank divides its loss by the global minibatch size, so the sum is the
gradient of the global mean loss (train_ppo2.0.py:70-87); and (3) one all-gather of the segment's
done/reached flags per update (5 B per transition), so that every rank applies the curriculum
(model.py:188-221) to the GLOBAL episode stream in canonical order (step-major, then global env id) and
all ranks hold the same radius / explore bonus.  NCCL on the GPUs, gloo in CPU tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None, device: torch.device | None = None):
    """Initialises torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun).  Returns
    (rank, world_size, process_group or None)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world <= 1:
        return 0, 1, None
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return rank, world, dist.group.WORLD


def env_shard(rank: int, envs_per_rank: int):
    """Global env ids owned by ``rank``: [rank*n, (rank+1)*n).  Philox counters use the global id,
    so a run does not depend on how many ranks the envs are spread over."""
    base = rank * envs_per_rank
    return base, range(base, base + envs_per_rank)


def allreduce_stats(stats: torch.Tensor, process_group=None) -> torch.Tensor:
    """Sum-reduces the {sum, sumsq, count} advantage statistics in place."""
    if process_group is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=process_group)
    return stats


def allreduce_gradient(flat_grad: torch.Tensor, process_group=None) -> torch.Tensor:
    """Sum-reduces the flat gradient in place (one collective per minibatch, 145 KB)."""
    if process_group is not None:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=process_group)
    return flat_grad


def global_minibatch(local_minibatch: int, process_group=None) -> int:
    return local_minibatch * (dist.get_world_size(process_group) if process_group is not None else 1)


def normalisation_from_stats(stats: torch.Tensor):
    """(mean, denominator) exactly as csrc/learner_kernels.cu::gae_normalise_kernel derives them
    from {sum, sumsq, count}: unbiased std, replaced by 1 if < 1e-6 or NaN, + 1e-6."""
    s1, s2, n = (float(x) for x in stats.tolist())
    mean = s1 / n
    var = max((s2 - n * mean * mean) / (n - 1.0), 0.0) if n > 1 else float("nan")
    sd = var ** 0.5
    if not (sd >= 1e-6):
        sd = 1.0
    return mean, sd + 1e-6


def gather_episode_flags(dones: torch.Tensor, reached: torch.Tensor, process_group=None):
    """``dones`` float32 [T, N], ``reached`` uint8 [T, N] of this rank -> the same flags of ALL ranks in
    canonical order [T, world * N] (global env id = rank * N + local id)."""
    if process_group is None:
        return dones, reached
    world = dist.get_world_size(process_group)
    T, N = dones.shape
    d_all = torch.empty(world, T, N, dtype=dones.dtype, device=dones.device)
    r_all = torch.empty(world, T, N, dtype=reached.dtype, device=reached.device)
    dist.all_gather(list(d_all.unbind(0)), dones.contiguous(), group=process_group)
    dist.all_gather(list(r_all.unbind(0)), reached.contiguous(), group=process_group)
    return (d_all.permute(1, 0, 2).reshape(T, world * N).contiguous(),
            r_all.permute(1, 0, 2).reshape(T, world * N).contiguous())


def gather_flag_codes(code: torch.Tensor, process_group=None) -> torch.Tensor:
    """``code`` uint8 [T, N] (bit 0 done, bit 1 reached) of this rank -> [world, T, N] of all ranks (1 byte per
    transition over the wire; the curriculum kernel indexes this layout directly)."""
    if process_group is None:
        return code.unsqueeze(0)
    world = dist.get_world_size(process_group)
    out = torch.empty((world,) + tuple(code.shape), dtype=code.dtype, device=code.device)
    if code.is_cuda:
        dist.all_gather_into_tensor(out, code.contiguous(), group=process_group)
    else:
        dist.all_gather(list(out.unbind(0)), code.contiguous(), group=process_group)
    return out


class PeerComm:
    """The exchange steps of a data-parallel iteration without NCCL: every rank's block (two gradient buffers, two
    flag-code segments, small vectors) is mapped into every peer (CUDA IPC over NVLink).
    ``plume_allreduce_clip_adam`` does all-reduce + clip + Adam in one kernel, ``allreduce_small`` sums the
    advantage statistics, ``publish_codes`` + ``plume_curriculum_update_peer`` replace the flag all-gather
    (csrc/comm_kernels.cu).  The 64-byte IPC handles travel through one ``all_gather_object`` at start-up."""

    def __init__(self, process_group, n_params: int, device, code_bytes: int = 0):
        import ctypes as C

        from . import _lib
        self.lib = _lib.load()
        self.process_group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.device = torch.device(device)
        self._h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_comm_create(self.world, self.rank, int(n_params), int(code_bytes),
                                                  C.byref(self._h), handle), "plume_comm_create")
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(handle), group=process_group)
            blob = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(gathered))
            _lib.check(self.lib.plume_comm_connect(self._h, blob), "plume_comm_connect")
        dist.barrier(group=process_group)
        # asynchronous read-back of the error word: a pinned host word + an event per iteration
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._err_event = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def allreduce_small(self, values: torch.Tensor) -> None:
        """``values`` (<= 32 doubles, device) <- sum over the ranks, in rank order."""
        from . import _lib
        assert values.dtype == torch.float64 and values.is_contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_comm_allreduce_small(self._h, values.data_ptr(), int(values.numel()),
                                                           self._stream()), "plume_comm_allreduce_small")

    def publish_codes(self, code: torch.Tensor, stream=None) -> None:
        """Copies this rank's ``[T, N]`` uint8 flag codes into its mapped block and signals the peers."""
        from . import _lib
        code = code.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_comm_publish_codes(self._h, code.data_ptr(), int(code.numel()),
                                                         stream if stream is not None else self._stream()),
                       "plume_comm_publish_codes")

    def error_code(self) -> int:
        import ctypes as C

        from . import _lib
        err = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_comm_error(self._h, C.byref(err), self._stream()), "plume_comm_error")
        return int(err.value)

    _MESSAGES = {1: "peer-memory exchange: a rank did not publish in time (the step was NOT applied)",
                 2: "peer-memory exchange: grid barrier timed out (the step was NOT applied)"}

    def check(self) -> None:
        """Synchronous check: raises if any exchange kernel of this communicator timed out."""
        err = self.error_code()
        if err:
            raise RuntimeError(self._MESSAGES.get(err, "comm error") + "; call reset() on every rank to continue")

    def check_async_begin(self) -> None:
        """Queues a read-back of the error word behind the work enqueued so far (no host synchronisation)."""
        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_comm_error_async(self._h, self._err_host.data_ptr(), self._stream()),
                       "plume_comm_error_async")
            self._err_event = torch.cuda.Event()
            self._err_event.record(torch.cuda.current_stream(self.device))

    def check_async_end(self) -> None:
        """Waits for the read-back queued by ``check_async_begin`` (i.e. for the previous iteration) and raises on
        a recorded timeout."""
        if self._err_event is None:
            return
        self._err_event.synchronize()
        self._err_event = None
        err = int(self._err_host[0])
        if err:
            raise RuntimeError(self._MESSAGES.get(err, "comm error") + "; call reset() on every rank to continue")

    def reset(self) -> None:
        """Collective recovery after a reported timeout: clears the error word and the exchange counters."""
        from . import _lib
        dist.barrier(group=self.process_group)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_comm_reset(self._h), "plume_comm_reset")
        self._err_host.zero_()
        self._err_event = None
        dist.barrier(group=self.process_group)

    def close(self) -> None:
        if self._h:
            self.lib.plume_comm_destroy(self._h)
            self._h = None
