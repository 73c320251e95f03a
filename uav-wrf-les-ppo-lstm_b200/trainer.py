"""``PlumeTrainer``: the batched equivalent of ``train_ppo()`` (PPOV2.1/train_ppo2.0.py:109-267)
for one GPU (one process per GPU; ranks hold disjoint env shards and all-reduce the gradient).

One ``train_iteration()`` = one fused rollout segment of ``horizon`` lockstep steps over all envs
(policy + env + LSTM stop head, one kernel), the on-device curriculum, GAE and
``epochs x minibatches`` optimiser steps.  Nothing returns to the host inside an iteration."""
from __future__ import annotations

import torch

from . import _lib
from .config import PlumeConfig, config_for
from .env import VecMethaneEnv
from .learner import (MATERIALISE_PERM_MIN, FusedAdam, PPOTrainer, UpdateWorkspace, materialise_permutations,
                      update_model)
from .model import PeakAndStopPredictor, PPOActorCritic
from .rollout import RolloutEngine, overlap_split


class PlumeTrainer:
    def __init__(self, num_envs: int = 4096, horizon: int = 256, version: str = "2.1", device="cuda",
                 seed: int = 0, minibatch_size: int | None = None, stop_head: bool = True,
                 rank: int = 0, world_size: int = 1, process_group=None, with_info: bool = False,
                 with_trend: bool = True, cfg: PlumeConfig | None = None, gradient_exchange: str = "peer",
                 with_trajectory: bool = False, lstm_hidden: int = 32, overlap_stop_head: bool = True):
        self.cfg = cfg or config_for(version)
        self.device = torch.device(device)
        self.rank, self.world_size, self.process_group = rank, world_size, process_group
        self.num_envs, self.horizon = int(num_envs), int(horizon)
        torch.manual_seed(seed)                       # identical initial weights on every rank
        self.env = VecMethaneEnv(num_envs, device=self.device, version=version, seed=seed,
                                 field_mode="procedural", auto_reset=True, env_id_base=rank * num_envs,
                                 config=self.cfg)
        self.model = PPOActorCritic(device=self.device)
        self.head = PeakAndStopPredictor(hidden_dim=lstm_hidden, device=self.device) if stop_head else None
        self.engine = RolloutEngine(self.env, self.model, self.head, horizon=horizon, with_info=with_info,
                                    with_trend=with_trend and stop_head, with_trajectory=with_trajectory)
        # the stop head of the first rows runs on the SMs the lockstep kernel leaves idle (rollout.overlap_split)
        if overlap_stop_head and stop_head and self.device.type == "cuda":
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            self.engine.overlap_chunks = overlap_split(self.horizon, self.num_envs, sms)
        self.optimizer = FusedAdam(self.model, lr=self.cfg.learning_rate, max_grad_norm=self.cfg.max_grad_norm)
        # multi-GPU, gradient_exchange="peer" (default): every exchange step of the iteration -- gradient all-reduce
        # fused with clip + Adam, the advantage statistics, the curriculum's flag codes -- goes over NVLink peer
        # memory (csrc/comm_kernels.cu), no host-launched collective.  "nccl" keeps torch.distributed for all three.
        if gradient_exchange not in ("peer", "nccl"):
            raise ValueError("gradient_exchange must be 'peer' or 'nccl'")
        self.comm = None
        if process_group is not None and world_size > 1 and gradient_exchange == "peer":
            from .dist import PeerComm
            self.comm = PeerComm(process_group, _lib.MLP_PARAMS, self.device, code_bytes=num_envs * horizon)
            self.optimizer.attach_comm(self.comm)
            # the flag codes are published on a side stream right after the lockstep kernel, under the stop-head kernel
            self._code_stream = torch.cuda.Stream(device=self.device)
            self._loop_done = torch.cuda.Event()
            self._codes_done = torch.cuda.Event()
            self.engine.after_loop = self._publish_codes
        self.curriculum = PPOTrainer(self.env, self.model, self.optimizer, cfg=self.cfg)
        self.minibatch_size = int(minibatch_size or (num_envs * horizon) // 4)
        self.workspace = UpdateWorkspace(self.device, min(self.minibatch_size, num_envs * horizon))
        self.iteration = 0
        self.last_losses = None
        # True: the stream goes on behind the lockstep kernels, so that curriculum, flag exchange and GAE (which do not
        # read the stop head's outputs) run under its last launch, joined before the update.  Measured: 10.12 against
        # 10.09 ms on one GPU, no difference on two (profiles/r2_notes.md) -- the head's CTAs hold every SM -- so off.
        self.defer_head_join = False
        n_mb = (num_envs * horizon + self.minibatch_size - 1) // self.minibatch_size
        # my kernels per iteration: rollout (+ deferred stop head), curriculum (count, scan, bin, apply), gae scan +
        # normalise, (weight prep, gradient, clip + Adam) per optimiser step; + sample records and the epoch
        # permutations when they are materialised (below)
        self.launches_per_iteration = 1 + (1 if stop_head else 0) + 4 + 2 + 3 * self.cfg.epochs * n_mb
        if self.engine.overlap_chunks is not None:
            self.launches_per_iteration += 2 * (len(self.engine.overlap_chunks) - 1)
        # the epoch permutations do not depend on the rollout: they are written out on a side stream while the
        # (latency-bound) rollout kernel runs
        self._perm_stream = None
        if self.device.type == "cuda" and num_envs * horizon >= MATERIALISE_PERM_MIN:
            self._perm_stream = torch.cuda.Stream(device=self.device)
            self._perm_done = torch.cuda.Event()
            self._perm_free = torch.cuda.Event()
            self.launches_per_iteration += self.cfg.epochs + 1

    def _permutations_async(self):
        """Launches this iteration's ``epochs`` permutations on the side stream; returns the ``[epochs, T*N]`` tensor."""
        M = self.num_envs * self.horizon
        out = self.workspace.perm_buffer(self.cfg.epochs, M)
        with torch.cuda.stream(self._perm_stream):
            if self.iteration > 0:
                self._perm_stream.wait_event(self._perm_free)     # the previous update has consumed the buffer
            materialise_permutations(_lib.load(), out, M, self.iteration, range(self.cfg.epochs),
                                     self._perm_stream.cuda_stream)
            self._perm_done.record(self._perm_stream)
        return out

    def _publish_codes(self, buf, T: int) -> None:
        cur = torch.cuda.current_stream(self.device)
        self._loop_done.record(cur)
        self._code_stream.wait_event(self._loop_done)
        self.comm.publish_codes(buf.flag_code[:T], stream=self._code_stream.cuda_stream)
        self._codes_done.record(self._code_stream)

    def check(self) -> None:
        """Raises if an exchange kernel of the PREVIOUS iteration timed out (that step was not applied)."""
        if self.comm is not None:
            self.comm.check_async_end()

    def train_iteration(self, check_nan: bool = False, rollout_events=None):
        """``rollout_events``: optional (start, end) CUDA events recorded around the rollout (bench.py)."""
        self.check()                               # waits for the previous iteration's error word, not for this one
        if rollout_events is not None:
            rollout_events[0].record(torch.cuda.current_stream(self.device))
        buf = self.engine.collect(join=not self.defer_head_join,
                                  end_event=rollout_events[1] if rollout_events is not None else None)
        perms = None
        if self._perm_stream is not None:          # launched after the rollout, so its CTAs are placed first
            with torch.cuda.device(self.device):
                perms = self._permutations_async()
        if self.comm is not None:
            torch.cuda.current_stream(self.device).wait_event(self._codes_done)
            self.curriculum.update_from_rollout(buf, self.process_group, comm=self.comm, codes_published=True)
        else:
            self.curriculum.update_from_rollout(buf, self.process_group)
        if perms is not None:
            torch.cuda.current_stream(self.device).wait_event(self._perm_done)
        self.engine.join_stop_head()
        self.last_losses = update_model(buf, self.model, self.optimizer, cfg=self.cfg,
                                        minibatch_size=self.minibatch_size, workspace=self.workspace,
                                        process_group=self.process_group, perm_seed=self.iteration, perms=perms,
                                        check_nan=check_nan)
        if perms is not None:
            self._perm_free.record(torch.cuda.current_stream(self.device))
        if self.comm is not None:
            self.comm.check_async_begin()
        self.iteration += 1
        return self.last_losses

    def rollout_only(self):
        return self.engine.collect()

    # -- host-resident model: the end-to-end path with host buffers -------------------------------
    def make_host_buffers(self):
        """Pinned host buffers of the public host API: parameters in/out, metrics out."""
        n_steps = self.cfg.epochs * ((self.num_envs * self.horizon + self.minibatch_size - 1) // self.minibatch_size)
        lstm_n = sum(p.numel() for p in self.head.parameters()) if self.head is not None else 0
        return {
            "params_in": torch.empty(_lib.MLP_PARAMS, dtype=torch.float32).pin_memory(),
            "lstm_in": torch.empty(max(lstm_n, 1), dtype=torch.float32).pin_memory(),
            "curriculum_in": torch.empty(2, dtype=torch.float64).pin_memory(),
            "params_out": torch.empty(_lib.MLP_PARAMS, dtype=torch.float32).pin_memory(),
            "losses_out": torch.empty(n_steps, 4, dtype=torch.float64).pin_memory(),
            "curriculum_out": torch.empty(8, dtype=torch.float64).pin_memory(),
            "episodes_out": torch.empty(2, dtype=torch.float32).pin_memory(),
        }

    def train_iteration_host(self, hb: dict) -> None:
        """One iteration for a caller that keeps the model on the host: uploads the parameters
        (and the curriculum scalars) from pinned memory, runs the iteration, downloads the updated
        parameters, the per-step losses, the curriculum state and the episode/success counts."""
        self.model.flat.copy_(hb["params_in"], non_blocking=True)
        if self.head is not None:
            off = 0
            for p in self.head.parameters():
                n = p.numel()
                p.data.copy_(hb["lstm_in"][off:off + n].view_as(p), non_blocking=True)
                off += n
        self.env.curriculum.copy_(hb["curriculum_in"], non_blocking=True)
        losses = self.train_iteration()
        hb["params_out"].copy_(self.model.flat, non_blocking=True)
        hb["losses_out"].copy_(losses, non_blocking=True)
        hb["curriculum_out"].copy_(self.curriculum.device_state(), non_blocking=True)
        buf = self.engine.buffer
        hb["episodes_out"].copy_(torch.stack([buf.dones[: buf.filled].sum(), buf.reached[: buf.filled].float().sum()]),
                                 non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        hb["params_in"].copy_(hb["params_out"])
        hb["curriculum_in"][0] = hb["curriculum_out"][0]
        hb["curriculum_in"][1] = hb["curriculum_out"][1]

    def host_bytes_per_iteration(self, hb: dict):
        h2d = sum(hb[k].numel() * hb[k].element_size() for k in ("params_in", "lstm_in", "curriculum_in"))
        d2h = sum(hb[k].numel() * hb[k].element_size() for k in ("params_out", "losses_out", "curriculum_out",
                                                                "episodes_out"))
        return h2d, d2h
