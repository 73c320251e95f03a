"""The learner: ``update_model`` (= ``_update_model``, PPOV2.1/train_ppo2.0.py:14-87), the fused
clip+Adam optimiser (train_ppo2.0.py:86-87,113) and the curriculum ``PPOTrainer``
(model.py:178-221).  GAE, the loss/gradient, the optimiser step and the batched curriculum run
in csrc/learner_kernels.cu and csrc/ppo_kernels.cu; the only collective is one all-reduce of
the flat gradient per minibatch (plus three doubles for global advantage statistics)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import dist as pdist
from .config import PlumeConfig, config_for


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class FusedAdam:
    """``torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8)`` + ``clip_grad_norm_`` in one kernel
    over the model's flat parameter/gradient buffers."""

    def __init__(self, model_or_params, lr: float = 3e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float = 0.5):
        model = model_or_params
        if not hasattr(model, "flat"):
            raise TypeError("FusedAdam takes the PPOActorCritic module (its parameters share one flat buffer)")
        self.model = model
        self.lr, self.betas, self.eps, self.max_grad_norm = float(lr), tuple(betas), float(eps), float(max_grad_norm)
        self.exp_avg = torch.zeros_like(model.flat)
        self.exp_avg_sq = torch.zeros_like(model.flat)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=model.flat.device)
        self.step_count = 0
        self.launches = 0
        self.comm = None          # dist.PeerComm: all-reduce fused into the step (multi-GPU)

    def attach_comm(self, comm) -> None:
        """With a ``dist.PeerComm`` attached, ``step()`` = gradient all-reduce over NVLink peer memory + clip +
        Adam in ONE kernel; callers must then not all-reduce ``flat_grad`` themselves."""
        self.comm = comm

    def zero_grad(self) -> None:
        self.model.flat_grad.zero_()

    def step(self) -> None:
        m = self.model
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(m.flat.device):
            if self.comm is not None:
                rc = lib.plume_allreduce_clip_adam(self.comm._h, m.flat.data_ptr(), m.flat_grad.data_ptr(),
                                                   self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                                   _lib.MLP_PARAMS, self.max_grad_norm, self.lr, self.betas[0],
                                                   self.betas[1], self.eps, self.step_count,
                                                   self.grad_norm.data_ptr(), _stream(m.flat.device))
            else:
                rc = lib.plume_clip_adam(m.flat.data_ptr(), m.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), _lib.MLP_PARAMS, self.max_grad_norm, self.lr,
                                         self.betas[0], self.betas[1], self.eps, self.step_count,
                                         self.grad_norm.data_ptr(), _stream(m.flat.device))
        _lib.check(rc, "plume_clip_adam")
        self.launches += 1

    def state_dict(self) -> dict:
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_count,
                "lr": self.lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd: dict) -> None:
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])


class UpdateWorkspace:
    """Scratch memory of the update, allocated once and reused."""

    def __init__(self, device, max_minibatch: int):
        lib = _lib.load()
        self.max_minibatch = int(max_minibatch)
        self.bytes = int(lib.plume_ppo_workspace_bytes(self.max_minibatch))
        self.ws = torch.empty(self.bytes, dtype=torch.uint8, device=device)
        self.stats = torch.zeros(3, dtype=torch.float64, device=device)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._perm = None
        self._packed = None

    def packed_buffer(self, m: int) -> torch.Tensor:
        """``[m, 12]`` float32 sample records (plume_ppo_pack), allocated on first use."""
        if self._packed is None or self._packed.shape[0] < m:
            self._packed = torch.empty(m, 12, dtype=torch.float32, device=self.ws.device)
        return self._packed

    def perm_buffer(self, rows: int, m: int) -> torch.Tensor:
        """``[rows, m]`` int64 scratch for materialised epoch permutations (allocated on first use)."""
        if self._perm is None or self._perm.shape[0] < rows or self._perm.shape[1] != m:
            self._perm = torch.empty(rows, m, dtype=torch.int64, device=self.ws.device)
        return self._perm


# Above this many transitions the epoch permutation is written out by ``plume_permutation`` (one pass at full
# occupancy) instead of being evaluated inside the gradient kernel, where the ~300 dependent instructions of the
# Feistel index sit on a 4-warp path of every 128-sample tile (profiles/r1_notes.md, r1t)
MATERIALISE_PERM_MIN = 32768


def materialise_permutations(lib, out: torch.Tensor, m: int, perm_seed: int, epochs, stream) -> None:
    """out[i] = the Feistel permutation of [0, m) keyed by (perm_seed, epochs[i]) -- the index set the gradient kernel
    would otherwise derive per sample (same values, so results are unchanged)."""
    for row, epoch in enumerate(epochs):
        _lib.check(lib.plume_permutation(m, perm_seed, int(epoch), 0, m, out[row].data_ptr(), stream),
                   "plume_permutation")


def compute_advantages(buffer, cfg: PlumeConfig, ws: UpdateWorkspace, process_group=None, variant: str = "quirk",
                       last_values=None, comm=None) -> None:
    """P5: GAE reverse scan per env column + global normalisation; fills ``buffer.advantages``
    and ``buffer.returns`` (train_ppo2.0.py:17-39).  ``variant``: "quirk" (the V2.x/V1.1 update, the
    parity target), "bootstrap" (PPOV1.1/train_ppo1.0.py:66-89, needs ``last_values`` [N] = V(next state))
    or "v12" (PPOV1.2)."""
    lib = _lib.load()
    T, N, dev = buffer.filled, buffer.num_envs, buffer.device
    var = _lib.GAE_VARIANTS[variant]
    if last_values is not None:
        last_values = torch.as_tensor(last_values, dtype=torch.float32, device=dev).reshape(N).contiguous()
    ws.stats.zero_()
    with torch.cuda.device(dev):
        _lib.check(lib.plume_gae_scan_variant(buffer.rewards.data_ptr(), buffer.values.data_ptr(),
                                              buffer.dones.data_ptr(), _lib.ptr(last_values), T, N, cfg.gamma, cfg.lam,
                                              var, buffer.advantages.data_ptr(), ws.stats.data_ptr(), _stream(dev)),
                   "plume_gae_scan_variant")
        if comm is not None:             # three doubles over NVLink peer memory, no host-launched collective
            comm.allreduce_small(ws.stats)
        else:
            pdist.allreduce_stats(ws.stats, process_group)
        _lib.check(lib.plume_gae_normalise_variant(buffer.advantages.data_ptr(), buffer.values.data_ptr(), T * N,
                                                   ws.stats.data_ptr(), var, buffer.returns.data_ptr(), _stream(dev)),
                   "plume_gae_normalise_variant")


def update_model(buffer, model, optimizer, cfg: PlumeConfig | None = None, perms=None,
                 minibatch_size: int | None = None, workspace: UpdateWorkspace | None = None,
                 process_group=None, perm_seed: int = 0, check_nan: bool = True, record=None,
                 kernel_path: str = "auto"):
    """Drop-in for ``_update_model(buffer, model, optimizer)`` (train_ppo2.0.py:14-87).

    ``optimizer`` must be a ``FusedAdam`` over ``model`` (``torch.optim.Adam`` cannot see the kernels' flat
    gradient buffer: its ``zero_grad()`` drops ``p.grad``).  ``kernel_path``: "auto" (tcgen05 from 1024 samples per
    minibatch), "tensor" or "simt" -- which of the two kernel families computes the gradient.

    ``perms``: optional list of ``cfg.epochs`` int64 index permutations of the flat ``[T*N]``
    transition set (what ``torch.randperm`` gives the reference, :43); by default the kernels use
    a stateless Feistel bijection keyed by ``(perm_seed, epoch)``.  ``minibatch_size`` defaults
    to ``cfg.batch_size`` (256).  With ``process_group`` the advantage statistics and the
    gradient are all-reduced (each rank holds its own envs).  Returns a ``[steps, 4]`` float64
    device tensor of (loss, policy loss, value loss, entropy) per optimiser step."""
    cfg = cfg or config_for("2.1")
    if not isinstance(optimizer, FusedAdam) or optimizer.model is not model:
        raise TypeError("update_model needs a FusedAdam built over this model (the gradient kernels accumulate into "
                        "model.flat_grad, which torch.optim optimisers neither zero nor read)")
    kpath = _lib.KERNEL_PATHS[kernel_path]
    lib = _lib.load()
    dev = buffer.device
    T, N = buffer.filled, buffer.num_envs
    M = T * N
    if M == 0:
        return None
    mb = int(minibatch_size or cfg.batch_size)
    if workspace is None or workspace.max_minibatch < min(mb, M):
        workspace = UpdateWorkspace(dev, min(mb, M))
    comm = getattr(optimizer, "comm", None)
    compute_advantages(buffer, cfg, workspace, process_group, comm=comm)
    batch = _lib.PpoBatch(M, buffer.obs.data_ptr(), buffer.actions.data_ptr(), buffer.log_probs.data_ptr(),
                          buffer.advantages.data_ptr(), buffer.returns.data_ptr(), buffer.values.data_ptr(), None)
    if M >= MATERIALISE_PERM_MIN:
        # every sample is gathered epochs times: interleave the six arrays once into 48-byte records
        packed = workspace.packed_buffer(M)
        with torch.cuda.device(dev):
            _lib.check(lib.plume_ppo_pack(C.byref(batch), packed.data_ptr(), _stream(dev)), "plume_ppo_pack")
        batch.packed = packed.data_ptr()
    n_mb = (M + mb - 1) // mb
    losses = torch.zeros(cfg.epochs * n_mb, 4, dtype=torch.float64, device=dev)
    nccl_exchange = process_group is not None and comm is None
    if record is None and not nccl_exchange:
        # the optimiser loop runs behind the ABI (plume_ppo_update): same launches, no Python call per minibatch
        perm_all = None
        if perms is not None:
            if torch.is_tensor(perms) and perms.dim() == 2:
                perm_all = perms.to(device=dev, dtype=torch.int64).contiguous()
            else:
                perm_all = torch.stack([torch.as_tensor(p, dtype=torch.int64) for p in perms]).to(dev).contiguous()
            assert perm_all.shape == (cfg.epochs, M), "perms: one permutation of the M transitions per epoch"
        elif M >= MATERIALISE_PERM_MIN:
            perm_all = workspace.perm_buffer(cfg.epochs, M)
            materialise_permutations(lib, perm_all, M, perm_seed, range(cfg.epochs), _stream(dev))
        world = pdist.global_minibatch(1, process_group)
        with torch.cuda.device(dev):
            rc = lib.plume_ppo_update(model.flat.data_ptr(), model.flat_grad.data_ptr(), optimizer.exp_avg.data_ptr(),
                                      optimizer.exp_avg_sq.data_ptr(), C.byref(batch), _lib.ptr(perm_all), perm_seed,
                                      cfg.epochs, mb, world, cfg.clip_epsilon, cfg.entropy_beta, optimizer.max_grad_norm,
                                      optimizer.lr, optimizer.betas[0], optimizer.betas[1], optimizer.eps,
                                      optimizer.step_count + 1, comm._h if comm is not None else None,
                                      losses.data_ptr(), optimizer.grad_norm.data_ptr(), workspace.nan_flag.data_ptr(),
                                      workspace.ws.data_ptr(), workspace.bytes, kpath, _stream(dev))
        _lib.check(rc, "plume_ppo_update")
        optimizer.step_count += cfg.epochs * n_mb
        optimizer.launches += cfg.epochs * n_mb
    else:
        step = 0
        with torch.cuda.device(dev):
            for epoch in range(cfg.epochs):
                perm = None
                if perms is not None:
                    perm = torch.as_tensor(perms[epoch], dtype=torch.int64, device=dev).contiguous()
                elif M >= MATERIALISE_PERM_MIN:
                    perm = workspace.perm_buffer(1, M)[0]
                    materialise_permutations(lib, perm.unsqueeze(0), M, perm_seed, [epoch], _stream(dev))
                for start in range(0, M, mb):
                    size = min(mb, M - start)
                    optimizer.zero_grad()
                    rc = lib.plume_ppo_grad(model.flat.data_ptr(), C.byref(batch), _lib.ptr(perm), perm_seed, epoch, start,
                                            size, pdist.global_minibatch(size, process_group), cfg.clip_epsilon,
                                            cfg.entropy_beta, model.flat_grad.data_ptr(), losses[step].data_ptr(),
                                            workspace.nan_flag.data_ptr(), workspace.ws.data_ptr(), workspace.bytes,
                                            kpath, _stream(dev))
                    _lib.check(rc, "plume_ppo_grad")
                    if comm is None:
                        pdist.allreduce_gradient(model.flat_grad, process_group)      # NCCL; else fused into step()
                    optimizer.step()
                    if record is not None:
                        record.append(optimizer.grad_norm.clone())
                    step += 1
    if check_nan and int(workspace.nan_flag.item()) != 0:            # train_ppo2.0.py:57-61
        workspace.nan_flag.zero_()
        raise RuntimeError("NaN in probs")
    return losses


class PPOTrainer:
    """Curriculum (model.py:178-221).  ``update(success)`` is the reference's per-episode host
    logic (exact for the single-env driver); ``update_from_rollout(buffer)`` applies the same rule
    on the device to every finished episode of a ``[T,N]`` segment in canonical order
    (step-major, then env index) without a host round trip."""

    def __init__(self, env, model=None, optimizer=None, cfg: PlumeConfig | None = None):
        self.env, self.model, self.optimizer = env, model, optimizer
        self.cfg = cfg or getattr(env, "cfg", None) or config_for("2.1")
        self.success_history: list = []
        self.current_radius = self.cfg.initial_radius
        self.explore_bonus = self.cfg.explore_bonus
        self._dev_state = None
        self._window_radius = None

    def update(self, success) -> None:
        c = self.cfg
        self.env.current_radius = self.current_radius
        self.env.explore_bonus = self.explore_bonus
        self.success_history.append(bool(success))
        if len(self.success_history) > c.window_size:
            self.success_history.pop(0)
        full = len(self.success_history) >= c.window_size
        if full:
            rate = float(np.mean(self.success_history[-c.window_size:]))
            self.explore_bonus *= c.decay_factor ** (1 + rate)
        self.explore_bonus = max(self.explore_bonus, 0.1)
        if full:
            rate = float(np.mean(self.success_history[-c.window_size:]))
            if rate > c.success_threshold:
                self.current_radius = max(c.min_radius,
                                          self.current_radius * c.radius_decay ** (2 + 3 * (rate - c.success_threshold)))
            elif rate < 0.25:
                self.current_radius = min(c.initial_radius, self.current_radius * 1.1)
            env_r = self.env.current_radius
            if abs(self.current_radius - env_r) > 5:
                self.current_radius = env_r + 5 * float(np.sign(self.current_radius - env_r))
            self.success_history = []

    # -- batched, on device ----------------------------------------------------------------
    def device_state(self) -> torch.Tensor:
        if self._dev_state is None:
            s = torch.zeros(8, dtype=torch.float64, device=self.env.device)
            s[0], s[1], s[2], s[3] = self.current_radius, self.explore_bonus, self.current_radius, self.explore_bonus
            self._dev_state = s
        return self._dev_state

    def window_radius(self) -> torch.Tensor:
        """double[MAX_WINDOWS + 2] written by the last ``update_from_rollout``: [0] length of the carried partial
        window at entry, [1] count, [2 + b] the trainer's radius in force for the episodes of window b -- the
        per-episode 'Current_Radius' of the reference's CSV (train_ppo2.0.py:247)."""
        if self._window_radius is None:
            self._window_radius = torch.zeros(_lib.CURRICULUM_MAX_WINDOWS + 2, dtype=torch.float64,
                                              device=self.env.device)
        return self._window_radius

    def flag_codes(self, buffer) -> torch.Tensor:
        T = buffer.filled
        if getattr(buffer, "flag_code_valid", False):
            return buffer.flag_code[:T]
        # rows appended with store(): derive the packed flags
        if not getattr(buffer, "reached_stored", True):
            raise ValueError("update_from_rollout: rows were stored without `reached` (PPOBuffer.store(..., reached=)); "
                             "the curriculum would count every episode as a failure")
        return ((buffer.dones[:T] != 0).to(torch.uint8) | (buffer.reached[:T] != 0).to(torch.uint8) * 2).contiguous()

    def update_from_rollout(self, buffer, process_group=None, comm=None, codes_published: bool = False) -> None:
        """With ``process_group`` the flags of all ranks are gathered first (or, with a ``dist.PeerComm``, read
        straight from the peers' mapped buffers), so every rank replays the same global episode stream and ends
        with identical curriculum scalars."""
        lib = _lib.load()
        c, st = self.cfg, self.device_state()
        T = buffer.filled
        consts = (c.initial_radius, c.min_radius, c.radius_decay, c.success_threshold, c.window_size, c.decay_factor)
        wr = self.window_radius()
        with torch.cuda.device(self.env.device):
            if comm is not None:
                if not codes_published:
                    comm.publish_codes(self.flag_codes(buffer))
                rc = lib.plume_curriculum_update_peer(comm._h, T, buffer.num_envs, st.data_ptr(),
                                                      self.env.curriculum.data_ptr(), *consts, wr.data_ptr(),
                                                      _stream(self.env.device))
            else:
                codes = pdist.gather_flag_codes(self.flag_codes(buffer).contiguous(), process_group)   # [world, T, N]
                rc = lib.plume_curriculum_update_packed(codes.data_ptr(), T, buffer.num_envs, codes.shape[0],
                                                        st.data_ptr(), self.env.curriculum.data_ptr(), *consts,
                                                        wr.data_ptr(), _stream(self.env.device))
        _lib.check(rc, "plume_curriculum_update")

    def sync_from_device(self) -> dict:
        s = self.device_state().cpu()
        if s[4] < 0:
            raise RuntimeError("curriculum kernel overflow: too many finished episodes in one segment")
        self.current_radius, self.explore_bonus = float(s[0]), float(s[1])
        return {"radius": float(s[0]), "explore_bonus": float(s[1]), "window_len": int(s[4]),
                "window_successes": int(s[5]), "episodes": int(s[6]), "successes": int(s[7])}
