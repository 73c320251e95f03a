"""Batched drop-in for the reference ``MethaneEnv`` (PPOV2.1/environment.py:19-178).

``VecMethaneEnv`` keeps the reference's method names and attributes with a leading ``N``
dimension and CUDA tensors; ``MethaneEnv`` is the ``N = 1`` view that returns numpy/python
scalars exactly like the reference so the unmodified ``train_ppo*.py`` loop can drive it.
All arithmetic runs in libplume_b200.so (csrc/env_kernels.cu); torch only owns the device
memory and the stream.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .config import FIELD_F32, FIELD_F64, FIELD_MODES, FIELD_PROCEDURAL, PlumeConfig, config_for

INFO_KEYS = _lib.INFO_KEYS


def host_tables(cfg):
    """The two per-step quantities of ``MethaneEnv.step`` that only depend on an integer, evaluated on the
    host exactly as the reference does and handed to the kernels as float32 tables:
    ``step_count / MAX_STEPS`` (environment.py:85, python int / int -> float64 -> float32) and
    ``visit_count**0.75 + 1`` (environment.py:140, python int ** float -> C pow, weak float -> float32)."""
    m = int(cfg.max_steps)
    step_frac = (np.arange(m + 1, dtype=np.float64) / float(m)).astype(np.float32)
    visit_denom = np.array([float(v) ** 0.75 + 1 for v in range(m + 2)], dtype=np.float64).astype(np.float32)
    return np.ascontiguousarray(step_frac), np.ascontiguousarray(visit_denom)


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("VecMethaneEnv needs a CUDA device: the plume kernels have no CPU fallback")
    return device


class VecMethaneEnv:
    """N plume environments stepped in lockstep on one GPU.

    Parameters mirror the reference where it has them; the rest select the batched layout:
    ``field_mode`` in {"procedural", "f32", "f64"/"materialised"}; ``seed`` keys the Philox
    streams; ``env_id_base`` is the global id of env 0 (``rank * num_envs``) so that results
    do not depend on how envs are sharded over GPUs.
    """

    def __init__(self, num_envs: int = 1, device="cuda", version: str = "2.1", seed: int = 0,
                 field_mode: str = "procedural", auto_reset: bool = False, env_id_base: int = 0,
                 config: PlumeConfig | None = None, plume_model: str = "isotropic", fast_reward: bool = False):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.cfg = config if config is not None else config_for(version)
        self.num_envs = int(num_envs)
        self.field_mode = FIELD_MODES[field_mode] if isinstance(field_mode, str) else int(field_mode)
        self.auto_reset = bool(auto_reset)
        if self.auto_reset and self.field_mode != FIELD_PROCEDURAL:
            raise ValueError("auto_reset needs field_mode='procedural' (materialised fields are regenerated "
                             "by reset(), which launches the field kernel)")
        self.seed = int(seed)
        # fast_reward: flags / indices keep the reference's float64 arithmetic, reward terms and the concentration
        # observation are float32 (fp32 rel 1e-5 instead of bit-exact float64 rewards); see PLUME_FLAG_FAST_REWARD
        self.fast_reward = bool(fast_reward)
        # "isotropic" = the reference code (parity target); "dispersion" = the README's plume/state/reward
        self.plume_model = _lib.PLUME_MODELS[plume_model] if isinstance(plume_model, str) else int(plume_model)
        self.grid_size = self.cfg.grid_size                                   # environment.py:23
        self.action_space = SimpleNamespace(n=_lib.NUM_ACTIONS)                # environment.py:24
        self.observation_space = SimpleNamespace(shape=(_lib.OBS_DIM,), dtype=np.float32)
        self.cell_size = self.cfg.cell_size                                    # environment.py:37
        self.min_radius = self.cfg.min_radius
        self.radius_decay = self.cfg.radius_decay
        N, G, dev = self.num_envs, self.cfg.grid_size, self.device
        with torch.cuda.device(dev):
            z = lambda dt, *s: torch.zeros(*s, dtype=dt, device=dev)
            self.pos_x, self.pos_y = z(torch.float32, N), z(torch.float32, N)
            self.src_x, self.src_y = z(torch.float64, N), z(torch.float64, N)
            self.step_count_t = z(torch.int32, N)
            self.episode_idx = z(torch.int32, N)
            self.visited_t = z(torch.int16, N, _lib.VISIT_STRIDE)
            self.last_move_t = z(torch.int8, N)
            self.radius_t = torch.full((N,), self.cfg.initial_radius, dtype=torch.float64, device=dev)
            self.explore_bonus_t = torch.full((N,), self.cfg.explore_bonus, dtype=torch.float64, device=dev)
            # the two curriculum scalars resets latch (model.py:189-190)
            self.curriculum = torch.tensor([self.cfg.initial_radius, self.cfg.explore_bonus],
                                           dtype=torch.float64, device=dev)
            g = np.arange(G)
            self.sin_tab = torch.from_numpy(np.sin(0.05 * g)).to(dev)          # environment.py:59
            self.cos_tab = torch.from_numpy(np.cos(0.07 * g)).to(dev)
            sf, vd = host_tables(self.cfg)
            self.step_frac_tab = torch.from_numpy(sf).to(dev)                  # environment.py:85
            self.visit_denom_tab = torch.from_numpy(vd).to(dev)                # environment.py:140
            # tke at the current float32 cell, carried from one step to the next (procedural mode)
            self.cell_tke_t = z(torch.float64, N)
            self.cell_conc_t = z(torch.float64, N)
            self.cell_key_t = z(torch.int32, N)
            self.conc_field_t = self.tke_field_t = None
            if self.field_mode in (FIELD_F32, FIELD_F64):
                fdt = torch.float32 if self.field_mode == FIELD_F32 else torch.float64
                self.conc_field_t = torch.empty(N, G, G, dtype=fdt, device=dev)
                self.tke_field_t = torch.empty(N, G, G, dtype=fdt, device=dev)
            self.obs = z(torch.float32, N, _lib.OBS_DIM)
            self.final_obs = z(torch.float32, N, _lib.OBS_DIM)
            self.reward = z(torch.float64, N)
            self.done = z(torch.uint8, N)
            self.reached = z(torch.uint8, N)
            self.info_t = z(torch.float32, _lib.INFO_DIM, N)
        self._ccfg = _lib.make_env_config(self.cfg, self.field_mode, self.seed, self.plume_model)
        self._cstate = _lib.EnvState(
            N, int(env_id_base), self.pos_x.data_ptr(), self.pos_y.data_ptr(), self.src_x.data_ptr(),
            self.src_y.data_ptr(), self.step_count_t.data_ptr(), self.episode_idx.data_ptr(),
            self.visited_t.data_ptr(), self.radius_t.data_ptr(), self.explore_bonus_t.data_ptr(),
            _lib.ptr(self.conc_field_t), _lib.ptr(self.tke_field_t), self.sin_tab.data_ptr(),
            self.cos_tab.data_ptr(), self.curriculum.data_ptr(), self.last_move_t.data_ptr(),
            self.step_frac_tab.data_ptr(), self.visit_denom_tab.data_ptr(), self.cell_tke_t.data_ptr(),
            self.cell_conc_t.data_ptr(), self.cell_key_t.data_ptr())
        self.launches = 0
        self.reset()                                                           # environment.py:40

    # -- curriculum scalars (written by PPOTrainer.update, model.py:189-190) -------------------
    @property
    def current_radius(self) -> float:
        return float(self.curriculum[0].item())

    @current_radius.setter
    def current_radius(self, value: float) -> None:
        self.curriculum[0] = float(value)
        if not self.auto_reset:      # the reference's attribute takes effect immediately
            self.radius_t.fill_(float(value))

    @property
    def explore_bonus(self) -> float:
        return float(self.curriculum[1].item())

    @explore_bonus.setter
    def explore_bonus(self, value: float) -> None:
        self.curriculum[1] = float(value)
        if not self.auto_reset:
            self.explore_bonus_t.fill_(float(value))

    # -- helpers ----------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, name, *args):
        with torch.cuda.device(self.device):
            rc = getattr(self.lib, name)(*args)
        _lib.check(rc, name)
        self.launches += 1

    def _idx(self, env_ids):
        if env_ids is None:
            return None, self.num_envs
        ids = torch.as_tensor(env_ids, dtype=torch.int32, device=self.device).contiguous()
        return ids, int(ids.numel())

    # -- P0/P1: reset ------------------------------------------------------------------------
    def reset(self, env_ids=None, u_src=None, fields=None):
        """Resets all envs (or ``env_ids``).  ``u_src`` [n,2] float64 injects the two
        ``rand()`` of environment.py:44; ``fields`` = (conc, tke) [n,G,G] injects the plume
        (field_mode f64/f32) instead of generating it from the Philox stream."""
        ids, n = self._idx(env_ids)
        us = None
        if u_src is not None:
            us = torch.as_tensor(u_src, dtype=torch.float64, device=self.device).reshape(n, 2).contiguous()
        self._call("plume_env_reset", C.byref(self._ccfg), C.byref(self._cstate), _lib.ptr(ids), n,
                   _lib.ptr(us), self._stream())
        if self.field_mode != FIELD_PROCEDURAL:
            if fields is not None:
                conc, tke = fields
                sel = slice(None) if ids is None else ids.long()
                self.conc_field_t[sel] = torch.as_tensor(conc, device=self.device).to(self.conc_field_t.dtype)
                self.tke_field_t[sel] = torch.as_tensor(tke, device=self.device).to(self.tke_field_t.dtype)
            else:
                self._call("plume_generate_fields", C.byref(self._ccfg), C.byref(self._cstate), _lib.ptr(ids), n,
                           None, None, self._stream())
        return self.observe()

    def set_source(self, src, env_ids=None):
        """Overrides the source position (float64 [n,2]) of already reset envs."""
        ids, n = self._idx(env_ids)
        s = torch.as_tensor(src, dtype=torch.float64, device=self.device).reshape(n, 2)
        sel = slice(None) if ids is None else ids.long()
        self.src_x[sel] = s[:, 0]
        self.src_y[sel] = s[:, 1]
        self.cell_key_t[sel] = 0          # the carried concentration belongs to the old source

    def observe(self) -> torch.Tensor:
        """P3 ``_get_obs`` for every env, [N,6] float32."""
        self._call("plume_env_observe", C.byref(self._ccfg), C.byref(self._cstate), self.obs.data_ptr(),
                   self._stream())
        return self.obs

    # -- P2: step ----------------------------------------------------------------------------
    def step(self, actions, step_noise=None, noise_out=None):
        """One lockstep step.  Returns ``(obs [N,6] f32, reward [N] f64, done [N] bool, info)``
        with the reference's five info keys as [N] float32 tensors (plus ``reached`` and, with
        auto-reset, ``final_obs``)."""
        a = torch.as_tensor(actions, device=self.device).to(torch.int32).reshape(self.num_envs).contiguous()
        zn = None
        if step_noise is not None:
            zn = torch.as_tensor(step_noise, dtype=torch.float64, device=self.device).reshape(self.num_envs, 2).contiguous()
        flags = (_lib.FLAG_AUTO_RESET if self.auto_reset else 0) | (_lib.FLAG_FAST_REWARD if self.fast_reward else 0)
        self._call("plume_env_step", C.byref(self._ccfg), C.byref(self._cstate), a.data_ptr(), _lib.ptr(zn), flags,
                   self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self.reached.data_ptr(),
                   self.info_t.data_ptr(), self.final_obs.data_ptr() if self.auto_reset else None,
                   _lib.ptr(noise_out), self._stream())
        info = {k: self.info_t[i] for i, k in enumerate(INFO_KEYS)}
        info["reached"] = self.reached.bool()
        if self.auto_reset:
            info["final_obs"] = self.final_obs
        return self.obs, self.reward, self.done.bool(), info

    # -- field access ------------------------------------------------------------------------
    def field_noise(self, env_ids=None):
        """(z, u) [n,G,G] float32: the randn/rand draws of the listed envs' current plume."""
        ids, n = self._idx(env_ids)
        G = self.cfg.grid_size
        z = torch.empty(n, G, G, dtype=torch.float32, device=self.device)
        u = torch.empty_like(z)
        ccfg = _lib.make_env_config(self.cfg, FIELD_PROCEDURAL, self.seed)
        self._call("plume_generate_fields", C.byref(ccfg), C.byref(self._cstate), _lib.ptr(ids), n, z.data_ptr(),
                   u.data_ptr(), self._stream())
        return z, u

    def field_noise_at(self, env_local, x, y):
        el = torch.as_tensor(env_local, dtype=torch.int32, device=self.device).contiguous()
        xx = torch.as_tensor(x, dtype=torch.int32, device=self.device).contiguous()
        yy = torch.as_tensor(y, dtype=torch.int32, device=self.device).contiguous()
        z = torch.empty(el.numel(), dtype=torch.float32, device=self.device)
        u = torch.empty_like(z)
        self._call("plume_field_noise_at", C.byref(self._ccfg), C.byref(self._cstate), el.data_ptr(), xx.data_ptr(),
                   yy.data_ptr(), int(el.numel()), z.data_ptr(), u.data_ptr(), self._stream())
        return z, u

    def materialise_fields(self, env_ids=None, dtype=torch.float64):
        """(conc, tke) [n,G,G] of the listed envs, generated by the field kernel (K1) -- the
        accessor behind ``conc_field`` when the env runs in procedural mode."""
        ids, n = self._idx(env_ids)
        G = self.cfg.grid_size
        sel = slice(None) if ids is None else ids.long()
        if self.field_mode != FIELD_PROCEDURAL:
            return self.conc_field_t[sel], self.tke_field_t[sel]
        mode = FIELD_F64 if dtype == torch.float64 else FIELD_F32
        conc = torch.empty(n, G, G, dtype=dtype, device=self.device)
        tke = torch.empty_like(conc)
        # a compact view: state arrays indexed by the global env id, fields by the list position
        full_conc = torch.empty(0)
        st = _lib.EnvState.from_buffer_copy(self._cstate)
        if ids is None:
            st.conc_field, st.tke_field = conc.data_ptr(), tke.data_ptr()
            ccfg = _lib.make_env_config(self.cfg, mode, self.seed, self.plume_model)
            self._call("plume_generate_fields", C.byref(ccfg), C.byref(st), None, n, None, None, self._stream())
            return conc, tke
        # per-env generation into the compact buffers
        for j, e in enumerate(ids.tolist()):
            full = _lib.EnvState.from_buffer_copy(self._cstate)
            off = (e * G * G) * conc.element_size()
            full.conc_field = conc[j].data_ptr() - off
            full.tke_field = tke[j].data_ptr() - off
            one = torch.tensor([e], dtype=torch.int32, device=self.device)
            ccfg = _lib.make_env_config(self.cfg, mode, self.seed, self.plume_model)
            self._call("plume_generate_fields", C.byref(ccfg), C.byref(full), one.data_ptr(), 1, None, None,
                       self._stream())
        del full_conc
        return conc, tke

    def conc_at(self, x, y) -> torch.Tensor:
        """``conc_field[x, y]`` for every env (x, y integer [N] tensors), float64."""
        xx = torch.as_tensor(x, device=self.device).to(torch.int32).contiguous()
        yy = torch.as_tensor(y, device=self.device).to(torch.int32).contiguous()
        conc = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        self._call("plume_field_at", C.byref(self._ccfg), C.byref(self._cstate), xx.data_ptr(), yy.data_ptr(),
                   conc.data_ptr(), None, self._stream())
        return conc

    # -- reference attribute names ----------------------------------------------------------
    @property
    def agent_pos(self) -> torch.Tensor:
        return torch.stack([self.pos_x, self.pos_y], dim=1)

    @property
    def source_pos(self) -> torch.Tensor:
        return torch.stack([self.src_x, self.src_y], dim=1)

    @property
    def step_count(self) -> torch.Tensor:
        return self.step_count_t

    @property
    def visited(self) -> torch.Tensor:
        D = self.cfg.grid_divisions
        return self.visited_t[:, :100].reshape(self.num_envs, 10, 10)[:, :D, :D]

    @property
    def gaussian_params(self) -> dict:                                       # environment.py:64-69
        return {"mu_x": self.src_x, "mu_y": self.src_y, "sigma": self.cfg.sigma, "peak": self.cfg.conc_peak}

    @property
    def c_config(self):
        return self._ccfg

    @property
    def c_state(self):
        return self._cstate


class MethaneEnv(VecMethaneEnv):
    """``N = 1`` drop-in with the reference's exact return types: numpy observation,
    float reward, bool done, dict of floats (one host round trip per step)."""

    def __init__(self, device="cuda", version: str = "2.1", seed: int = 0, field_mode: str = "f64", **kw):
        self.trajectory: list = []
        self._fields_host = None          # (conc, tke) numpy arrays of the CURRENT episode, downloaded on first use
        super().__init__(1, device=device, version=version, seed=seed, field_mode=field_mode, **kw)

    def reset(self, *a, **k):
        self.trajectory = []
        self._fields_host = None
        return super().reset(*a, **k)[0].cpu().numpy()

    def set_source(self, *a, **k):
        self._fields_host = None
        return super().set_source(*a, **k)

    def _get_obs(self):
        return self.observe()[0].cpu().numpy()

    def step(self, action, step_noise=None):
        """``step(action)`` as in the reference; ``step_noise`` (float64 [2]) optionally injects the ``randn(2)`` of
        environment.py:108 (parity tests replay the reference's draws)."""
        if step_noise is not None:
            step_noise = np.asarray(step_noise, dtype=np.float64).reshape(1, 2)
        obs, reward, done, info = super().step(torch.tensor([int(action)]), step_noise=step_noise)
        torch.cuda.current_stream(self.device).synchronize()
        o = obs[0].cpu().numpy()
        out_info = {k: float(info[k][0].item()) for k in INFO_KEYS}
        reached = bool(info["reached"][0].item())
        self.trajectory.append({"pos": self.agent_pos, "conc": o[2], "tke": o[3], "reached": reached})
        return o, float(reward[0].item()), bool(done[0].item()), out_info

    @property
    def agent_pos(self):
        return np.array([self.pos_x[0].item(), self.pos_y[0].item()], dtype=np.float32)

    @property
    def source_pos(self):
        return np.array([self.src_x[0].item(), self.src_y[0].item()], dtype=np.float64)

    @property
    def step_count(self):
        return int(self.step_count_t[0].item())

    def _host_fields(self):
        # the reference driver indexes env.conc_field every step (train_ppo2.0.py:167-170): one download per
        # episode, not per access (the plume only changes in reset())
        if self._fields_host is None:
            conc, tke = self.materialise_fields()
            self._fields_host = (conc[0].cpu().numpy(), tke[0].cpu().numpy())
        return self._fields_host

    @property
    def conc_field(self):
        return self._host_fields()[0]

    @property
    def tke_field(self):
        return self._host_fields()[1]

    @property
    def gaussian_params(self):
        s = self.source_pos
        return {"mu_x": s[0], "mu_y": s[1], "sigma": self.cfg.sigma, "peak": self.cfg.conc_peak}
