"""Trajectory and per-episode logging in the reference's on-disk layouts.

* ``training_data.nc`` (``NetCDFWriter``, PPOV2.1/model.py:351-419, schema in PPOV2.1/nc_info.txt): per episode a
  row of ``x``, ``y``, ``concentration`` [max_steps] (NaN fill), ``is_source`` int8, and the scalars ``source_x``,
  ``source_y``, ``source_concentration``, ``gaussian_sigma``, ``peak_concentration``.  The driver writes every
  episode with ``source_x/y = gaussian_params['mu_x'/'mu_y']`` and ``source_conc = peak``
  (train_ppo2.0.py:222-233; the conditional first write of :207-220 is always overwritten by it), and the writer's
  quirk is kept: the LAST logged step's x/y are overwritten with the source coordinates and flagged ``is_source``
  (model.py:410-412).  netCDF4 is not available in this image, so the same variable names go into an ``.npz``.
* the 11-column per-episode statistics of ``training_results*.csv`` (train_ppo2.0.py:128-134,236-248):
  ``Final_Conc`` = ``conc_field`` at the final cell when the source was reached, else 0.0 (:145,194-196);
  ``Current_Radius`` = the trainer's radius when the episode ended, before its own curriculum update (:247,251).

Episodes are assembled on the device by ``plume_trajectory_log`` (csrc/trajectory_kernels.cu: five small launches
per ``[T, N]`` segment, no torch op); an episode that spans rollout segments is carried in a per-env row until it
finishes.  Episode order is the canonical one of the curriculum (step-major, then env index)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

CSV_COLUMNS = ["Episode", "Total_Reward", "Success", "Conc_Reward", "Explore_Reward", "Move_Penalty", "TKE_Penalty",
               "Boundary_Penalty", "Steps", "Final_Conc", "Current_Radius"]


class TrajectoryLogger:
    def __init__(self, env, max_episodes: int = 2000):
        self.env, self.cfg = env, env.cfg
        self.lib = _lib.load()
        self.max_episodes, self.S = int(max_episodes), int(env.cfg.max_steps)
        dev, N, S, E = env.device, env.num_envs, self.S, self.max_episodes
        nan = float("nan")
        self.x = torch.full((E, S), nan, dtype=torch.float32, device=dev)
        self.y = torch.full((E, S), nan, dtype=torch.float32, device=dev)
        self.conc = torch.full((E, S), nan, dtype=torch.float32, device=dev)
        self.steps = torch.zeros(E, dtype=torch.int32, device=dev)
        self.source = torch.full((E, 2), nan, dtype=torch.float32, device=dev)
        self.success = torch.zeros(E, dtype=torch.uint8, device=dev)
        self.radius = torch.zeros(E, dtype=torch.float64, device=dev)
        self.sums = torch.zeros(E, 6, dtype=torch.float64, device=dev)   # total reward + the 5 info components
        self.final_conc = torch.zeros(E, dtype=torch.float32, device=dev)
        self.count_t = torch.zeros(1, dtype=torch.int32, device=dev)
        # carry of the unfinished episode of every env
        self.c_x = torch.empty(N, S, dtype=torch.float32, device=dev)
        self.c_y = torch.empty(N, S, dtype=torch.float32, device=dev)
        self.c_conc = torch.empty(N, S, dtype=torch.float32, device=dev)
        self.c_sums = torch.zeros(N, 6, dtype=torch.float64, device=dev)
        self.c_len = torch.zeros(N, dtype=torch.int32, device=dev)
        p = lambda t: t.data_ptr()
        self._c = _lib.TrajLog(E, S, N, 0, p(self.x), p(self.y), p(self.conc), p(self.steps), p(self.source),
                               p(self.success), p(self.radius), p(self.sums), p(self.final_conc), p(self.c_x),
                               p(self.c_y), p(self.c_conc), p(self.c_sums), p(self.c_len), p(self.count_t))
        self._ws = None
        self._count_host = 0

    @property
    def count(self) -> int:
        """Episodes logged so far (reads the device counter)."""
        self._count_host = int(self.count_t.item())
        return self._count_host

    def consume(self, buf, trainer=None, comm=None, sync: bool = True, radius: float | None = None) -> int:
        """Adds the finished episodes of one rollout segment (a buffer built with ``with_trajectory=True``; with
        ``with_info=True`` the five reward components are summed too).  ``trainer``: the ``PPOTrainer`` whose
        ``update_from_rollout`` has ALREADY run on this segment -- its per-window radii give every episode the
        reference's 'Current_Radius'; ``radius`` gives it explicitly (the single-env driver, which logs
        ``trainer.current_radius`` before ``trainer.update``); without either the env's current radius is used.  ``comm``: the ``PeerComm`` of a
        multi-GPU run (global episode ordinals).  Returns how many episodes were appended (``sync=False``: 0, the
        call stays asynchronous)."""
        if buf.pos_out is None or buf.conc_out is None:
            raise ValueError("TrajectoryLogger needs RolloutEngine(..., with_trajectory=True)")
        T, N, dev = buf.filled, buf.num_envs, buf.device
        need = int(self.lib.plume_trajectory_workspace_bytes(T, N))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        before = self._count_host if not sync else int(self.count_t.item())
        wr, fallback = None, 0.0
        if radius is not None:
            fallback = float(radius)
        elif trainer is not None and getattr(trainer, "_window_radius", None) is not None:
            wr = trainer.window_radius()
        else:
            fallback = float(self.env.curriculum[0].item())
        cb = buf.c_rollout_buffers(None, None, None)
        with torch.cuda.device(dev):
            rc = self.lib.plume_trajectory_log(C.byref(self._c), C.byref(cb), T, comm._h if comm is not None else None,
                                               int(self.cfg.window_size), _lib.ptr(wr), fallback,
                                               self._ws.data_ptr(), int(self._ws.numel()),
                                               torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "plume_trajectory_log")
        if not sync:
            return 0
        return self.count - before

    # -- output ---------------------------------------------------------------------------------
    def nc_variables(self) -> dict:
        """The variables of training_data.nc (nc_info.txt) for the episodes logged so far."""
        E, S = self.count, self.S
        x, y = self.x[:E].clone(), self.y[:E].clone()
        is_source = torch.zeros(E, S, dtype=torch.int8, device=x.device)
        rows = torch.arange(E, device=x.device)
        last = (self.steps[:E].long() - 1).clamp_(0, S - 1)
        is_source[rows, last] = 1                                   # model.py:410
        x[rows, last] = self.source[:E, 0]                          # model.py:411-412 (sic)
        y[rows, last] = self.source[:E, 1]
        peak = np.full(E, self.cfg.conc_peak, dtype=np.float32)
        return {"episode": np.arange(E, dtype=np.int32), "step": np.arange(S, dtype=np.int32),
                "x": x.cpu().numpy(), "y": y.cpu().numpy(), "concentration": self.conc[:E].cpu().numpy(),
                "is_source": is_source.cpu().numpy(), "source_concentration": peak.copy(),
                "source_x": self.source[:E, 0].cpu().numpy(), "source_y": self.source[:E, 1].cpu().numpy(),
                "gaussian_sigma": np.full(E, self.cfg.sigma, dtype=np.float32), "peak_concentration": peak}

    def csv_rows(self) -> np.ndarray:
        """[E, 11] in the column order of training_results*.csv (CSV_COLUMNS)."""
        E = self.count
        s = self.sums[:E].cpu().numpy()
        return np.column_stack([np.arange(1, E + 1), s[:, 0], self.success[:E].cpu().numpy().astype(np.int64), s[:, 1],
                                s[:, 2], s[:, 3], s[:, 4], s[:, 5], self.steps[:E].cpu().numpy(),
                                self.final_conc[:E].cpu().numpy().astype(np.float64), self.radius[:E].cpu().numpy()])

    def save(self, nc_like_path: str, csv_path: str | None = None) -> None:
        np.savez_compressed(nc_like_path, **self.nc_variables())
        if csv_path:
            np.savetxt(csv_path, self.csv_rows(), delimiter=",", header=",".join(CSV_COLUMNS), comments="")
