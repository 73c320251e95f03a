"""Trajectory and per-episode logging in the reference's on-disk layouts.

* ``training_data.nc`` (PPOV2.1/model.py:351-419, schema in PPOV2.1/nc_info.txt): per episode a row of
  ``x``, ``y``, ``concentration`` [max_steps] (NaN fill), ``is_source`` int8, and the scalars ``source_x``,
  ``source_y``, ``source_concentration``, ``gaussian_sigma``, ``peak_concentration``.  The writer's quirk is
  kept: the LAST logged step's x/y are overwritten with the source coordinates and flagged ``is_source``
  (model.py:410-412).  netCDF4 is not available in this image, so the same variable names go into an
  ``.npz``.
* the 11-column per-episode statistics of ``training_results*.csv`` (train_ppo2.0.py:128-134,236-248).

Episodes are assembled on the device from the ``[T, N]`` rollout buffers with scatter ops (no per-step
Python loop); an episode that spans rollout segments is carried in a per-env row until it finishes.
Episode order is the canonical one of the curriculum (step-major, then env index)."""
from __future__ import annotations

import numpy as np
import torch

CSV_COLUMNS = ["Episode", "Total_Reward", "Success", "Conc_Reward", "Explore_Reward", "Move_Penalty", "TKE_Penalty",
               "Boundary_Penalty", "Steps", "Final_Conc", "Current_Radius"]


class TrajectoryLogger:
    def __init__(self, env, max_episodes: int = 2000):
        self.env, self.cfg = env, env.cfg
        self.max_episodes, self.S = int(max_episodes), int(env.cfg.max_steps)
        dev, N, S, E = env.device, env.num_envs, self.S, self.max_episodes
        nan = float("nan")
        self.x = torch.full((E, S), nan, dtype=torch.float32, device=dev)
        self.y = torch.full((E, S), nan, dtype=torch.float32, device=dev)
        self.conc = torch.full((E, S), nan, dtype=torch.float32, device=dev)
        self.steps = torch.zeros(E, dtype=torch.int32, device=dev)
        self.source = torch.full((E, 2), nan, dtype=torch.float32, device=dev)
        self.success = torch.zeros(E, dtype=torch.bool, device=dev)
        self.radius = torch.zeros(E, dtype=torch.float64, device=dev)
        self.sums = torch.zeros(E, 6, dtype=torch.float64, device=dev)   # total reward + the 5 info components
        self.final_conc = torch.zeros(E, dtype=torch.float32, device=dev)
        self.count = 0
        # carry of the unfinished episode of every env
        self.c_x = torch.full((N, S), nan, dtype=torch.float32, device=dev)
        self.c_y = torch.full((N, S), nan, dtype=torch.float32, device=dev)
        self.c_conc = torch.full((N, S), nan, dtype=torch.float32, device=dev)
        self.c_sums = torch.zeros(N, 6, dtype=torch.float64, device=dev)

    @torch.no_grad()
    def consume(self, buf) -> int:
        """Adds the finished episodes of one rollout segment (needs a buffer built with
        ``with_trajectory=True, with_info=True`` and a stop head or ``conc_sample``).  Returns how many
        episodes were appended."""
        if buf.pos_out is None or buf.info is None or buf.conc_sample is None:
            raise ValueError("TrajectoryLogger needs RolloutEngine(..., with_info=True, with_trajectory=True)")
        T, N, S, dev = buf.filled, buf.num_envs, self.S, buf.device
        done = buf.dones[:T] != 0
        k = torch.round(buf.obs[:T, :, 4].double() * S).long().clamp_(0, S - 1)      # step index inside the episode
        tt = torch.arange(T, device=dev).unsqueeze(1).expand(T, N)
        # time of the done that closes the episode a transition belongs to (T = not in this segment)
        nd = torch.where(done, tt, torch.full_like(tt, T))
        next_done = torch.flip(torch.cummin(torch.flip(nd, [0]), 0).values, [0])
        order = torch.cumsum(done.reshape(-1).long(), 0).reshape(T, N) - 1          # canonical rank of each done
        n_new = int(done.sum().item())
        n_take = max(0, min(n_new, self.max_episodes - self.count))
        cols = torch.arange(N, device=dev).unsqueeze(0).expand(T, N)
        closes = next_done < T
        slot = torch.full((T, N), -1, dtype=torch.long, device=dev)
        slot[closes] = order[next_done[closes], cols[closes]] + self.count
        slot[slot >= self.max_episodes] = -1
        x, y = buf.pos_out[:T, :, 0], buf.pos_out[:T, :, 1]
        conc = buf.conc_sample[:T] * float(self.cfg.conc_peak)
        info = buf.info[:T].permute(0, 2, 1).double()                               # [T, N, 5]
        contrib = torch.cat([buf.rewards[:T].double().unsqueeze(-1), info], dim=-1)  # [T, N, 6]
        # (1) episodes that end here and started in an earlier segment take the carried rows first
        first_done = done & (torch.cumsum(done.long(), 0) == 1)
        fd_t, fd_n = first_done.nonzero(as_tuple=True)
        fd_slot = slot[fd_t, fd_n]
        ok = fd_slot >= 0
        fd_n, fd_slot = fd_n[ok], fd_slot[ok]
        self.x[fd_slot], self.y[fd_slot], self.conc[fd_slot] = self.c_x[fd_n], self.c_y[fd_n], self.c_conc[fd_n]
        self.sums[fd_slot] = self.c_sums[fd_n]
        nan = float("nan")
        self.c_x[fd_n], self.c_y[fd_n], self.c_conc[fd_n] = nan, nan, nan
        self.c_sums[fd_n] = 0.0
        # (2) scatter this segment's transitions into their episode rows / into the carry
        fin = slot >= 0
        self.x[slot[fin], k[fin]] = x[fin]
        self.y[slot[fin], k[fin]] = y[fin]
        self.conc[slot[fin], k[fin]] = conc[fin]
        self.sums.index_put_((slot[fin],), contrib[fin], accumulate=True)
        open_ = ~closes
        self.c_x[cols[open_], k[open_]] = x[open_]
        self.c_y[cols[open_], k[open_]] = y[open_]
        self.c_conc[cols[open_], k[open_]] = conc[open_]
        self.c_sums.index_put_((cols[open_],), contrib[open_], accumulate=True)
        # (3) per-episode scalars at the closing transition
        d_t, d_n = done.nonzero(as_tuple=True)
        d_slot = slot[d_t, d_n]
        ok = d_slot >= 0
        d_t, d_n, d_slot = d_t[ok], d_n[ok], d_slot[ok]
        self.steps[d_slot] = (k[d_t, d_n] + 1).int()
        self.source[d_slot] = buf.src_out[d_t, d_n]
        self.success[d_slot] = buf.reached[d_t, d_n] != 0
        self.final_conc[d_slot] = conc[d_t, d_n]
        self.radius[d_slot] = float(self.env.curriculum[0].item())
        self.count += n_take
        return n_take

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
                                np.full(E, self.cfg.conc_peak), self.radius[:E].cpu().numpy()])

    def save(self, nc_like_path: str, csv_path: str | None = None) -> None:
        np.savez_compressed(nc_like_path, **self.nc_variables())
        if csv_path:
            np.savetxt(csv_path, self.csv_rows(), delimiter=",", header=",".join(CSV_COLUMNS), comments="")
