"""CPU restatement of the reference plume environment -- TEST INFRASTRUCTURE ONLY.

This file is the parity oracle for the batched CUDA plume kernels.  It restates, in
numpy, what ``MethaneEnv`` does in the reference (all citations are
``/root/reference/PPOV2.1/environment.py:line`` unless another file is named; the
V2.0 and V1.1 variants differ only in sigma, the clip upper bound and MAX_STEPS, see
:class:`PlumeConfig`).  It is **pinned**: ``tests/test_oracle_vs_reference.py`` runs the
imported, unmodified reference next to it on injected noise/action traces and demands
bit-equal observations, rewards, flags and visit tables, and ``tests/golden/`` holds
vectors produced by the reference itself (``oracle/make_golden.py``).

Two forms are provided:

* :class:`OracleVecEnv` -- N environments stepped in lockstep with numpy arrays; every
  operation keeps the *scalar* dtype the reference computes in (float64 position math,
  float32 observation-derived reward terms under numpy>=2 promotion rules), so that for
  each environment the results are bit-identical to the scalar reference.
* :class:`OracleScalarEnv` -- one environment, python/numpy scalars, the same cost
  profile as the reference loop; used as the timed CPU baseline ("port").

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  The product never does.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

# action -> unit displacement, environment.py:98-100 (0 stay, 1 +y, 2 -y, 3 +x, 4 -x)
MOVE_TABLE = np.array([[0, 0], [0, 1], [0, -1], [1, 0], [-1, 0]], dtype=np.float64)


@dataclass(frozen=True)
class PlumeConfig:
    """Constants of one reference version (PPOV*/config.py)."""

    version: str = "2.1"
    grid_size: int = 500            # config.py:6
    max_steps: int = 1000           # config.py:7 (V1.1: 5000, PPOV1.1/config.py:7)
    conc_peak: float = 100.0        # config.py:8 / :13
    turbulence_intensity: float = 3.0   # config.py:9
    sigma: float = 15.0             # config.py:12 (V2.0/V1.1: grid/16, PPOV2.0/environment.py:54)
    clip_hi: float = 499.0          # environment.py:112 (V1.1: grid-1e-6, PPOV1.1/environment.py:105)
    explore_bonus: float = 0.6      # config.py:25
    grid_divisions: int = 10        # config.py:27
    initial_radius: float = 50.0    # config.py:31
    min_radius: float = 5.0
    radius_decay: float = 0.9
    success_threshold: float = 0.6
    window_size: int = 120
    decay_factor: float = 0.999
    conc_reward_coef: float = 2.0   # config.py:38
    tke_penalty_factor: float = 0.4
    boundary_penalty: float = 0.1
    boundary_decay_start: float = 0.15
    gamma: float = 0.99
    lam: float = 0.95
    clip_epsilon: float = 0.2
    entropy_beta: float = 0.01
    learning_rate: float = 3e-5
    batch_size: int = 256
    epochs: int = 5
    # "isotropic" = the reference code (pinned); "dispersion" = the README's plume/state/reward
    # (README.md:48-53,95-100: no reference code exists for it => this restatement is UNPINNED)
    plume_model: str = "isotropic"

    @property
    def cell_size(self) -> int:     # environment.py:37
        return self.grid_size // self.grid_divisions

    @property
    def move_step(self) -> float:   # environment.py:98
        return self.grid_size * 0.05


def config_for(version: str) -> PlumeConfig:
    base = PlumeConfig()
    if version == "2.1":
        return base
    if version == "2.0":
        return replace(base, version="2.0", sigma=base.grid_size / 16)
    if version == "1.1":
        return replace(base, version="1.1", sigma=base.grid_size / 16, max_steps=5000,
                       clip_hi=base.grid_size - 1e-6)
    raise ValueError(version)


def visit_denominator_table(max_count: int) -> np.ndarray:
    """``visit_count**0.75 + 1`` exactly as python evaluates it (int ** float -> C pow),
    environment.py:140."""
    return np.array([float(v) ** 0.75 + 1 for v in range(max_count + 2)], dtype=np.float64)


# --------------------------------------------------------------------------------------
# plume field, environment.py:52-63
# --------------------------------------------------------------------------------------
def wave_tables(cfg: PlumeConfig):
    """sin(0.05 x) and cos(0.07 y) over the integer grid, environment.py:59."""
    g = np.arange(cfg.grid_size)
    return np.sin(0.05 * g), np.cos(0.07 * g)


def plume_fields(cfg: PlumeConfig, src, z_field, u_field):
    """(conc_field, tke_field) in float64 for one source position, environment.py:53-63.

    ``z_field``/``u_field`` stand for ``np.random.randn(G,G)`` / ``np.random.rand(G,G)``.
    Index order is ``field[x, y]`` (x = first axis, ``np.mgrid``)."""
    G = cfg.grid_size
    x, y = np.mgrid[:G, :G]
    dist = np.sqrt((x - src[0]) ** 2 + (y - src[1]) ** 2)
    base = cfg.conc_peak * np.exp(-dist ** 2 / (2 * (cfg.sigma) ** 2))
    turbulence = cfg.turbulence_intensity * (
        np.abs(np.asarray(z_field, dtype=np.float64))
        + 0.3 * np.sin(0.05 * x) * np.cos(0.07 * y)
        + 0.2 * np.asarray(u_field, dtype=np.float64)
    )
    conc = np.clip(base + turbulence, 0, cfg.conc_peak)
    return conc, turbulence


WIND_MAX_SPEED = 5.0
DISPERSION_REF = 10.0


def dispersion_base(cfg: PlumeConfig, ddx, ddy, wind_c, wind_s):
    """README plume (README.md:50,97; parity unpinned): Gaussian dispersion with sigma_y = 0.3 x^0.71 in the
    wind-rotated frame, amplitude peak * min(1, sigma_y(10) / sigma_y(x)), zero upwind of the source."""
    xd = ddx * wind_c + ddy * wind_s
    yc = ddy * wind_c - ddx * wind_s
    down = xd > 0.0
    sig = 0.3 * np.power(np.where(down, xd, 1.0), 0.71)
    sig0 = 0.3 * DISPERSION_REF ** 0.71
    amp = np.where(sig0 < sig, sig0 / sig, 1.0)
    return np.where(down, cfg.conc_peak * amp * np.exp(-(yc * yc) / (2.0 * sig * sig)), 0.0)


def plume_cells(cfg: PlumeConfig, src_x, src_y, x, y, z, u, wind=None):
    """Same arithmetic as :func:`plume_fields` for individual cells (vectorised over the
    leading dimension): returns (conc, tke) float64 at integer cells (x, y) given the
    cell's two noise draws.  ``wind`` = (cos, sin) arrays selects the README dispersion base."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    if wind is not None:
        base = dispersion_base(cfg, x - src_x, y - src_y, wind[0], wind[1])
    else:
        dist = np.sqrt((x - src_x) ** 2 + (y - src_y) ** 2)
        base = cfg.conc_peak * np.exp(-dist ** 2 / (2 * (cfg.sigma) ** 2))
    turbulence = cfg.turbulence_intensity * (
        np.abs(np.asarray(z, dtype=np.float64)) + 0.3 * np.sin(0.05 * x) * np.cos(0.07 * y)
        + 0.2 * np.asarray(u, dtype=np.float64))
    return np.clip(base + turbulence, 0, cfg.conc_peak), turbulence


class MaterialisedFields:
    """conc/tke arrays ``[N,G,G]`` float64 held in memory, like the reference does."""

    def __init__(self, cfg: PlumeConfig, n: int):
        G = cfg.grid_size
        self.cfg = cfg
        self.conc = np.zeros((n, G, G), dtype=np.float64)
        self.tke = np.zeros((n, G, G), dtype=np.float64)

    def regenerate(self, i: int, src, z_field, u_field):
        self.conc[i], self.tke[i] = plume_fields(self.cfg, src, z_field, u_field)

    def at(self, idx, x, y):
        return self.conc[idx, x, y], self.tke[idx, x, y]


class CellNoiseFields:
    """Fields evaluated cell by cell from a noise callback ``noise(idx, x, y) -> (z, u)``
    (float32-representable values, e.g. read back from the CUDA generator).  Used when
    N is too large to hold ``[N,G,G]`` float64 arrays."""

    def __init__(self, cfg: PlumeConfig, n: int, noise):
        self.cfg = cfg
        self.noise = noise
        self.src = np.zeros((n, 2), dtype=np.float64)
        self.wind = None          # [n, 3] (cos, sin, speed) in the README dispersion model

    def regenerate(self, i: int, src, z_field=None, u_field=None):
        self.src[i] = src

    def at(self, idx, x, y):
        z, u = self.noise(idx, x, y)
        wind = None if self.wind is None else (self.wind[idx, 0], self.wind[idx, 1])
        return plume_cells(self.cfg, self.src[idx, 0], self.src[idx, 1], x, y, z, u, wind)


# --------------------------------------------------------------------------------------
# vectorised environment
# --------------------------------------------------------------------------------------
class OracleVecEnv:
    """N reference environments in lockstep; per-env results are bit-identical to
    ``MethaneEnv`` given the same injected draws."""

    def __init__(self, cfg: PlumeConfig, n: int, fields=None):
        self.cfg = cfg
        self.n = n
        D = cfg.grid_divisions
        self.fields = fields if fields is not None else MaterialisedFields(cfg, n)
        self.pos32 = np.zeros((n, 2), dtype=np.float32)          # agent_pos after astype(float32), :113
        self.src = np.zeros((n, 2), dtype=np.float64)            # source_pos, :44
        self.step_count = np.zeros(n, dtype=np.int64)
        self.visited = np.zeros((n, D, D), dtype=np.int64)       # defaultdict(int), :38
        self.current_radius = np.full(n, cfg.initial_radius, dtype=np.float64)   # :32
        self.explore_bonus = np.full(n, cfg.explore_bonus, dtype=np.float64)     # :39
        # python-float explore_bonus is "weak" under numpy>=2 promotion (float32 result);
        # once PPOTrainer.update has multiplied it by a np.float64 (model.py:197-199) it is
        # a strong float64 and explore_reward is computed in float64.
        self.explore_bonus_strong = False
        self.episode_idx = np.zeros(n, dtype=np.int64)
        self.last_reached = np.zeros(n, dtype=bool)
        self.wind = np.zeros((n, 3), dtype=np.float64)           # README model: (cos, sin, speed) per env
        self.last_move = np.zeros(n, dtype=np.int64)             # README model: last non-zero action
        self._pow = visit_denominator_table(cfg.max_steps)
        self._all = np.arange(n)

    # -- reset, :42-50 -----------------------------------------------------------------
    def reset_env(self, i: int, u_src, z_field=None, u_field=None, count_episode: bool = True):
        cfg = self.cfg
        padding = 50
        self.src[i] = np.asarray(u_src, dtype=np.float64) * (cfg.grid_size - 2 * padding) + padding   # :43-44
        self.fields.regenerate(i, self.src[i], z_field, u_field)                                     # :45
        self.pos32[i] = 0.0                                                                          # :46
        self.step_count[i] = 0
        self.visited[i] = 0
        if count_episode:
            self.episode_idx[i] += 1

    def set_source(self, i: int, src, z_field=None, u_field=None):
        """Reset with an explicit source position instead of a uniform draw."""
        self.src[i] = np.asarray(src, dtype=np.float64)
        self.fields.regenerate(i, self.src[i], z_field, u_field)
        self.pos32[i] = 0.0
        self.step_count[i] = 0
        self.visited[i] = 0
        self.last_move[i] = 0

    def set_wind(self, i: int, wind):
        """README model: (cos, sin, speed) of env ``i`` for its current episode."""
        self.wind[i] = wind
        if hasattr(self.fields, "wind"):
            if self.fields.wind is None:
                self.fields.wind = np.zeros((self.n, 3), dtype=np.float64)
            self.fields.wind[i] = wind

    # -- _get_obs, :71-87 --------------------------------------------------------------
    def _cells32(self):
        G = self.cfg.grid_size
        x = np.clip(self.pos32[:, 0].astype(np.int64), 0, G - 1)
        y = np.clip(self.pos32[:, 1].astype(np.int64), 0, G - 1)
        return x, y

    def observe(self) -> np.ndarray:
        cfg = self.cfg
        x, y = self._cells32()
        gx = x // cfg.cell_size
        gy = y // cfg.cell_size
        visit = self.visited[self._all, gx, gy]
        explore_level = np.minimum(visit / 5.0, 1.0)
        conc, tke = self.fields.at(self._all, x, y)
        obs = np.empty((self.n, 6), dtype=np.float32)
        obs[:, 0] = self.pos32[:, 0] / np.float32(cfg.grid_size)     # f32 / python int -> f32
        obs[:, 1] = self.pos32[:, 1] / np.float32(cfg.grid_size)
        obs[:, 2] = conc / cfg.conc_peak
        obs[:, 3] = tke / (cfg.turbulence_intensity * 3)
        obs[:, 4] = self.step_count / cfg.max_steps
        obs[:, 5] = explore_level
        if cfg.plume_model == "dispersion":      # README state: [CH4], wind vector, UAV position
            obs[:, 3] = self.wind[:, 0] * self.wind[:, 2] / WIND_MAX_SPEED
            obs[:, 5] = self.wind[:, 1] * self.wind[:, 2] / WIND_MAX_SPEED
        return obs

    # -- step, :89-178 -----------------------------------------------------------------
    def step(self, actions, z_step):
        """``actions`` int[N]; ``z_step`` [N,2] stands for ``np.random.randn(2)`` (:108).
        Returns (obs f32[N,6], reward f64[N], done bool[N], info dict of arrays)."""
        cfg = self.cfg
        G = cfg.grid_size
        a = np.asarray(actions, dtype=np.int64)
        z_step = np.asarray(z_step, dtype=np.float64)
        self.step_count += 1                                                         # :90

        px, py = self._cells32()                                                     # :93-94
        conc_prev, tke_prev = self.fields.at(self._all, px, py)
        prev_conc = conc_prev / cfg.conc_peak                                        # :95

        move_step = cfg.move_step                                                    # :98
        d = MOVE_TABLE[a] * move_step                                                # :99-100
        dnorm = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])                       # np.linalg.norm, exact: 0 or 25
        move_magnitude = dnorm / (cfg.grid_size * 0.05)                              # :101
        move_penalty = -0.15 * (1 - move_magnitude)                                  # :102

        turbulence_effect = (move_step * 0.2) * (
            (z_step * tke_prev[:, None]) / (cfg.turbulence_intensity * 3))           # :107-108

        new_pos = (self.pos32.astype(np.float64) + d) + turbulence_effect            # :111
        new_pos = np.clip(new_pos, 0, cfg.clip_hi)                                   # :112
        self.pos32 = new_pos.astype(np.float32)                                      # :113

        cx = np.clip(new_pos[:, 0].astype(np.int64), 0, G - 1)                       # :116-117
        cy = np.clip(new_pos[:, 1].astype(np.int64), 0, G - 1)
        conc_cur, _ = self.fields.at(self._all, cx, cy)
        current_conc = conc_cur / cfg.conc_peak                                      # :118
        conc_gradient = (current_conc - prev_conc) / (dnorm + 1e-6)                  # :119

        boundary_dist = np.minimum(
            np.minimum(new_pos[:, 0] / G, (G - new_pos[:, 0]) / G),
            np.minimum(new_pos[:, 1] / G, (G - new_pos[:, 1]) / G))                  # :121-126
        near = (boundary_dist < cfg.boundary_decay_start) & (conc_gradient < -0.01)  # :128
        gap = cfg.boundary_decay_start - boundary_dist
        boundary_penalty = np.where(near, -cfg.boundary_penalty * (gap * gap), 0.0)  # :129-131

        gx = (new_pos[:, 0] // cfg.cell_size).astype(np.int64)                       # :134-135
        gy = (new_pos[:, 1] // cfg.cell_size).astype(np.int64)
        self.visited[self._all, gx, gy] += 1                                         # :136
        visit_count = self.visited[self._all, gx, gy]                                # :137

        obs = self.observe()                                                         # :140 and :143
        denom = self._pow[visit_count]
        if self.explore_bonus_strong:
            explore_reward = (self.explore_bonus * (1 - obs[:, 5]).astype(np.float64)) / denom
        else:
            explore_reward = (self.explore_bonus.astype(np.float32) * (1 - obs[:, 5])) / denom.astype(np.float32)

        conc_reward = np.float32(cfg.conc_reward_coef) * obs[:, 2]                   # f32, :147
        tke_term = np.float32(cfg.tke_penalty_factor) * obs[:, 3]                    # f32, :150
        total = conc_reward + explore_reward                                         # f32 (f64 if strong)
        total = total.astype(np.float64) + move_penalty                              # f64 from here on
        total = total - tke_term.astype(np.float64)
        total = total + boundary_penalty                                             # :146-152
        if cfg.plume_model == "dispersion":
            # README reward R = d[CH4] - 0.2 |d theta| (README.md:52,99); theta = heading of the move
            prev_move = self.last_move
            turned = (a != 0) & (prev_move != 0) & (a != prev_move)
            same_axis = (a <= 2) == (prev_move <= 2)
            dtheta = np.where(turned, np.where(same_axis, np.pi, np.pi / 2), 0.0)
            self.last_move = np.where(a != 0, a, prev_move)
            dconc = current_conc - prev_conc
            total = dconc - 0.2 * dtheta
            conc_reward = dconc.astype(np.float32)
            explore_reward = np.zeros(self.n, dtype=np.float32)
            tke_term = np.zeros(self.n, dtype=np.float32)
            move_penalty = -(0.2 * dtheta)
            boundary_penalty = np.zeros(self.n)

        dx = self.pos32[:, 0].astype(np.float64) - self.src[:, 0]
        dy = self.pos32[:, 1].astype(np.float64) - self.src[:, 1]
        distance = np.sqrt(dx * dx + dy * dy)                                        # :155
        reached = distance <= self.current_radius                                    # :156
        bonus = np.minimum(500.0, 150 * (cfg.initial_radius / self.current_radius))  # :158
        total = np.where(reached, total + bonus, total)

        done = (self.step_count >= cfg.max_steps) | reached                          # :161
        self.last_reached = reached
        info = {
            "concentration_reward": conc_reward,
            "explore_reward": explore_reward,
            "move_penalty": move_penalty,
            "tke_penalty": -tke_term,
            "boundary_penalty": boundary_penalty,
            "reached": reached,
            "current_conc": current_conc,
            "conc_gradient": conc_gradient,
            "cell64": np.stack([cx, cy], axis=1),
        }
        return obs, total, done, info


# --------------------------------------------------------------------------------------
# scalar environment -- timed CPU baseline ("port" of the reference loop)
# --------------------------------------------------------------------------------------
class OracleScalarEnv:
    """One environment with python/numpy scalars: the same sequence of small numpy calls
    the reference executes per step, used to time the CPU baseline.  Draws its noise from
    a ``numpy.random.Generator`` (the reference uses the unseeded global stream)."""

    def __init__(self, cfg: PlumeConfig, rng: np.random.Generator):
        self.cfg = cfg
        self.rng = rng
        self.current_radius = cfg.initial_radius
        self.explore_bonus = cfg.explore_bonus
        self.visited: dict = {}
        self.reached = False
        self.reset()

    def reset(self):
        cfg = self.cfg
        self.source_pos = self.rng.random(2) * (cfg.grid_size - 100) + 50
        z = self.rng.standard_normal((cfg.grid_size, cfg.grid_size))
        u = self.rng.random((cfg.grid_size, cfg.grid_size))
        self.conc_field, self.tke_field = plume_fields(cfg, self.source_pos, z, u)
        self.agent_pos = np.array([0.0, 0.0])
        self.step_count = 0
        self.visited.clear()
        return self._obs()

    def _obs(self):
        cfg = self.cfg
        x = min(max(int(self.agent_pos[0]), 0), cfg.grid_size - 1)
        y = min(max(int(self.agent_pos[1]), 0), cfg.grid_size - 1)
        level = min(self.visited.get((x // cfg.cell_size, y // cfg.cell_size), 0) / 5.0, 1.0)
        return np.array([self.agent_pos[0] / cfg.grid_size, self.agent_pos[1] / cfg.grid_size,
                         self.conc_field[x, y] / cfg.conc_peak,
                         self.tke_field[x, y] / (cfg.turbulence_intensity * 3),
                         self.step_count / cfg.max_steps, level], dtype=np.float32)

    def step(self, action: int):
        cfg = self.cfg
        G = cfg.grid_size
        self.step_count += 1
        px = min(max(int(self.agent_pos[0]), 0), G - 1)
        py = min(max(int(self.agent_pos[1]), 0), G - 1)
        prev_conc = self.conc_field[px, py] / cfg.conc_peak
        ms = cfg.move_step
        d = MOVE_TABLE[action] * ms
        dnorm = np.sqrt(d[0] * d[0] + d[1] * d[1])          # np.float64, like np.linalg.norm (keeps the f64 promotion)
        move_penalty = -0.15 * (1 - dnorm / ms)
        eff = (ms * 0.2) * (self.rng.standard_normal(2) * self.tke_field[px, py] / (cfg.turbulence_intensity * 3))
        new_pos = np.clip(self.agent_pos + d + eff, 0, cfg.clip_hi)
        self.agent_pos = new_pos.astype(np.float32)
        cx = min(max(int(new_pos[0]), 0), G - 1)
        cy = min(max(int(new_pos[1]), 0), G - 1)
        grad = (self.conc_field[cx, cy] / cfg.conc_peak - prev_conc) / (dnorm + 1e-6)
        bd = min(new_pos[0] / G, (G - new_pos[0]) / G, new_pos[1] / G, (G - new_pos[1]) / G)
        bpen = -cfg.boundary_penalty * (cfg.boundary_decay_start - bd) ** 2 \
            if (bd < cfg.boundary_decay_start and grad < -0.01) else 0
        key = (int(new_pos[0] // cfg.cell_size), int(new_pos[1] // cfg.cell_size))
        self.visited[key] = self.visited.get(key, 0) + 1
        vc = self.visited[key]
        explore = (self.explore_bonus * (1 - self._obs()[5])) / (vc ** 0.75 + 1)
        obs = self._obs()
        total = cfg.conc_reward_coef * obs[2] + explore + move_penalty - cfg.tke_penalty_factor * obs[3] + bpen
        dist = float(np.sqrt(np.sum((self.agent_pos - self.source_pos) ** 2)))
        self.reached = dist <= self.current_radius
        if self.reached:
            total += min(500, 150 * (cfg.initial_radius / self.current_radius))
        done = self.step_count >= cfg.max_steps or self.reached
        info = {"concentration_reward": cfg.conc_reward_coef * obs[2], "explore_reward": explore,
                "move_penalty": move_penalty, "tke_penalty": -cfg.tke_penalty_factor * obs[3],
                "boundary_penalty": bpen}
        return obs, total, done, info
