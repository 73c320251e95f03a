"""Constants of the reference versions (PPOV*/config.py) as frozen dataclasses.

The reference keeps them as module-level names edited by hand (config.py:6-44); the
batched implementation passes one immutable object to the kernels instead.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

FIELD_PROCEDURAL, FIELD_F32, FIELD_F64 = 0, 1, 2
FIELD_MODES = {"procedural": FIELD_PROCEDURAL, "f32": FIELD_F32, "materialised": FIELD_F64, "f64": FIELD_F64}


@dataclass(frozen=True)
class PlumeConfig:
    version: str = "2.1"
    # config.py:6-9
    grid_size: int = 500
    max_steps: int = 1000
    conc_peak: float = 100.0
    turbulence_intensity: float = 3.0
    # config.py:12 (V2.1); V2.0/V1.1 use grid/16 (PPOV2.0/environment.py:54)
    sigma: float = 15.0
    # environment.py:112; V1.1 clips at grid-1e-6 (PPOV1.1/environment.py:105)
    clip_hi: float = 499.0
    # config.py:16-22
    gamma: float = 0.99
    lam: float = 0.95
    clip_epsilon: float = 0.2
    entropy_beta: float = 0.01
    learning_rate: float = 3e-5
    batch_size: int = 256
    epochs: int = 5
    # config.py:25-27
    explore_bonus: float = 0.6
    decay_factor: float = 0.999
    grid_divisions: int = 10
    # config.py:31-35
    initial_radius: float = 50.0
    min_radius: float = 5.0
    radius_decay: float = 0.9
    success_threshold: float = 0.6
    window_size: int = 120
    # config.py:38-41
    conc_reward_coef: float = 2.0
    tke_penalty_factor: float = 0.4
    boundary_penalty: float = 0.1
    boundary_decay_start: float = 0.15
    # evaluators: PPOV2.0/config.py:43-44, PPOV2.1/evaluate_with_lstm.py:42,77,88
    success_distance_threshold: float = 50.0
    lstm_window: int = 20
    lstm_stop_threshold: float = 0.8
    max_grad_norm: float = 0.5          # train_ppo2.0.py:86

    @property
    def cell_size(self) -> int:
        return self.grid_size // self.grid_divisions


def config_for(version: str = "2.1") -> PlumeConfig:
    """The constants of PPOV1.1 / PPOV2.0 / PPOV2.1."""
    base = PlumeConfig()
    v = str(version)
    if v == "2.1":
        return base
    if v == "2.0":
        return replace(base, version="2.0", sigma=base.grid_size / 16, success_distance_threshold=40.0,
                       lstm_window=10)
    if v == "1.1":
        return replace(base, version="1.1", sigma=base.grid_size / 16, max_steps=5000,
                       clip_hi=base.grid_size - 1e-6)
    raise ValueError(f"unknown reference version {version!r} (expected 1.1, 2.0 or 2.1)")
