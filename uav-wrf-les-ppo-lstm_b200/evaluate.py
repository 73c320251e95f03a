"""Batched evaluators: the three reference evaluation drivers with N environments in lockstep.

* ``stop="lstm"``      PPOV2.1/evaluate_with_lstm.py:53-113 -- greedy policy, LSTM(1->32) stop head over the
                       last 20 concentrations, stop when ``stop_prob > 0.8``; success = deviation <= 50.
* ``stop="threshold"`` PPOV2.0/evaluate_with_lstm.py:10-37,70-113 -- ``ThresholdController``: every 10th step
                       the 3x128 LSTM predicts the source concentration from the MinMax-scaled last 10
                       samples, threshold = 0.95 * prediction; stop when the concentration or the mean of the
                       last 10 reaches it (from step 20 on); success = deviation <= 40.
* ``stop="fixed"``     PPOV1.1/evaluate_model.py:25-37,44-82 -- position std over the last 10 steps < 2 px and
                       the (sic, twice-scaled) concentration above 0.8 * CONC_PEAK; 2000-step guard;
                       success = deviation < current_radius.

All envs of a round start together and run without auto-reset, so every live env is at the same step and the
"every 10th step" / "from step 20" conditions are uniform; a finished env is masked out.  The policy forward,
the env step, the field accessor and the LSTM heads are the library's kernels; the few ``[N, 10]`` ring-buffer
updates in between are torch ops (evaluator glue, not the training hot path).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .config import config_for
from .env import VecMethaneEnv


@dataclass
class EvalResult:
    deviations: torch.Tensor       # [E] ||final agent_pos - source_pos||
    steps: torch.Tensor            # [E] steps taken
    success: torch.Tensor          # [E] bool
    stopped_early: torch.Tensor    # [E] bool: ended by the stop test
    stop_step: torch.Tensor        # [E] step at which the stop test fired (0 = never)

    def summary(self) -> dict:
        d = self.deviations.double()
        ok = self.success
        return {"episodes": int(d.numel()), "mean_deviation": float(d.mean()), "std_deviation": float(d.std(unbiased=False)),
                "success_rate": float(ok.double().mean()),
                "success_mean_deviation": float(d[ok].mean()) if bool(ok.any()) else 0.0,
                "stopped_early_rate": float(self.stopped_early.double().mean()),
                "mean_steps": float(self.steps.double().mean())}


class ThresholdController:
    """Batched ``ThresholdController`` (PPOV2.0/evaluate_with_lstm.py:10-37).  ``scaler`` = (data_min,
    data_max) of the fitted ``MinMaxScaler`` (the reference refits it on the single saved ``data_min_``,
    :52-56, which makes the transform ``x - data_min``)."""

    def __init__(self, model, scaler=(0.0, 0.0), num_envs: int = 1, window_size: int = 10, device="cuda"):
        self.model, self.window_size = model, int(window_size)
        self.min_activate_steps = 2 * self.window_size
        rng = float(scaler[1]) - float(scaler[0])
        self.scale = 1.0 / (rng if rng != 0.0 else 1.0)                  # sklearn _handle_zeros_in_scale
        self.offset = -float(scaler[0]) * self.scale
        dev = torch.device(device)
        self.current_threshold = torch.full((num_envs,), float("nan"), dtype=torch.float64, device=dev)
        self.conc_buffer = torch.zeros(num_envs, self.window_size, dtype=torch.float64, device=dev)
        self.count = 0

    def reset(self) -> None:
        self.current_threshold.fill_(float("nan"))
        self.conc_buffer.zero_()
        self.count = 0

    def push(self, current_conc: torch.Tensor) -> None:
        self.conc_buffer = torch.roll(self.conc_buffer, -1, dims=1)
        self.conc_buffer[:, -1] = current_conc
        self.count += 1

    def update_threshold(self) -> None:
        """``trajectory[-window:]`` is the same data as the conc buffer once ``window`` samples exist."""
        if self.count >= max(self.window_size, self.min_activate_steps):
            scaled = (self.conc_buffer * self.scale + self.offset).float()          # FloatTensor(scaled)
            pred = self.model(scaled.unsqueeze(-1), lengths=[self.window_size] * scaled.shape[0])
            self.current_threshold = pred.double() * 0.95

    def should_stop(self, current_conc: torch.Tensor, step_count: int) -> torch.Tensor:
        n = min(self.count, self.window_size)
        mean = self.conc_buffer[:, self.window_size - n:].mean(dim=1)
        has = ~torch.isnan(self.current_threshold)
        thr = torch.where(has, self.current_threshold, torch.zeros_like(self.current_threshold))
        return has & (step_count >= self.min_activate_steps) & ((current_conc >= thr) | (mean >= thr))


@torch.no_grad()
def evaluate_policy(model, stop: str | None = "lstm", head=None, version: str | None = None, num_envs: int = 1024,
                    rounds: int = 1, seed: int = 0, scaler=(0.0, 0.0), device="cuda", env: VecMethaneEnv | None = None,
                    trace: dict | None = None, plume_model: str = "isotropic") -> EvalResult:
    """Runs ``rounds`` x ``num_envs`` greedy evaluation episodes.  ``head``: ``PeakAndStopPredictor`` for
    ``stop="lstm"``, ``ConcentrationThresholdPredictor`` for ``stop="threshold"``.  ``trace`` (optional
    dict) receives the per-step tensors of the first round (positions, concentrations, stop flags) for
    parity tests."""
    version = version or {"lstm": "2.1", "threshold": "2.0", "fixed": "1.1", None: "2.1"}[stop]
    cfg = config_for(version)
    dev = torch.device(device)
    if env is None:
        env = VecMethaneEnv(num_envs, device=dev, version=version, seed=seed, field_mode="procedural",
                            plume_model=plume_model)
    N = env.num_envs
    max_steps = 2000 if stop == "fixed" else cfg.max_steps                  # PPOV1.1/evaluate_model.py:52
    window = {"lstm": cfg.lstm_window, "threshold": 10, "fixed": 10, None: 1}[stop]
    out = {k: [] for k in ("deviations", "steps", "success", "stopped_early", "stop_step")}
    ctrl = ThresholdController(head, scaler, N, 10, dev) if stop == "threshold" else None
    for rnd in range(rounds):
        obs = env.reset().clone()
        if ctrl is not None:
            ctrl.reset()
        alive = torch.ones(N, dtype=torch.bool, device=dev)
        steps = torch.zeros(N, dtype=torch.int32, device=dev)
        early = torch.zeros(N, dtype=torch.bool, device=dev)
        stop_step = torch.zeros(N, dtype=torch.int32, device=dev)
        final_pos = torch.zeros(N, 2, dtype=torch.float32, device=dev)
        conc_win = torch.zeros(N, window, dtype=torch.float64, device=dev)
        pos_win = torch.zeros(N, 10, 2, dtype=torch.float32, device=dev)
        for step in range(1, max_steps + 1):
            action, _, _, _ = model.act(obs, env=env, greedy=True)                    # argmax, evaluate_with_lstm.py:65
            obs, _, done, info = env.step(action)
            obs = obs.clone()
            pos = env.agent_pos
            conc = env.conc_at(pos[:, 0].int(), pos[:, 1].int())             # conc_field[int(x), int(y)]
            stop_now = torch.zeros(N, dtype=torch.bool, device=dev)
            if stop == "lstm":
                conc_win = torch.roll(conc_win, -1, dims=1)
                conc_win[:, -1] = conc
                if step >= window:                                           # evaluate_with_lstm.py:73-80
                    _, prob = head((conc_win / 100.0).float())
                    stop_now = prob > cfg.lstm_stop_threshold
            elif stop == "threshold":
                ctrl.push(conc)
                if step % 10 == 0:                                           # PPOV2.0/evaluate_with_lstm.py:89-90
                    ctrl.update_threshold()
                stop_now = ctrl.should_stop(conc, step)
            elif stop == "fixed":
                pos_win = torch.roll(pos_win, -1, dims=1)
                pos_win[:, -1] = pos
                if step >= 10:                                               # PPOV1.1/evaluate_model.py:25-37
                    pos_std = pos_win.std(dim=1, unbiased=False).mean(dim=1)
                    current = info["concentration_reward"].double() * cfg.conc_peak * cfg.conc_peak
                    stop_now = (pos_std < 2.0) & (current > 0.8 * cfg.conc_peak)
            if trace is not None and rnd == 0:
                for k, v in (("pos", pos), ("conc", conc), ("stop", stop_now), ("done", done), ("alive", alive),
                             ("conc_reward", info["concentration_reward"]), ("action", action)):
                    trace.setdefault(k, []).append(v.clone())
            finish = alive & (done | stop_now)
            steps = torch.where(finish, torch.full_like(steps, step), steps)
            early = torch.where(finish, stop_now, early)
            stop_step = torch.where(finish & stop_now, torch.full_like(stop_step, step), stop_step)
            final_pos = torch.where(finish.unsqueeze(1), pos, final_pos)
            alive = alive & ~finish
            if step % 16 == 0 and not bool(alive.any()):
                break
        # envs still running at the guard count as finished there
        pos = env.agent_pos
        steps = torch.where(alive, torch.full_like(steps, max_steps), steps)
        final_pos = torch.where(alive.unsqueeze(1), pos, final_pos)
        dev_ = (final_pos.double() - env.source_pos).norm(dim=1)
        if stop == "fixed":
            ok = dev_ < env.radius_t                                         # PPOV1.1/evaluate_model.py:75
        else:
            ok = dev_ <= cfg.success_distance_threshold
        for k, v in (("deviations", dev_), ("steps", steps), ("success", ok), ("stopped_early", early),
                     ("stop_step", stop_step)):
            out[k].append(v)
    return EvalResult(**{k: torch.cat(v) for k, v in out.items()})
