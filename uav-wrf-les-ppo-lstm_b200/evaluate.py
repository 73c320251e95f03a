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

All envs of a round start together, so every env's first episode is at the lockstep step count and the
"every 10th step" / "from step 20" conditions are uniform.  Policy forward, env step and the stop test of every step
run inside the fused rollout kernel (``plume_rollout`` with ``PLUME_FLAG_STOP_*`` / the in-loop LSTM head), the
bookkeeping of where each env's episode ended in ``plume_eval_collect``; nothing returns to Python per step.
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
    parity tests.

    A round is a sequence of FUSED rollout segments (policy + env step + stop test inside ``plume_rollout``, greedy
    actions): all envs start their episode together, every env's FIRST episode of the round is what counts
    (``plume_eval_collect`` records where it ends), and the round is over when every env has finished one -- envs
    that finish earlier are auto-reset and keep running, their later transitions are ignored.  ``stop="lstm"`` runs
    the in-loop LSTM head, ``"fixed"`` / ``"threshold"`` the kernel's own stop tests; the V2.0 controller's 3 x 128
    LSTM is evaluated for all envs at once between the 10-step segments (its "every 10th step")."""
    import ctypes as C

    from . import _lib
    from .rollout import RolloutEngine
    version = version or {"lstm": "2.1", "threshold": "2.0", "fixed": "1.1", None: "2.1"}[stop]
    cfg = config_for(version)
    dev = torch.device(device)
    if env is None:
        env = VecMethaneEnv(num_envs, device=dev, version=version, seed=seed, field_mode="procedural",
                            plume_model=plume_model, auto_reset=True)
    elif not env.auto_reset:
        raise ValueError("evaluate_policy drives the fused rollout: the env must be built with auto_reset=True")
    N = env.num_envs
    lib = _lib.load()
    max_steps = 2000 if stop == "fixed" else cfg.max_steps                  # PPOV1.1/evaluate_model.py:52
    seg = {"lstm": 50, "threshold": 10, "fixed": 50, None: 50}[stop]
    eng = RolloutEngine(env, model, head if stop == "lstm" else None, horizon=seg, with_info=True, with_trajectory=True)
    ctrl = ThresholdController(head, scaler, N, 10, dev) if stop == "threshold" else None
    ring = torch.zeros(N, 10, dtype=torch.float64, device=dev) if stop in ("fixed", "threshold") else None
    out = {k: [] for k in ("deviations", "steps", "success", "stopped_early", "stop_step")}
    stream = lambda: torch.cuda.current_stream(dev).cuda_stream
    for rnd in range(rounds):
        env.reset()
        eng.reset_windows()
        finished = torch.zeros(N, dtype=torch.uint8, device=dev)
        steps = torch.zeros(N, dtype=torch.int32, device=dev)
        early = torch.zeros(N, dtype=torch.uint8, device=dev)
        stop_step = torch.zeros(N, dtype=torch.int32, device=dev)
        deviation = torch.zeros(N, dtype=torch.float64, device=dev)
        radius_at_start = env.radius_t.clone()
        thr = torch.full((N,), float("nan"), dtype=torch.float64, device=dev) if ctrl is not None else None
        if ring is not None:
            ring.zero_()
        base = 0
        while base < max_steps:
            T = min(seg, max_steps - base)
            buf = eng.collect(greedy=True, horizon=T, stop_terminates=(stop == "lstm"),
                              defer_stop_head=False if stop == "lstm" else None,
                              stop_mode=stop if stop in ("fixed", "threshold") else None,
                              step_guard=max_steps if stop == "fixed" else 0, eval_ring=ring, stop_threshold=thr)
            alive_before = finished == 0
            pending = None
            if ctrl is not None and base + T >= ctrl.min_activate_steps:       # update_threshold at every 10th step
                scaled = (ring * ctrl.scale + ctrl.offset).float()             # FloatTensor(scaler.transform(window))
                thr = ctrl.model(scaled.unsqueeze(-1), lengths=[ctrl.window_size] * N).double() * 0.95
                pending = thr
            cb = buf.c_rollout_buffers(None, None, None, eval_ring=ring)
            with torch.cuda.device(dev):
                _lib.check(lib.plume_eval_collect(C.byref(cb), T, N, base, finished.data_ptr(), steps.data_ptr(),
                                                  early.data_ptr(), stop_step.data_ptr(), deviation.data_ptr(),
                                                  _lib.ptr(pending), stream()), "plume_eval_collect")
            if trace is not None and rnd == 0:
                trace.setdefault("segments", []).append({
                    "pos": buf.pos_out[:T].clone(), "conc": buf.conc_out[:T].double(),
                    "stop": (buf.stop_flag[:T] != 0) if buf.stop_flag is not None else torch.zeros(T, N, dtype=torch.bool, device=dev),
                    "done": buf.dones[:T] != 0, "conc_reward": buf.info[:T, 0].clone(), "action": buf.actions[:T].clone(),
                    "alive_at_start": alive_before.clone()})
            base += T
            if bool((finished != 0).all()):
                break
        # envs still running at the guard count as finished there (the env's own MAX_STEPS / the 2000-step guard end
        # every first episode by then; this only covers a guard that is not a multiple of the segment length)
        still = finished == 0
        if bool(still.any()):
            steps = torch.where(still, torch.full_like(steps, base), steps)
            deviation = torch.where(still, (env.agent_pos.double() - env.source_pos).norm(dim=1), deviation)
        if stop == "fixed":
            ok = deviation < radius_at_start                                 # PPOV1.1/evaluate_model.py:75
        else:
            ok = deviation <= cfg.success_distance_threshold
        for k, v in (("deviations", deviation), ("steps", steps), ("success", ok), ("stopped_early", early != 0),
                     ("stop_step", stop_step)):
            out[k].append(v)
    if trace is not None and "segments" in trace:
        segs = trace.pop("segments")
        for k in ("pos", "conc", "stop", "done", "conc_reward", "action"):
            trace[k] = list(torch.cat([s_[k] for s_ in segs]).unbind(0))
    return EvalResult(**{k: torch.cat(v) for k, v in out.items()})
