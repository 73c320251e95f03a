"""Fused rollout driver: T lockstep iterations of policy + env step (+ LSTM stop head) in ONE
persistent kernel launch (csrc/rollout_kernel.cu), replacing the reference's python
``while not done`` loop (PPOV2.1/train_ppo2.0.py:156-192, evaluate_with_lstm.py:61-82).

Where the stop decision does not end the episode (training rollouts: the reference only lets the
stop head terminate in evaluate_with_lstm.py:77-80) the head does not feed back into the loop, so
it is evaluated for the whole ``[T, N]`` segment by one throughput kernel afterwards
(``plume_stop_head_segment``) -- same windows, same arithmetic, identical results."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .buffer import PPOBuffer
from .config import FIELD_PROCEDURAL


class RolloutEngine:
    def __init__(self, env, model, stop_head=None, horizon: int = 256, with_info: bool = False,
                 with_trend: bool = False, with_trajectory: bool = False, stop_head_path: str = "auto"):
        if env.field_mode != FIELD_PROCEDURAL:
            raise ValueError("the fused rollout needs field_mode='procedural'")
        self.env, self.model, self.stop_head = env, model, stop_head
        self.lib = _lib.load()
        self.horizon = int(horizon)
        # which kernel family evaluates the deferred stop head: "auto" (tcgen05 where the hidden size has a
        # tensor-core kernel), "tensor", "simt" (bit-identical to the in-loop head)
        self.stop_head_path = _lib.KERNEL_PATHS[stop_head_path]
        self.after_loop = None        # optional callable run between the lockstep kernel and the stop-head kernel
        dev, N = env.device, env.num_envs
        self.window = env.cfg.lstm_window
        self.buffer = PPOBuffer(horizon, N, dev, with_info=with_info, with_stop=stop_head is not None,
                                with_trend=with_trend, with_trajectory=with_trajectory)
        self.conc_window = torch.zeros(N, self.window, dtype=torch.float32, device=dev)
        self._window_next = torch.zeros_like(self.conc_window)      # ring written by the deferred head
        self.window_fill = torch.zeros(N, dtype=torch.int32, device=dev)
        self.last_obs = torch.zeros(N, _lib.OBS_DIM, dtype=torch.float32, device=dev)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.launches = 0

    def reset_windows(self) -> None:
        self.window_fill.zero_()

    @torch.no_grad()
    def collect(self, greedy: bool = False, stop_terminates: bool = False, forced_actions=None,
                step_noise=None, noise_out=None, horizon: int | None = None,
                defer_stop_head: bool | None = None, stop_mode: str | None = None, step_guard: int = 0,
                eval_ring=None, stop_threshold=None) -> PPOBuffer:
        """Runs ``horizon`` lockstep iterations and returns the filled buffer (asynchronous).
        ``defer_stop_head`` (default: whenever the stop decision does not terminate episodes)
        evaluates the LSTM head and the trend features after the loop in one batched kernel.
        ``stop_mode`` ("fixed" / "threshold"; engines without a stop head): the V1.1 / V2.0 evaluator stop test runs
        inside the kernel and ends the episode; ``eval_ring`` [N,10] float64 carries the last 10 samples across
        segments, ``stop_threshold`` [N] float64 holds the V2.0 controller's current thresholds, ``step_guard`` ends
        an episode at that step count (evaluate_model.py:52)."""
        env, T = self.env, int(horizon or self.horizon)
        assert T <= self.horizon
        defer = (self.stop_head is not None and not stop_terminates) if defer_stop_head is None else bool(defer_stop_head)
        if defer and (self.stop_head is None or stop_terminates):
            raise ValueError("defer_stop_head needs a stop head whose decision does not end the episode")
        flags = _lib.FLAG_AUTO_RESET
        flags |= _lib.FLAG_DEFER_STOP_HEAD if defer else 0
        flags |= _lib.FLAG_FAST_REWARD if getattr(env, "fast_reward", False) else 0
        flags |= _lib.FLAG_GREEDY if greedy else 0
        flags |= _lib.FLAG_STOP_TERMINATES if stop_terminates else 0
        if stop_mode is not None:
            if self.stop_head is not None:
                raise ValueError("stop_mode needs an engine without an LSTM stop head")
            flags |= {"fixed": _lib.FLAG_STOP_FIXED, "threshold": _lib.FLAG_STOP_THRESHOLD}[stop_mode]
            if self.buffer.stop_flag is None:        # the kernel's stop decisions go where the LSTM head's would
                self.buffer.stop_flag = torch.zeros(self.horizon, env.num_envs, dtype=torch.uint8, device=env.device)
        if not 0 <= int(step_guard) < 65536:
            raise ValueError("step_guard must fit 16 bits")
        flags |= int(step_guard) << 16
        if forced_actions is not None:
            forced_actions = forced_actions.to(device=env.device, dtype=torch.int32).contiguous()
        if step_noise is not None:
            step_noise = step_noise.to(device=env.device, dtype=torch.float64).contiguous()
        bufs = self.buffer.c_rollout_buffers(self.conc_window, self.window_fill, self.last_obs, forced_actions,
                                             step_noise, noise_out, eval_ring, stop_threshold)
        if self.stop_head is not None:
            lp = self.stop_head.c_params(self.window, env.cfg.lstm_stop_threshold)
        else:
            lp = _lib.LstmParams()
            lp.window = self.window
        with torch.cuda.device(env.device):
            rc = self.lib.plume_rollout(C.byref(env.c_config), C.byref(env.c_state), self.model.flat.data_ptr(),
                                        C.byref(lp), C.byref(bufs), T, flags, self.nan_flag.data_ptr(),
                                        torch.cuda.current_stream(env.device).cuda_stream)
            _lib.check(rc, "plume_rollout")
            self.launches += 1
            if self.after_loop is not None:
                self.after_loop(self.buffer, T)
            if defer:
                b = self.buffer
                rc = self.lib.plume_stop_head_segment(C.byref(lp), b.conc_sample.data_ptr(), b.fill_t.data_ptr(),
                                                      _lib.ptr(b.src_dist), T, env.num_envs,
                                                      self.conc_window.data_ptr(), self._window_next.data_ptr(),
                                                      env.cfg.conc_peak, b.stop_prob.data_ptr(),
                                                      b.stop_flag.data_ptr(), b.peak_pred.data_ptr(),
                                                      _lib.ptr(b.trend), self.stop_head_path,
                                                      torch.cuda.current_stream(env.device).cuda_stream)
                _lib.check(rc, "plume_stop_head_segment")
                self.conc_window, self._window_next = self._window_next, self.conc_window
                self.launches += 1
        self.buffer.filled = T
        self.buffer.flag_code_valid = True
        return self.buffer

    def check_nan(self) -> None:
        """model.py:41-43: raise if any logit was NaN during the last rollout(s)."""
        if int(self.nan_flag.item()) != 0:
            self.nan_flag.zero_()
            raise RuntimeError("NaN in model output")
