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


def overlap_split(horizon: int, num_envs: int, sm_count: int) -> tuple | None:
    """Row counts ``(h0, horizon - h0)`` of a rollout collected in two launches, or None where that does not pay.

    The lockstep kernel occupies one SM per 32-env tile and is latency-bound; the deferred stop head is a throughput
    kernel whose CTAs cannot share an SM with it (registers, shared memory).  With the head of the first ``h0`` rows on a
    second stream, its CTAs run on the ``idle = sm_count - tiles`` SMs while the lockstep kernel produces the other rows.
    ``h0`` is the largest multiple of 8 whose head finishes under them: per row the head costs 0.82 x the lockstep
    iteration at 4096 envs on all 148 SMs (5.9 us against 7.2 us, bench shape) and scales with the env count.  More
    launches do not help: a head waiting for SMs takes the ones a finished segment frees before the next segment starts,
    and every extra lockstep launch costs ~50 us of prologue (measured: ``profiles/r2_notes.md``)."""
    tiles = (num_envs + 31) // 32
    idle = sm_count - tiles
    if idle <= 0 or horizon < 32:
        return None
    h0 = int(horizon * idle / (idle + sm_count * 0.82 * num_envs / 4096.0)) // 8 * 8
    h0 = min(h0, horizon // 2)
    if h0 < 8 or h0 * 5.9 * num_envs / 4096.0 < 100.0:        # the head time hidden (us) against ~50 us for the second launch
        return None
    return (h0, horizon - h0)


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
        # ``overlap_chunks`` (row counts summing to the horizon, or None): the deferred-head rollout is launched as that
        # many shorter segments on a high-priority stream and the stop head of each follows on a second stream, so the
        # (MUFU-bound) head of the rows already written runs on the SMs the (latency-bound, <= 128-CTA) lockstep kernel
        # leaves idle.  Same kernels on the same rows: results are identical to the single launch.
        self.overlap_chunks = None
        self._streams = None
        self._head_pending = False

    def reset_windows(self) -> None:
        self.window_fill.zero_()

    @torch.no_grad()
    def collect(self, greedy: bool = False, stop_terminates: bool = False, forced_actions=None,
                step_noise=None, noise_out=None, horizon: int | None = None,
                defer_stop_head: bool | None = None, stop_mode: str | None = None, step_guard: int = 0,
                eval_ring=None, stop_threshold=None, join: bool = True, end_event=None) -> PPOBuffer:
        """Runs ``horizon`` lockstep iterations and returns the filled buffer (asynchronous).
        ``defer_stop_head`` (default: whenever the stop decision does not terminate episodes)
        evaluates the LSTM head and the trend features after the loop in one batched kernel.
        ``stop_mode`` ("fixed" / "threshold"; engines without a stop head): the V1.1 / V2.0 evaluator stop test runs
        inside the kernel and ends the episode; ``eval_ring`` [N,10] float64 carries the last 10 samples across
        segments, ``stop_threshold`` [N] float64 holds the V2.0 controller's current thresholds, ``step_guard`` ends
        an episode at that step count (evaluate_model.py:52).
        ``join=False`` (overlapped collection only): the caller's stream continues after the LOCKSTEP kernels -- rewards,
        values, flags are complete, the stop head's outputs are not until ``join_stop_head()``; work that does not read
        them (curriculum, GAE) then runs under the last stop-head launch.  ``end_event``: recorded behind the last launch
        of the collection, on whichever stream that is."""
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
        chunks = self.overlap_chunks if (defer and forced_actions is None and step_noise is None
                                         and noise_out is None) else None
        if chunks is not None and (len(chunks) < 2 or sum(chunks) != T or min(chunks) < 1):
            chunks = None
        with torch.cuda.device(env.device):
            self.join_stop_head()                 # a collection the caller left pending
            if chunks is not None:
                self._collect_overlapped(bufs, lp, chunks, flags, join, end_event)
                self.buffer.filled = T
                self.buffer.flag_code_valid = True
                return self.buffer
            rc = self.lib.plume_rollout(C.byref(env.c_config), C.byref(env.c_state), self.model.flat.data_ptr(),
                                        C.byref(lp), C.byref(bufs), T, flags, self.nan_flag.data_ptr(),
                                        torch.cuda.current_stream(env.device).cuda_stream)
            _lib.check(rc, "plume_rollout")
            self.launches += 1
            if self.after_loop is not None:
                self.after_loop(self.buffer, T)
            if defer:
                self._stop_head(lp, 0, T, torch.cuda.current_stream(env.device).cuda_stream)
            if end_event is not None:
                end_event.record(torch.cuda.current_stream(env.device))
        self.buffer.filled = T
        self.buffer.flag_code_valid = True
        return self.buffer

    def _stop_head(self, lp, t0: int, h: int, stream: int) -> None:
        """Deferred head of rows [t0, t0 + h): continues the window ring of the rows before it."""
        b, env = self.buffer, self.env
        o = (lambda x: _lib.ptr(None if x is None else x[t0:]))
        rc = self.lib.plume_stop_head_segment(C.byref(lp), o(b.conc_sample), o(b.fill_t), o(b.src_dist), h, env.num_envs,
                                              self.conc_window.data_ptr(), self._window_next.data_ptr(),
                                              env.cfg.conc_peak, o(b.stop_prob), o(b.stop_flag), o(b.peak_pred),
                                              o(b.trend), self.stop_head_path, stream)
        _lib.check(rc, "plume_stop_head_segment")
        self.conc_window, self._window_next = self._window_next, self.conc_window
        self.launches += 1

    def join_stop_head(self) -> None:
        """Makes the current stream wait for the stop head of a collection started with ``join=False``."""
        if self._head_pending:
            torch.cuda.current_stream(self.env.device).wait_event(self._streams[5])
            self._head_pending = False

    def _collect_overlapped(self, bufs0, lp, chunks, flags, join=True, end_event=None) -> None:
        env, dev = self.env, self.env.device
        if self._streams is None:
            lo, hi = torch.cuda.Stream.priority_range()
            self._streams = (torch.cuda.Stream(device=dev, priority=hi), torch.cuda.Stream(device=dev),
                             torch.cuda.Event(), [torch.cuda.Event() for _ in range(64)], torch.cuda.Event(),
                             torch.cuda.Event())
        loop_s, head_s, start, rows, loop_done, head_done = self._streams
        main = torch.cuda.current_stream(dev)
        start.record(main)
        loop_s.wait_event(start)
        head_s.wait_event(start)
        t0 = 0
        for k, h in enumerate(chunks):
            bufs = bufs0 if t0 == 0 else self.buffer.c_rollout_buffers(self.conc_window, self.window_fill,
                                                                       self.last_obs, t0=t0)
            rc = self.lib.plume_rollout(C.byref(env.c_config), C.byref(env.c_state), self.model.flat.data_ptr(),
                                        C.byref(lp), C.byref(bufs), h, flags, self.nan_flag.data_ptr(),
                                        loop_s.cuda_stream)
            _lib.check(rc, "plume_rollout")
            self.launches += 1
            rows[k % 64].record(loop_s)
            head_s.wait_event(rows[k % 64])
            self._stop_head(lp, t0, h, head_s.cuda_stream)
            t0 += h
        loop_done.record(loop_s)
        main.wait_event(loop_done)
        if self.after_loop is not None:
            self.after_loop(self.buffer, t0)
        head_done.record(head_s)
        if end_event is not None:
            end_event.record(head_s)
        self._head_pending = True
        if join:
            self.join_stop_head()

    def check_nan(self) -> None:
        """model.py:41-43: raise if any logit was NaN during the last rollout(s)."""
        if int(self.nan_flag.item()) != 0:
            self.nan_flag.zero_()
            raise RuntimeError("NaN in model output")
