"""Device-resident rollout buffer: the batched form of ``PPOBuffer`` (PPOV2.1/model.py:132-173).

The reference appends python scalars to six lists and converts them to CPU tensors in
``get()``.  Here the six arrays live in HBM as ``[T, N]`` struct-of-arrays (step-major, env
fastest) which is what the rollout kernel writes, what the GAE scan reads along ``T`` and what
the minibatch gather indexes as a flat ``[T*N]`` set.  ``store/get/clear/len(states)`` keep the
reference's call pattern (train_ppo2.0.py:182-191)."""
from __future__ import annotations

import torch

from . import _lib


class PPOBuffer:
    def __init__(self, horizon: int = 256, num_envs: int = 1, device="cuda", with_info: bool = False,
                 with_stop: bool = True, with_trend: bool = False, with_episode: bool = True,
                 with_trajectory: bool = False):
        self.horizon, self.num_envs, self.device = int(horizon), int(num_envs), torch.device(device)
        T, N, dev = self.horizon, self.num_envs, self.device
        z = lambda dt, *s: torch.zeros(*s, dtype=dt, device=dev)
        self.obs = z(torch.float32, T, N, _lib.OBS_DIM)
        self.actions = z(torch.int32, T, N)
        self.rewards = z(torch.float32, T, N)
        self.values = z(torch.float32, T, N)
        self.log_probs = z(torch.float32, T, N)
        self.dones = z(torch.float32, T, N)
        self.reached = z(torch.uint8, T, N)
        self.flag_code = z(torch.uint8, T, N)        # bit 0 done, bit 1 reached (curriculum input)
        self.flag_code_valid = False                  # written by the rollout kernel (not by store())
        self.stop_prob = z(torch.float32, T, N) if with_stop else None
        self.stop_flag = z(torch.uint8, T, N) if with_stop else None
        self.peak_pred = z(torch.float32, T, N) if with_stop else None
        self.trend = z(torch.float32, T, N, 4) if with_trend else None
        self.info = z(torch.float32, T, _lib.INFO_DIM, N) if with_info else None
        self.episode_idx = z(torch.int32, T, N) if with_episode else None
        # inputs of the deferred stop head (csrc/lstm_kernels.cu::stop_head_segment_kernel)
        self.conc_sample = z(torch.float32, T, N) if with_stop else None
        self.fill_t = z(torch.uint8, T, N) if with_stop else None
        self.src_dist = z(torch.float64, T, N) if ((with_stop and with_trend) or with_trajectory) else None
        # trajectory logging (trajectory_log.TrajectoryLogger): post-step positions, source at the episode's last step
        self.pos_out = z(torch.float32, T, N, 2) if with_trajectory else None
        self.src_out = z(torch.float32, T, N, 2) if with_trajectory else None
        self.conc_out = z(torch.float32, T, N) if with_trajectory else None
        self.advantages = z(torch.float32, T, N)
        self.returns = z(torch.float32, T, N)
        self.filled = 0          # rows written
        self.reached_stored = True    # False once a row was stored without its `reached` flags

    # -- reference API ----------------------------------------------------------------------
    def clear(self) -> None:
        self.filled = 0
        self.flag_code_valid = False
        self.reached_stored = True

    def store(self, state, action, reward, value, log_prob, done, reached=None, info=None, pos=None, conc=None,
              src=None) -> None:
        """Appends one lockstep row (each argument ``[N]``-shaped, ``state`` ``[N,6]``).  ``reached`` (optional, the
        env's ``info["reached"]``) feeds the batched curriculum (``PPOTrainer.update_from_rollout``); the
        reference's six-argument call leaves it at "not reached" and drives ``PPOTrainer.update(success)`` itself.
        ``info`` (the step's info dict), ``pos`` (``env.agent_pos``), ``conc`` (``env.conc_field[x, y]``) and ``src``
        (``env.source_pos``) are what the reference driver logs per step (train_ppo2.0.py:166-180); a buffer built
        with ``with_info`` / ``with_trajectory`` keeps them for ``TrajectoryLogger.consume``."""
        t = self.filled
        if t >= self.horizon:
            raise IndexError("PPOBuffer is full: call clear() after update_model()")
        dev = self.device
        self.obs[t] = torch.as_tensor(state, dtype=torch.float32, device=dev).reshape(self.num_envs, -1)
        self.actions[t] = torch.as_tensor(action, device=dev).to(torch.int32).reshape(self.num_envs)
        self.rewards[t] = torch.as_tensor(reward, device=dev).to(torch.float32).reshape(self.num_envs)
        self.values[t] = torch.as_tensor(value, device=dev).to(torch.float32).reshape(self.num_envs)
        self.log_probs[t] = torch.as_tensor(log_prob, device=dev).to(torch.float32).reshape(self.num_envs)
        self.dones[t] = torch.as_tensor(done, device=dev).to(torch.float32).reshape(self.num_envs)
        if reached is None:
            self.reached[t].zero_()
            self.reached_stored = False
        else:
            self.reached[t] = torch.as_tensor(reached, device=dev).to(torch.uint8).reshape(self.num_envs)
        N = self.num_envs
        if info is not None and self.info is not None:
            for k, key in enumerate(_lib.INFO_KEYS):
                self.info[t, k] = torch.as_tensor(info[key], device=dev).to(torch.float32).reshape(N)
        if pos is not None and self.pos_out is not None:
            self.pos_out[t] = torch.as_tensor(pos, device=dev).to(torch.float32).reshape(N, 2)
        if conc is not None and self.conc_out is not None:
            self.conc_out[t] = torch.as_tensor(conc, device=dev).to(torch.float32).reshape(N)
        if src is not None and self.src_out is not None:
            self.src_out[t] = torch.as_tensor(src, device=dev).to(torch.float32).reshape(N, 2)
        self.flag_code_valid = False
        self.filled = t + 1

    @property
    def states(self) -> torch.Tensor:
        """Flat ``[filled*N, 6]`` view; ``len(buffer.states)`` counts transitions like the
        reference's list (train_ppo2.0.py:189)."""
        return self.obs[: self.filled].reshape(-1, _lib.OBS_DIM)

    def get(self):
        """(states, actions, rewards, values, log_probs, dones) flattened step-major."""
        t = self.filled
        return (self.obs[:t].reshape(-1, _lib.OBS_DIM), self.actions[:t].reshape(-1).long(),
                self.rewards[:t].reshape(-1), self.values[:t].reshape(-1), self.log_probs[:t].reshape(-1),
                self.dones[:t].reshape(-1))

    def __len__(self) -> int:
        return self.filled * self.num_envs

    # -- C view -----------------------------------------------------------------------------
    def c_rollout_buffers(self, conc_window, window_fill, last_obs, forced_actions=None, step_noise=None,
                          noise_out=None, eval_ring=None, stop_threshold=None, t0: int = 0) -> _lib.RolloutBuffers:
        """``t0``: the segment starts at row ``t0`` of every ``[T][N]`` array (a rollout collected in several launches)."""
        p = _lib.ptr if t0 == 0 else (lambda x: _lib.ptr(None if x is None else x[t0:]))
        q = _lib.ptr
        return _lib.RolloutBuffers(p(self.obs), p(self.actions), p(self.rewards), p(self.values), p(self.log_probs),
                                   p(self.dones), p(self.reached), p(self.stop_prob), p(self.stop_flag),
                                   p(self.peak_pred), p(self.trend), p(self.info), p(self.episode_idx),
                                   p(forced_actions), p(step_noise), p(noise_out), q(conc_window), q(window_fill),
                                   q(last_obs), p(self.conc_sample), p(self.fill_t), p(self.src_dist),
                                   p(self.pos_out), p(self.src_out), p(self.conc_out), q(eval_ring), q(stop_threshold),
                                   p(self.flag_code))
