"""CPU restatement of the reference learner, models and stop heads -- TEST INFRASTRUCTURE ONLY.

torch-CPU float32 is the arithmetic the reference itself runs here (it never moves the
PPO path to a GPU), so this oracle uses torch CPU ops where the reference does and numpy
where the reference does.  Citations are ``/root/reference/PPOV2.1/...:line``.

Pinned by ``tests/test_oracle_vs_reference.py`` (bit-equal parameters after a full
``_update_model`` on identical data and permutations; bit-equal forward outputs with the
reference modules loaded from the same ``state_dict``) and by ``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .plume_oracle import PlumeConfig


# --------------------------------------------------------------------------------------
# P4: actor-critic MLP, model.py:16-46
# --------------------------------------------------------------------------------------
class OracleActorCritic(nn.Module):
    """6 -> 256 -> LN -> ReLU -> 128 -> LN -> ReLU -> {actor 5, critic 1}; parameter names
    (``feature.0.weight`` ... ``critic.bias``) are the reference's state_dict keys."""

    def __init__(self, input_size: int = 6, output_size: int = 5):
        super().__init__()
        trunk = [nn.Linear(input_size, 256), nn.LayerNorm(256), nn.ReLU(),
                 nn.Linear(256, 128), nn.LayerNorm(128), nn.ReLU()]
        self.feature = nn.Sequential(*trunk)
        self.actor = nn.Linear(128, output_size)
        self.critic = nn.Linear(128, 1)
        for m, gain in ((self.feature[0], np.sqrt(2)), (self.feature[3], np.sqrt(2)),
                        (self.actor, 0.01), (self.critic, 1.0)):      # model.py:27-36
            nn.init.orthogonal_(m.weight, gain=gain)
            nn.init.constant_(m.bias, 0.0)

    def forward(self, x):
        h = self.feature(x)
        logits = self.actor(h)
        if torch.isnan(logits).any():                                   # model.py:41-43
            raise RuntimeError("NaN in model output")
        return torch.softmax(logits, dim=-1), self.critic(h)


def categorical_log_prob(probs: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """``Categorical(probs).log_prob(a)``: probs are re-normalised, clamped to
    [eps, 1-eps] and logged (torch/distributions/categorical.py, utils.probs_to_logits);
    call sites train_ppo2.0.py:63-64,161,185."""
    p = probs / probs.sum(-1, keepdim=True)
    eps = torch.finfo(p.dtype).eps
    logits = torch.log(p.clamp(min=eps, max=1 - eps))
    return logits.gather(-1, actions.long().unsqueeze(-1)).squeeze(-1)


# --------------------------------------------------------------------------------------
# P5: GAE + normalisation, train_ppo2.0.py:17-39
# --------------------------------------------------------------------------------------
def gae_quirk(rewards, values, dones, gamma: float, lam: float) -> torch.Tensor:
    """The reference's reverse scan, applied independently to every column of ``[T,N]``
    float32 tensors (for N=1 this *is* the reference's flat buffer).

    Quirks kept on purpose: the last step bootstraps from its own value
    (``next_value = values[T-1]*(1-dones[T-1])``, :22-24); earlier steps mask with
    ``dones[t+1]`` rather than ``dones[t]`` (:26-27)."""
    rewards = torch.as_tensor(rewards, dtype=torch.float32)
    values = torch.as_tensor(values, dtype=torch.float32)
    dones = torch.as_tensor(dones, dtype=torch.float32)
    squeeze = rewards.dim() == 1
    if squeeze:
        rewards, values, dones = rewards[:, None], values[:, None], dones[:, None]
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    for t in reversed(range(T)):
        if t == T - 1:
            nnt = 1.0 - dones[t]
            nv = values[t] * nnt
        else:
            nnt = 1.0 - dones[t + 1]
            nv = values[t + 1] * nnt
        delta = rewards[t] + gamma * nv - values[t]
        adv[t] = delta + gamma * lam * nnt * last
        last = adv[t]
    return adv[:, 0] if squeeze else adv


def normalise_advantages(adv: torch.Tensor, values: torch.Tensor):
    """mean-centre, divide by the unbiased std (+1e-6; std replaced by 1 if <1e-6 or NaN),
    and -- sic -- returns = *normalised* advantage + value (:34-39).  Statistics are taken
    over every element of ``adv`` (the whole ``[T,N]`` batch)."""
    a = adv - adv.mean()
    s = a.std()
    if s < 1e-6 or torch.isnan(s):
        s = 1.0
    a = a / (s + 1e-6)
    return a, a + values


def gae_standard(rewards, values, dones, last_value, gamma: float, lam: float) -> torch.Tensor:
    """P5' (not a parity target): bootstrap with V(s_T), PPOV1.1/train_ppo1.0.py:66-89."""
    rewards = torch.as_tensor(rewards, dtype=torch.float32)
    values = torch.as_tensor(values, dtype=torch.float32)
    dones = torch.as_tensor(dones, dtype=torch.float32)
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    last = torch.zeros_like(rewards[0])
    for t in reversed(range(T)):
        nv = torch.as_tensor(last_value, dtype=torch.float32) if t == T - 1 else values[t + 1]
        nnt = 1.0 - dones[t]
        delta = rewards[t] + gamma * nv * nnt - values[t]
        adv[t] = delta + gamma * lam * nnt * last
        last = adv[t]
    return adv


def gae_bootstrap_v10(rewards, values, dones, next_value, gamma: float, lam: float):
    """P5' literal transcription of PPOV1.1/train_ppo1.0.py:72-89 (same loop in PPOV1.0/ppo0.0.py:330-353 and
    train_ppo_gail.py) for one flat buffer: returns (normalised advantages, returns = RAW advantage + value)."""
    rewards = torch.as_tensor(rewards, dtype=torch.float32)
    values = torch.as_tensor(values, dtype=torch.float32)
    dones = torch.as_tensor(dones, dtype=torch.float32)
    next_value = torch.as_tensor(next_value, dtype=torch.float32)
    advantages = torch.zeros_like(rewards)
    returns = torch.zeros_like(rewards)
    gae = 0
    for t in reversed(range(len(rewards))):
        if t == len(rewards) - 1:
            next_non_terminal = 1.0 - dones[t]
            next_value_t = next_value
        else:
            next_non_terminal = 1.0 - dones[t + 1]
            next_value_t = values[t + 1]
        delta = rewards[t] + gamma * next_value_t * next_non_terminal - values[t]
        gae = delta + gamma * lam * next_non_terminal * gae
        advantages[t] = gae
        returns[t] = advantages[t] + values[t]
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return advantages, returns


def gae_v12(rewards, values, dones, gamma: float, lam: float):
    """P5' literal transcription of PPOV1.2/ppo注释版.py:366-380 for one flat buffer: returns (normalised
    advantages, returns = normalised advantage + value)."""
    rewards = torch.as_tensor(rewards, dtype=torch.float32)
    values = torch.as_tensor(values, dtype=torch.float32)
    dones = torch.as_tensor(dones, dtype=torch.float32)
    advantages = torch.zeros_like(rewards)
    last_advantage = 0
    for t in reversed(range(len(rewards))):
        if t < len(rewards) - 1:
            next_value = values[t + 1] * (1 - dones[t])
        else:
            next_value = 0
        delta = rewards[t] + gamma * next_value - values[t]
        advantages[t] = delta + gamma * lam * last_advantage * (1 - dones[t])
        last_advantage = advantages[t]
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return advantages, advantages + values


# --------------------------------------------------------------------------------------
# P6/P7: clipped-surrogate loss and optimiser step, train_ppo2.0.py:42-87
# --------------------------------------------------------------------------------------
def ppo_loss(model, states, actions, old_log_probs, advantages, returns, old_values, cfg: PlumeConfig):
    """Loss of one minibatch; returns (total, policy_loss, value_loss, entropy)."""
    probs, values = model(states)                                                    # :54
    if torch.isnan(probs).any():                                                     # :57-61
        raise RuntimeError("NaN in probs")
    logp = categorical_log_prob(probs, actions)                                      # :63-64
    ratio = (logp - old_log_probs).exp()                                             # :67
    surr1 = ratio * advantages
    surr2 = torch.clamp(ratio, 1 - cfg.clip_epsilon, 1 + cfg.clip_epsilon) * advantages
    policy_loss = -torch.min(surr1, surr2).mean()                                    # :70
    v = values.squeeze()
    v_clipped = old_values + (v - old_values).clamp(-cfg.clip_epsilon, cfg.clip_epsilon)   # :73
    value_loss = 0.5 * torch.max((v - returns).pow(2), (v_clipped - returns).pow(2)).mean()  # :74-77
    entropy = -torch.sum(probs * torch.log(probs + 1e-8), dim=1).mean()              # :80
    total = policy_loss + value_loss - cfg.entropy_beta * entropy                    # :82
    return total, policy_loss, value_loss, entropy


def ppo_update(model, optimizer, states, actions, rewards, values, log_probs, dones, cfg: PlumeConfig,
               perms=None, adv_ret=None, record=None):
    """Full ``_update_model`` on flat ``[M]`` tensors (``[M,6]`` states).

    ``perms``: list of ``cfg.epochs`` index permutations (the reference draws
    ``torch.randperm`` per epoch, :43); ``adv_ret``: optional precomputed
    (advantages, returns) for batched ``[T,N]`` layouts flattened by the caller."""
    if adv_ret is None:
        adv = gae_quirk(rewards, values, dones, cfg.gamma, cfg.lam)
        adv, ret = normalise_advantages(adv, values)
    else:
        adv, ret = adv_ret
    M = states.shape[0]
    for e in range(cfg.epochs):
        perm = torch.randperm(M) if perms is None else torch.as_tensor(perms[e], dtype=torch.long)
        for idx in perm.split(cfg.batch_size):
            if len(idx) == 0:
                continue
            total, pl, vl, ent = ppo_loss(model, states[idx], actions[idx], log_probs[idx], adv[idx],
                                          ret[idx], values[idx], cfg)
            optimizer.zero_grad()
            total.backward()
            gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)          # :86
            if record is not None:
                record.append({"loss": float(total.detach()), "policy_loss": float(pl.detach()),
                               "value_loss": float(vl.detach()), "entropy": float(ent.detach()),
                               "grad_norm": float(gnorm)})
            optimizer.step()                                                         # :87
    return adv, ret


# --------------------------------------------------------------------------------------
# P8: curriculum, model.py:178-221
# --------------------------------------------------------------------------------------
class OracleCurriculum:
    """``PPOTrainer.update`` restated; ``env`` is any object with ``current_radius`` and
    ``explore_bonus`` attributes."""

    def __init__(self, env, cfg: PlumeConfig):
        self.env = env
        self.cfg = cfg
        self.success_history: list = []
        self.current_radius = cfg.initial_radius
        self.explore_bonus = cfg.explore_bonus

    def update(self, success) -> None:
        c = self.cfg
        self.env.current_radius = self.current_radius              # :189 (env lags by one episode)
        self.env.explore_bonus = self.explore_bonus                # :190
        self.success_history.append(success)                       # :192
        if len(self.success_history) > c.window_size:
            self.success_history.pop(0)
        full = len(self.success_history) >= c.window_size
        if full:                                                   # :197-199
            rate = np.mean(self.success_history[-c.window_size:])
            self.explore_bonus *= (c.decay_factor ** (1 + rate))
        self.explore_bonus = max(self.explore_bonus, 0.1)          # :201
        if full:                                                   # :203-221
            rate = np.mean(self.success_history[-c.window_size:])
            if rate > c.success_threshold:
                self.current_radius = max(
                    c.min_radius, self.current_radius * (c.radius_decay ** (2 + 3 * (rate - c.success_threshold))))
            elif rate < 0.25:
                self.current_radius = min(c.initial_radius, self.current_radius * 1.1)
            if abs(self.current_radius - self.env.current_radius) > 5:
                self.current_radius = self.env.current_radius + 5 * np.sign(
                    self.current_radius - self.env.current_radius)
            self.success_history = []


# --------------------------------------------------------------------------------------
# P4L: LSTM stop heads
# --------------------------------------------------------------------------------------
class OraclePeakAndStop(nn.Module):
    """V2.1 ``PeakAndStopPredictor`` (evaluate_with_lstm.py:11-27): LSTM(1->H) from zero
    state over the window, last hidden -> peak (linear) and stop probability (sigmoid)."""

    def __init__(self, input_dim: int = 1, hidden_dim: int = 32, num_layers: int = 1):
        super().__init__()
        self.lstm = nn.LSTM(input_dim, hidden_dim, num_layers=num_layers, batch_first=True)
        self.fc_peak = nn.Linear(hidden_dim, 1)
        self.fc_stop = nn.Sequential(nn.Linear(hidden_dim, 1), nn.Sigmoid())

    def forward(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(-1)
        _, (h_n, _) = self.lstm(x)
        h = h_n[-1]
        return self.fc_peak(h).squeeze(-1), self.fc_stop(h).squeeze(-1)


def lstm_window_stop(model: OraclePeakAndStop, conc_history: np.ndarray, window: int = 20, threshold: float = 0.8):
    """evaluate_with_lstm.py:73-80 for a batch: ``conc_history`` [B, >=window] raw
    concentrations (0..100); returns (peak, stop_prob, stop_flag)."""
    seq = np.asarray(conc_history, dtype=np.float64)[:, -window:].reshape(-1, window, 1) / 100.0
    with torch.no_grad():
        peak, prob = model(torch.FloatTensor(seq))
    return peak.numpy(), prob.numpy(), prob.numpy() > threshold


class OracleThresholdPredictor(nn.Module):
    """V2.0 ``ConcentrationThresholdPredictor`` (PPOV2.0/model.py:203-240): 3-layer
    LSTM(1->H) (dropout inactive in eval), FC H->64 -> LN -> ReLU -> (Dropout) -> 1."""

    def __init__(self, input_size: int = 1, hidden_size: int = 128):
        super().__init__()
        self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=3,
                            batch_first=True, dropout=0.3)
        self.fc = nn.Sequential(nn.Linear(hidden_size, 64), nn.LayerNorm(64), nn.ReLU(),
                                nn.Dropout(0.1), nn.Linear(64, 1))
        for name, p in self.named_parameters():                     # PPOV2.0/model.py:222-227
            if "weight" in name and p.dim() > 1:
                nn.init.xavier_uniform_(p)
            elif "bias" in name:
                nn.init.zeros_(p)

    def forward(self, x, lengths=None):
        # PPOV2.0/model.py:229-240: packed sequences (all windows have the full length on the hot
        # path, evaluate_with_lstm.py:25; the packed kernel path is kept so the bits agree)
        if lengths is None:
            lengths = [x.shape[1]] * x.shape[0]
        packed = nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True, enforce_sorted=False)
        out, _ = self.lstm(packed)
        unpacked, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
        last = unpacked[torch.arange(unpacked.size(0)), torch.tensor(lengths) - 1]
        return self.fc(last).squeeze()


class OracleThresholdController:
    """V2.0 ``ThresholdController`` (PPOV2.0/evaluate_with_lstm.py:10-37) with a MinMax
    scaler given by (data_min, data_max) (``MinMaxScaler.transform``: (x-min)/(max-min))."""

    def __init__(self, model, data_min: float, data_max: float, window_size: int = 10):
        self.model = model
        self.data_min, self.data_max = data_min, data_max
        self.window_size = window_size
        self.current_threshold = None
        self.conc_buffer: list = []
        self.min_activate_steps = 2 * window_size

    def update_threshold(self, trajectory):
        if len(trajectory) >= max(self.window_size, self.min_activate_steps):
            w = np.array(trajectory[-self.window_size:]).reshape(-1, 1)
            rng = self.data_max - self.data_min
            scale = 1.0 / (rng if rng != 0 else 1.0)
            scaled = w * scale + (0.0 - self.data_min * scale)       # sklearn MinMaxScaler.transform
            with torch.no_grad():
                pred = self.model(torch.FloatTensor(scaled).unsqueeze(0), lengths=[self.window_size])
            self.current_threshold = pred.item() * 0.95

    def should_stop(self, current_conc, step_count) -> bool:
        self.conc_buffer.append(current_conc)
        if len(self.conc_buffer) > self.window_size:
            self.conc_buffer.pop(0)
        return bool(step_count >= self.min_activate_steps and self.current_threshold is not None
                    and (current_conc >= self.current_threshold
                         or np.mean(self.conc_buffer) >= self.current_threshold))


def fixed_threshold_stop(positions, conc_reward_last: float, cfg: PlumeConfig,
                         window: int = 10, stability: float = 2.0) -> bool:
    """V1.1 stop test (PPOV1.1/evaluate_model.py:25-37): mean over axes of the std of the
    last ``window`` positions < 2 px and ``conc > 0.8*CONC_PEAK`` where (sic) ``conc`` is
    ``info['concentration_reward']*CONC_PEAK*CONC_PEAK`` (:59-62 then :35)."""
    if len(positions) < window:
        return False
    last = np.asarray(positions[-window:])
    pos_std = np.std(last, axis=0).mean()
    current = (conc_reward_last * cfg.conc_peak) * cfg.conc_peak
    return bool((pos_std < stability) and (current > 0.8 * cfg.conc_peak))


# --------------------------------------------------------------------------------------
# P4t: trend features (dead code in the reference = the spec), model.py:113-127
# --------------------------------------------------------------------------------------
def trend_label(conc, pos_last, src, conc_peak: float = 100.0):
    """Returns (label, trend_score, dist_score, conc_score) for one window."""
    conc = np.asarray(conc, dtype=np.float64)
    dist = np.linalg.norm(np.asarray(pos_last, dtype=np.float64) - np.asarray(src, dtype=np.float64))
    dist_score = np.exp(-dist / 50.0)
    grad = np.gradient(conc)
    trend_score = np.tanh(np.mean(grad[-3:]) / 5.0)
    conc_score = np.clip(conc[-1] / conc_peak, 0, 1)
    label = 0.4 * dist_score + 0.3 * (trend_score + 1) / 2 + 0.3 * conc_score
    return float(np.clip(label, 0.01, 0.99)), float(trend_score), float(dist_score), float(conc_score)


# --------------------------------------------------------------------------------------
# P9 + driver loop: single-env training loop used as the timed CPU baseline
# --------------------------------------------------------------------------------------
def cpu_train_loop(env, model, optimizer, cfg: PlumeConfig, n_steps: int, curriculum=None, seed: int = 0):
    """The reference's rollout/update loop (train_ppo2.0.py:137-192,251) for ``n_steps``
    env steps: batch-1 policy forward, Categorical sample, env.step, store, update every
    ``cfg.batch_size`` transitions.  Returns (#env steps, #updates, #episodes)."""
    torch.manual_seed(seed)
    buf = {k: [] for k in ("s", "a", "r", "v", "lp", "d")}
    steps = updates = episodes = 0
    state = env.reset()
    while steps < n_steps:
        with torch.no_grad():
            probs, value = model(torch.FloatTensor(state).unsqueeze(0))
        dist = torch.distributions.Categorical(probs)
        action = dist.sample().item()
        nxt, reward, done, _ = env.step(action)
        buf["s"].append(np.array(state, dtype=np.float32))
        buf["a"].append(int(action))
        buf["r"].append(float(reward))
        buf["v"].append(float(value.item()))
        buf["lp"].append(float(dist.log_prob(torch.tensor(action)).item()))
        buf["d"].append(float(done))
        steps += 1
        if len(buf["s"]) >= cfg.batch_size:
            ppo_update(model, optimizer, torch.FloatTensor(np.stack(buf["s"])), torch.LongTensor(buf["a"]),
                       torch.FloatTensor(buf["r"]), torch.FloatTensor(buf["v"]), torch.FloatTensor(buf["lp"]),
                       torch.FloatTensor(buf["d"]), cfg)
            for v in buf.values():
                v.clear()
            updates += 1
        state = nxt
        if done:
            episodes += 1
            if curriculum is not None:
                curriculum.update(bool(env.reached))
            state = env.reset()
    return steps, updates, episodes
