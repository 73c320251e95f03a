"""TEST INFRASTRUCTURE -- CPU restatement of the supervised stop-head training.

Follows /root/reference/PPOV2.1/train_lstm.py and model.py:67-91 line by line (cited below).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this; the product path never does.  The
reference's own arithmetic here is torch (nn.LSTM, AdamW, clip_grad_norm_, ReduceLROnPlateau -- torch 2.11.0, the
version installed in this image), so the restatement calls the same torch ops on the CPU.  Pinned against the
unmodified reference by tests/test_oracle_vs_reference.py (bit-equal checkpoints) and tests/golden/lstm_train_s13.npz.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------
# model.py:67-91 + train_lstm.py:28-65
# --------------------------------------------------------------------------------------
def eligible_episodes(nc: dict, window: int = 20) -> np.ndarray:
    """Episodes with at least ``window`` non-NaN x entries (model.py:71-73), file order = insertion order of
    ``episode_dict`` (train_lstm.py:33-39; keys are the source coordinates, assumed distinct)."""
    valid = (~np.isnan(np.asarray(nc["x"]))).sum(axis=1)
    return np.nonzero(valid >= window)[0].astype(np.int32)


def select_episodes(n_eligible: int, seed: int) -> list:
    """train_lstm.py:40 ``random.sample(list(episode_dict.values()), min(1000, len))``: the positions drawn only
    depend on the population size, so the same seed selects the same positions."""
    rng = random.Random(seed)
    return rng.sample(list(range(n_eligible)), min(1000, n_eligible))


def build_dataset(nc: dict, episode_ids, window: int = 20, stop_radius: float = 10.0):
    """Two samples per selected episode from its FIRST window (``ep_segs[0]``, train_lstm.py:42): the 'negative'
    one (:45-52) and the 'positive' one (:54-63), which reads ``conc[-window:]`` of a window-long segment, i.e. the
    same values, and ``positions[-1]`` = the position at step window-1."""
    feats, labels = [], []
    for ep in episode_ids:
        valid = np.where(~np.isnan(nc["x"][ep]))[0]                  # model.py:71
        x = nc["x"][ep, valid][:window]
        y = nc["y"][ep, valid][:window]
        conc = np.array(nc["concentration"][ep, valid][:window])     # train_lstm.py:43
        src = np.array([nc["source_x"][ep], nc["source_y"][ep]])     # model.py:77
        neg = conc[:window].reshape(-1, 1) / 100.0                   # :46
        peak = conc[window - 1]                                      # :47
        labels.append([peak / 100.0, 0.0])
        feats.append(neg)
        pos = conc[-window:].reshape(-1, 1) / 100.0                  # :55
        last_pos = np.column_stack((x, y))[-1]                       # :56 (model.py:82)
        stop = 1.0 if np.linalg.norm(last_pos - src) <= stop_radius else 0.0   # :58
        labels.append([conc[-1] / 100.0, stop])                      # :59-63
        feats.append(pos)
    if not feats:
        return np.zeros((0, window), np.float32), np.zeros((0, 2), np.float32)
    f = torch.FloatTensor(np.stack(feats)).squeeze(-1).numpy()       # __getitem__, :73-77
    l = torch.FloatTensor(np.array(labels)).numpy()
    return f, l


# --------------------------------------------------------------------------------------
# train_lstm.py:84-100
# --------------------------------------------------------------------------------------
class PeakAndStopPredictor(nn.Module):
    def __init__(self, input_dim=1, hidden_dim=32, num_layers=1):
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


def loss_and_grad(model: nn.Module, feats: torch.Tensor, labels: torch.Tensor):
    """train_lstm.py:112-117 for one minibatch: (loss, flat gradient in named_parameters order)."""
    model.zero_grad()
    peak, stop = model(feats.unsqueeze(-1))
    loss = nn.MSELoss()(peak, labels[:, 0]) + nn.BCELoss()(stop, labels[:, 1])
    loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    return float(loss.item()), g


def epoch_order(n: int) -> torch.Tensor:
    """One DataLoader(shuffle=True) epoch (torch/utils/data/dataloader.py + sampler.py, torch 2.11): the iterator
    draws its base seed from the default generator, then RandomSampler draws a seed for a fresh generator."""
    torch.empty((), dtype=torch.int64).random_()                                      # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())                   # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def train(model: nn.Module, feats, labels, epochs: int = 100, batch_size: int = 64, orders=None, lr: float = 1e-3,
          weight_decay: float = 1e-4, max_norm: float = 1.0, patience: int = 5) -> dict:
    """train_lstm.py:104-127.  ``orders`` [epochs][n] fixes the shuffles; None draws them like the DataLoader."""
    feats = torch.as_tensor(feats, dtype=torch.float32)
    labels = torch.as_tensor(labels, dtype=torch.float32)
    n = feats.shape[0]
    opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)      # :105
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", patience=patience)  # :106
    best, hist, ckpts, used_orders, batch_losses, grad_norms = float("inf"), [], [], [], [], []
    for e in range(epochs):
        order = torch.as_tensor(orders[e]) if orders is not None else epoch_order(n)
        used_orders.append(order.numpy().astype(np.int64))
        model.train()
        epoch_loss, nb = 0.0, 0
        lr_used = opt.param_groups[0]["lr"]
        for b0 in range(0, n, batch_size):
            idx = order[b0:b0 + batch_size].long()
            f, l = feats[idx].unsqueeze(-1), labels[idx]
            opt.zero_grad()                                                           # :112
            peak, stop = model(f)
            loss = nn.MSELoss()(peak, l[:, 0]) + nn.BCELoss()(stop, l[:, 1])           # :114-116
            loss.backward()
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)         # :118
            opt.step()
            epoch_loss += loss.item()                                                  # :120
            batch_losses.append(loss.item())
            grad_norms.append(float(gn))
            nb += 1
        avg = epoch_loss / nb                                                          # :121
        sched.step(avg)
        if avg < best:                                                                 # :123-125
            ckpts.append({k: v.detach().clone() for k, v in model.state_dict().items()})
            best = avg
        hist.append((avg, lr_used, opt.param_groups[0]["lr"]))
    return {"history": hist, "checkpoints": ckpts, "orders": used_orders, "batch_losses": np.array(batch_losses),
            "grad_norms": np.array(grad_norms),
            "final": {k: v.detach().clone() for k, v in model.state_dict().items()}}


def synthetic_nc(n_episodes: int, seed: int, max_steps: int = 64, window: int = 20) -> dict:
    """A small training_data.nc-shaped record (nc_info.txt variable names, float32, NaN padded): random-walk
    trajectories with a concentration that rises towards the source; some episodes shorter than ``window``
    (skipped by model.py:72-73), some ending within the stop radius of the source."""
    rng = np.random.default_rng(seed)
    x = np.full((n_episodes, max_steps), np.nan, np.float32)
    y = np.full_like(x, np.nan)
    c = np.full_like(x, np.nan)
    sx = (rng.random(n_episodes) * 400 + 50).astype(np.float32)
    sy = (rng.random(n_episodes) * 400 + 50).astype(np.float32)
    for e in range(n_episodes):
        L = int(rng.integers(window - 6, max_steps + 1))
        near = rng.random() < 0.45          # starts close to the source: label stop = 1 after `window` steps
        p0 = np.array([sx[e], sy[e]]) + (rng.normal(size=2) * 4 if near else rng.normal(size=2) * 120)
        steps = rng.normal(size=(L, 2)) * (0.8 if near else 12.0)
        pos = np.clip(p0 + np.cumsum(steps, axis=0), 0, 499)
        d = np.linalg.norm(pos - np.array([sx[e], sy[e]]), axis=1)
        conc = np.clip(100 * np.exp(-d ** 2 / 450.0) + 3 * np.abs(rng.normal(size=L)), 0, 100)
        x[e, :L], y[e, :L], c[e, :L] = pos[:, 0], pos[:, 1], conc
    return {"episode": np.arange(n_episodes, dtype=np.int32), "step": np.arange(max_steps, dtype=np.int32), "x": x,
            "y": y, "concentration": c, "source_x": sx, "source_y": sy,
            "gaussian_sigma": np.full(n_episodes, 15.0, np.float32)}
