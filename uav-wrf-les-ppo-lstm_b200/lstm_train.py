"""Supervised training of the V2.1 stop head on the GPU (PPOV2.1/train_lstm.py).

``build_dataset`` is ``TrajectoryDataset._preprocess`` (train_lstm.py:28-65) over the ``training_data.nc``
variables a ``TrajectoryLogger`` holds on the device; ``LstmTrainer`` is the loop of ``train()``
(train_lstm.py:102-125): minibatches of 64 in shuffled order, loss = MSE(peak) + BCE(stop),
``clip_grad_norm_(1.0)``, ``AdamW(lr=1e-3, weight_decay=1e-4)``, ``ReduceLROnPlateau('min', patience=5)`` on the
epoch mean, best-loss checkpoint.  Forward, back-propagation through time, gradient reduction, clipping and the
AdamW update of one minibatch are ONE kernel launch (csrc/lstm_train_kernels.cu); an epoch is one C call.  The
learning-rate schedule and the checkpoint decision are host scalars, as in the reference.  No CPU fallback."""
from __future__ import annotations

import copy
import math

import numpy as np
import torch

from . import _lib
from .model import PeakAndStopPredictor


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def eligible_episodes(steps, window: int = 20) -> np.ndarray:
    """Episodes ``load_trajectory_segments`` keeps (model.py:71-73: at least ``window`` logged steps), in file
    order -- the keys of ``episode_dict`` (train_lstm.py:33-39) from which ``random.sample`` draws."""
    steps = steps.cpu().numpy() if isinstance(steps, torch.Tensor) else np.asarray(steps)
    return np.nonzero(steps >= window)[0].astype(np.int32)


@torch.no_grad()
def build_dataset(nc: dict, episode_ids, window: int = 20, stop_radius: float = 10.0, device="cuda"):
    """``nc``: the training_data.nc variables (``TrajectoryLogger.nc_variables()`` or device tensors with the same
    names: ``x``, ``y``, ``concentration`` [E,S] float32, ``source_x``, ``source_y`` [E]).  ``episode_ids``: the
    selected episodes in the order ``random.sample`` returned them (train_lstm.py:40).  Returns
    ``(features [2M, window], labels [2M, 2])`` float32 on the device."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("build_dataset runs on CUDA only (no CPU fallback)")
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous()
    conc, x, y = t(nc["concentration"]), t(nc["x"]), t(nc["y"])
    sx, sy = t(nc["source_x"]), t(nc["source_y"])
    ids = torch.as_tensor(np.asarray(episode_ids), dtype=torch.int32, device=dev).contiguous()
    M, S = int(ids.numel()), int(conc.shape[1])
    feats = torch.empty(2 * M, window, dtype=torch.float32, device=dev)
    labels = torch.empty(2 * M, 2, dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.plume_lstm_dataset(conc.data_ptr(), x.data_ptr(), y.data_ptr(), sx.data_ptr(), sy.data_ptr(), S,
                                          ids.data_ptr(), M, window, float(stop_radius), feats.data_ptr(),
                                          labels.data_ptr(), _stream(dev)), "plume_lstm_dataset")
    return feats, labels


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, 'min', patience=5) with torch's defaults
    (factor 0.1, threshold 1e-4 'rel', cooldown 0, min_lr 0, eps 1e-8) on a scalar learning rate."""

    def __init__(self, lr: float, factor: float = 0.1, patience: int = 5, threshold: float = 1e-4,
                 cooldown: int = 0, min_lr: float = 0.0, eps: float = 1e-8):
        self.lr, self.factor, self.patience, self.threshold = float(lr), factor, patience, threshold
        self.cooldown, self.min_lr, self.eps = cooldown, min_lr, eps
        self.best, self.num_bad_epochs, self.cooldown_counter = math.inf, 0, 0

    def step(self, metric: float) -> float:
        current = float(metric)
        if current < self.best * (1.0 - self.threshold):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.lr * self.factor, self.min_lr)
            if self.lr - new_lr > self.eps:
                self.lr = new_lr
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0
        return self.lr


class LstmTrainer:
    """``train()`` of train_lstm.py:102-125 for a ``PeakAndStopPredictor(hidden_dim=32)`` on the GPU."""

    def __init__(self, model: PeakAndStopPredictor, features: torch.Tensor, labels: torch.Tensor,
                 batch_size: int = 64, lr: float = 1e-3, weight_decay: float = 1e-4, betas=(0.9, 0.999),
                 eps: float = 1e-8, max_norm: float = 1.0, patience: int = 5):
        dev = features.device
        if dev.type != "cuda":
            raise RuntimeError("LstmTrainer runs on CUDA only (no CPU fallback)")
        self.model, self.device = model, dev
        self.flat = model.flatten_()
        if self.flat.device != dev:
            raise ValueError("model and dataset must live on the same device")
        self.features = features.to(torch.float32).contiguous()
        self.labels = labels.to(torch.float32).contiguous()
        self.n, self.window = int(features.shape[0]), int(features.shape[1])
        self.batch_size = int(batch_size)
        self.weight_decay, self.betas, self.eps, self.max_norm = float(weight_decay), betas, float(eps), float(max_norm)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.opt_step = 0
        self.scheduler = ReduceLROnPlateau(lr, patience=patience)
        self.lib = _lib.load()
        nbytes = int(self.lib.plume_lstm_train_workspace_bytes(self.batch_size))
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.n_batches = (self.n + self.batch_size - 1) // self.batch_size
        self.batch_losses = torch.zeros(self.n_batches, dtype=torch.float32, device=dev)
        self.grad_norms = torch.zeros(self.n_batches, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(_lib.LSTM_TRAIN_PARAMS, dtype=torch.float32, device=dev)
        self.best_loss, self.best_state = math.inf, None
        self.history = []        # (epoch mean loss, lr used) per epoch
        self.launches = 0

    @property
    def lr(self) -> float:
        return self.scheduler.lr

    def epoch_order(self, generator=None) -> torch.Tensor:
        """The sample order of one DataLoader(shuffle=True) epoch: torch's RandomSampler draws a seed from the
        default generator and permutes with a fresh generator (torch/utils/data/sampler.py)."""
        if generator is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            generator = torch.Generator()
            generator.manual_seed(seed)
        return torch.randperm(self.n, generator=generator)

    @torch.no_grad()
    def train_epoch(self, order=None) -> float:
        """One pass over the dataset in ``order`` ([n] sample ids; default: a fresh shuffle).  Returns the mean of
        the minibatch losses (train_lstm.py:122), steps the plateau scheduler and keeps the best state."""
        if self.n == 0:
            raise ValueError("empty dataset (train_lstm.py:133-135 exits here)")
        if order is None:
            order = self.epoch_order()
        order = torch.as_tensor(order).to(device=self.device, dtype=torch.int32).contiguous()
        if order.numel() != self.n:
            raise ValueError("order must list every sample once")
        lr_used = self.lr
        with torch.cuda.device(self.device):
            _lib.check(self.lib.plume_lstm_train_epoch(
                self.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.model.hidden_dim,
                self.features.data_ptr(), self.labels.data_ptr(), order.data_ptr(), self.n, self.window,
                self.batch_size, self.max_norm, lr_used, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                self.opt_step + 1, self.workspace.data_ptr(), self.workspace.numel(), self.batch_losses.data_ptr(),
                self.grad_norms.data_ptr(), self.grad.data_ptr(), _stream(self.device)), "plume_lstm_train_epoch")
        self.opt_step += self.n_batches
        self.launches += self.n_batches
        # epoch_loss += loss.item() per minibatch, then / len(dataloader): float32 losses summed in float64
        avg = float(self.batch_losses.double().sum().item()) / self.n_batches
        if not math.isfinite(avg):
            raise RuntimeError("NaN/Inf loss in LSTM training")
        self.scheduler.step(avg)
        if avg < self.best_loss:                                   # train_lstm.py:124-126
            self.best_loss = avg
            self.best_state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        self.history.append((avg, lr_used))
        return avg

    def train(self, epochs: int = 100, orders=None, verbose: bool = False):
        for e in range(epochs):
            avg = self.train_epoch(None if orders is None else orders[e])
            if verbose:
                print(f"Epoch {e + 1:03d} | Loss: {avg:.4f} | LR: {self.lr:.2e}")
        return self.history

    def save_best(self, path: str) -> None:
        """``torch.save(model.state_dict(), "model/best_peak_and_stop.pth")`` of the best epoch."""
        torch.save(copy.deepcopy(self.best_state), path)
