"""Drop-ins for the reference models (PPOV2.1/model.py:16-46, evaluate_with_lstm.py:11-27,
PPOV2.0/model.py:203-240) whose ``forward`` runs the hand-written kernels.

The modules keep the reference's parameter names and shapes, so ``state_dict()`` /
``load_state_dict()`` interchange ``.pth`` files with the reference in both directions.
``PPOActorCritic`` stores all parameters in one flat fp32 buffer (the layout of
include/plume_b200.h); every ``nn.Parameter`` is a view into it, and so are the gradients,
which lets the update kernels, the single NCCL all-reduce and the fused Adam work on one
contiguous array.  ``forward`` is inference only; training goes through
``learner.update_model`` (the gradient is computed by csrc/ppo_kernels.cu, not autograd).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class PPOActorCritic(nn.Module):
    """6 -> 256 -> LayerNorm -> ReLU -> 128 -> LayerNorm -> ReLU -> {actor 5, critic 1}
    (model.py:17-36): orthogonal init (gain sqrt(2) hidden, 0.01 actor, 1.0 critic), zero bias."""

    def __init__(self, input_size: int = 6, output_size: int = 5, device="cuda"):
        super().__init__()
        if input_size != 6 or output_size != 5:
            raise ValueError("the plume kernels are specialised for the reference's 6 -> 5 policy")
        self.feature = nn.Sequential(nn.Linear(input_size, 256), nn.LayerNorm(256), nn.ReLU(),
                                     nn.Linear(256, 128), nn.LayerNorm(128), nn.ReLU())
        for layer in self.feature:
            if isinstance(layer, nn.Linear):
                nn.init.orthogonal_(layer.weight, gain=np.sqrt(2))
                nn.init.constant_(layer.bias, 0.0)
        self.actor = nn.Linear(128, output_size)
        self.critic = nn.Linear(128, 1)
        nn.init.orthogonal_(self.actor.weight, gain=0.01)
        nn.init.constant_(self.actor.bias, 0.0)
        nn.init.orthogonal_(self.critic.weight, gain=1.0)
        nn.init.constant_(self.critic.bias, 0.0)
        self.flat = None
        self.flat_grad = None
        self._flatten(torch.device(device))

    # -- flat storage ------------------------------------------------------------------------
    def _flatten(self, device: torch.device) -> None:
        flat = torch.zeros(_lib.MLP_PARAMS, dtype=torch.float32, device=device)
        grad = torch.zeros_like(flat)
        named = dict(self.named_parameters())
        for name, (off, shape) in _lib.MLP_OFFSETS.items():
            p = named[name]
            n = int(np.prod(shape))
            flat[off:off + n].copy_(p.detach().reshape(-1).to(device))
            p.data = flat[off:off + n].view(shape)
            p.grad = grad[off:off + n].view(shape)
        self.flat, self.flat_grad = flat, grad
        if device.type == "cuda":
            self.nan_flag = torch.zeros(1, dtype=torch.int32, device=device)

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._flatten(next(self.parameters()).device)
        return self

    @property
    def device(self):
        return self.flat.device

    # -- P4 forward ------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """``(probs [B,5], value [B,1])``; raises ``RuntimeError("NaN in model output")`` like
        model.py:41-43 (one flag read-back per call)."""
        probs, value = self.forward_async(x)
        if int(self.nan_flag.item()) != 0:
            self.nan_flag.zero_()
            raise RuntimeError("NaN in model output")
        return probs, value

    @torch.no_grad()
    def forward_async(self, x: torch.Tensor):
        """Same as ``forward`` without the NaN read-back (check ``nan_flag`` later)."""
        if self.flat.device.type != "cuda":
            raise RuntimeError("PPOActorCritic.forward runs on CUDA only (no CPU fallback)")
        lib = _lib.load()
        x = x.to(device=self.device, dtype=torch.float32).reshape(-1, 6).contiguous()
        B = x.shape[0]
        probs = torch.empty(B, 5, dtype=torch.float32, device=self.device)
        value = torch.empty(B, 1, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.plume_policy_forward(self.flat.data_ptr(), x.data_ptr(), B, probs.data_ptr(),
                                                value.data_ptr(), self.nan_flag.data_ptr(), _stream(self.device)),
                       "plume_policy_forward")
        return probs, value

    @torch.no_grad()
    def act(self, obs: torch.Tensor, env=None, uniforms=None, forced_actions=None, greedy: bool = False):
        """P4s: forward + ``Categorical`` sample / log_prob (train_ppo2.0.py:161-162,185).
        Returns ``(action int32 [B], log_prob [B], value [B], probs [B,5])``.  The draw uses
        ``uniforms`` if given, else the Philox action stream of ``env``."""
        lib = _lib.load()
        obs = obs.to(device=self.device, dtype=torch.float32).reshape(-1, 6).contiguous()
        B = obs.shape[0]
        actions = torch.empty(B, dtype=torch.int32, device=self.device)
        logp = torch.empty(B, dtype=torch.float32, device=self.device)
        value = torch.empty(B, dtype=torch.float32, device=self.device)
        probs = torch.empty(B, 5, dtype=torch.float32, device=self.device)
        u = None if uniforms is None else uniforms.to(device=self.device, dtype=torch.float32).contiguous()
        f = None if forced_actions is None else forced_actions.to(device=self.device, dtype=torch.int32).contiguous()
        cfg = C.byref(env.c_config) if env is not None else None
        st = C.byref(env.c_state) if env is not None else None
        with torch.cuda.device(self.device):
            _lib.check(lib.plume_policy_act(cfg, st, self.flat.data_ptr(), obs.data_ptr(), B, _lib.ptr(u), _lib.ptr(f),
                                            _lib.FLAG_GREEDY if greedy else 0, actions.data_ptr(), logp.data_ptr(),
                                            value.data_ptr(), probs.data_ptr(), self.nan_flag.data_ptr(),
                                            _stream(self.device)), "plume_policy_act")
        return actions, logp, value, probs


class PeakAndStopPredictor(nn.Module):
    """V2.1 stop head (evaluate_with_lstm.py:11-27): LSTM(1 -> hidden) from zero state over
    the window, ``fc_peak`` and ``fc_stop`` (sigmoid) on the last hidden state."""

    def __init__(self, input_dim: int = 1, hidden_dim: int = 32, num_layers: int = 1, device="cuda"):
        super().__init__()
        if input_dim != 1 or num_layers != 1:
            raise ValueError("the stop-head kernel implements the reference's LSTM(1 -> hidden, 1 layer)")
        self.hidden_dim = hidden_dim
        self.lstm = nn.LSTM(input_dim, hidden_dim, num_layers=num_layers, batch_first=True)
        self.fc_peak = nn.Linear(hidden_dim, 1)
        self.fc_stop = nn.Sequential(nn.Linear(hidden_dim, 1), nn.Sigmoid())
        self.flat = None
        self.to(torch.device(device))

    def flatten_(self) -> torch.Tensor:
        """Moves the eight parameters into one flat fp32 buffer (layout of include/plume_b200.h, = the
        ``named_parameters()`` order) and re-points every ``nn.Parameter`` at its slice, so the training kernel
        (csrc/lstm_train_kernels.cu) updates the tensors ``state_dict()`` returns.  Hidden size 32 only."""
        if self.hidden_dim != 32:
            raise ValueError("the training kernel implements the reference's hidden_dim=32 (train_lstm.py:85)")
        dev = next(self.parameters()).device
        if self.flat is not None and self.flat.device == dev:
            return self.flat
        flat = torch.zeros(_lib.LSTM_TRAIN_PARAMS, dtype=torch.float32, device=dev)
        named = dict(self.named_parameters())
        for name, (off, shape) in _lib.LSTM_TRAIN_OFFSETS.items():
            p = named[name]
            n = int(np.prod(shape))
            flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = flat[off:off + n].view(shape)
        self.flat = flat
        return flat

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        if getattr(self, "flat", None) is not None:     # moved / cast: rebuild the flat view on the new storage
            self.flat = None
            self.flatten_()
        return self

    def c_params(self, window: int = 20, threshold: float = 0.8) -> _lib.LstmParams:
        p = {k: v.detach() for k, v in self.named_parameters()}
        for v in p.values():
            assert v.is_contiguous() and v.dtype == torch.float32
        return _lib.LstmParams(self.hidden_dim, window, threshold, p["lstm.weight_ih_l0"].data_ptr(),
                               p["lstm.weight_hh_l0"].data_ptr(), p["lstm.bias_ih_l0"].data_ptr(),
                               p["lstm.bias_hh_l0"].data_ptr(), p["fc_peak.weight"].data_ptr(),
                               p["fc_peak.bias"].data_ptr(), p["fc_stop.0.weight"].data_ptr(),
                               p["fc_stop.0.bias"].data_ptr())

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """``x`` [B,T,1] or [B,T] -> ``(peak [B], stop_prob [B])``."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PeakAndStopPredictor.forward runs on CUDA only (no CPU fallback)")
        lib = _lib.load()
        if x.dim() == 3:
            x = x.squeeze(-1)
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, T = x.shape
        peak = torch.empty(B, dtype=torch.float32, device=dev)
        stop = torch.empty(B, dtype=torch.float32, device=dev)
        lp = self.c_params(T)
        with torch.cuda.device(dev):
            # hidden 32 / 64: weights resident in shared memory; any other size <= 256: generic kernel (L2 weights)
            rc = lib.plume_lstm_stop_head(lp.w_ih, lp.w_hh, lp.b_ih, lp.b_hh, lp.w_peak, lp.b_peak, lp.w_stop,
                                          lp.b_stop, self.hidden_dim, x.data_ptr(), B, T, peak.data_ptr(),
                                          stop.data_ptr(), _stream(dev))
            _lib.check(rc, "plume_lstm_stop_head")
        return peak, stop


class ConcentrationThresholdPredictor(nn.Module):
    """V2.0 threshold predictor (PPOV2.0/model.py:203-240): 3-layer LSTM(1 -> hidden),
    FC hidden->64 -> LayerNorm -> ReLU -> (Dropout) -> 1, xavier init.  The LSTM stack runs in
    csrc/lstm_kernels.cu (``plume_lstm_forward``)."""

    def __init__(self, input_size: int = 1, hidden_size: int = 128, device="cuda"):
        super().__init__()
        self.hidden_size = hidden_size
        self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=3, batch_first=True,
                            dropout=0.3)
        self.fc = nn.Sequential(nn.Linear(hidden_size, 64), nn.LayerNorm(64), nn.ReLU(), nn.Dropout(0.1),
                                nn.Linear(64, 1))
        for name, p in self.named_parameters():
            if "weight" in name and p.dim() > 1:
                nn.init.xavier_uniform_(p)
            elif "bias" in name:
                nn.init.zeros_(p)
        self.to(torch.device(device))

    @torch.no_grad()
    def lstm_last_hidden(self, x: torch.Tensor) -> torch.Tensor:
        dev = next(self.parameters()).device
        lib = _lib.load()
        if x.dim() == 3:
            x = x.squeeze(-1)
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, T = x.shape
        flat = torch.cat([p.detach().reshape(-1) for p in self.lstm.parameters()]).contiguous()
        h = torch.empty(B, self.hidden_size, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.plume_lstm_forward(flat.data_ptr(), 3, self.hidden_size, x.data_ptr(), B, T,
                                              h.data_ptr(), _stream(dev)), "plume_lstm_forward")
        return h

    @torch.no_grad()
    def forward(self, x: torch.Tensor, lengths=None) -> torch.Tensor:
        """``x`` [B,T,1] (or [B,T]) -> predicted source concentration [B] (PPOV2.0/model.py:229-240).
        ``lengths`` is accepted for signature parity; the evaluator always passes full windows
        (PPOV2.0/evaluate_with_lstm.py:26), which is what the kernels implement."""
        T = x.shape[1]
        if lengths is not None and any(int(l) != T for l in lengths):
            raise NotImplementedError("ragged windows: the reference evaluator only passes lengths == window size")
        h = self.lstm_last_hidden(x)
        dev = h.device
        lib = _lib.load()
        out = torch.empty(h.shape[0], dtype=torch.float32, device=dev)
        fc1, ln, fc2 = self.fc[0], self.fc[1], self.fc[4]
        with torch.cuda.device(dev):
            _lib.check(lib.plume_threshold_head(h.data_ptr(), h.shape[0], self.hidden_size, fc1.weight.data_ptr(),
                                                fc1.bias.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(),
                                                fc2.weight.data_ptr(), fc2.bias.data_ptr(), out.data_ptr(),
                                                _stream(dev)), "plume_threshold_head")
        return out
