"""ctypes binding of libplume_b200.so (the C ABI declared in include/plume_b200.h).

There is no CPU fallback: if the shared library has not been built, importing the
bindings raises, and every call fails loudly on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplume_b200.so")

OBS_DIM, NUM_ACTIONS, INFO_DIM, VISIT_STRIDE = 6, 5, 5, 104
INFO_KEYS = ("concentration_reward", "explore_reward", "move_penalty", "tke_penalty", "boundary_penalty")
FLAG_AUTO_RESET, FLAG_GREEDY, FLAG_STOP_TERMINATES, FLAG_DEFER_STOP_HEAD, FLAG_FAST_REWARD = 1, 2, 4, 8, 16
FLAG_STOP_FIXED, FLAG_STOP_THRESHOLD = 32, 64

# flat MLP parameter layout (include/plume_b200.h)
MLP_OFFSETS = {
    "feature.0.weight": (0, (256, 6)), "feature.0.bias": (1536, (256,)),
    "feature.1.weight": (1792, (256,)), "feature.1.bias": (2048, (256,)),
    "feature.3.weight": (2304, (128, 256)), "feature.3.bias": (35072, (128,)),
    "feature.4.weight": (35200, (128,)), "feature.4.bias": (35328, (128,)),
    "actor.weight": (35456, (5, 128)), "actor.bias": (36096, (5,)),
    "critic.weight": (36104, (1, 128)), "critic.bias": (36232, (1,)),
}
MLP_PARAMS = 36236
# flat layout of PeakAndStopPredictor(hidden 32) = named_parameters() order (include/plume_b200.h, N3)
LSTM_TRAIN_OFFSETS = {
    "lstm.weight_ih_l0": (0, (128, 1)), "lstm.weight_hh_l0": (128, (128, 32)), "lstm.bias_ih_l0": (4224, (128,)),
    "lstm.bias_hh_l0": (4352, (128,)), "fc_peak.weight": (4480, (1, 32)), "fc_peak.bias": (4512, (1,)),
    "fc_stop.0.weight": (4513, (1, 32)), "fc_stop.0.bias": (4545, (1,)),
}
LSTM_TRAIN_PARAMS = 4546

_vp = C.c_void_p


class EnvConfig(C.Structure):
    _fields_ = [("grid_size", C.c_int32), ("max_steps", C.c_int32), ("grid_divisions", C.c_int32),
                ("field_mode", C.c_int32), ("conc_peak", C.c_double), ("turbulence_intensity", C.c_double),
                ("sigma", C.c_double), ("clip_hi", C.c_double), ("conc_reward_coef", C.c_double),
                ("tke_penalty_factor", C.c_double), ("boundary_penalty", C.c_double),
                ("boundary_decay_start", C.c_double), ("initial_radius", C.c_double), ("seed", C.c_uint64),
                ("plume_model", C.c_int32), ("reserved0", C.c_int32)]


class EnvState(C.Structure):
    _fields_ = [("n_envs", C.c_int32), ("env_id_base", C.c_int32), ("pos_x", _vp), ("pos_y", _vp),
                ("src_x", _vp), ("src_y", _vp), ("step_count", _vp), ("episode_idx", _vp), ("visited", _vp),
                ("radius", _vp), ("explore_bonus", _vp), ("conc_field", _vp), ("tke_field", _vp),
                ("sin_tab", _vp), ("cos_tab", _vp), ("curriculum", _vp), ("last_move", _vp),
                ("step_frac_tab", _vp), ("visit_denom_tab", _vp), ("cell_tke", _vp), ("cell_conc", _vp), ("cell_key", _vp)]


class LstmParams(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("window", C.c_int32), ("threshold", C.c_float), ("w_ih", _vp),
                ("w_hh", _vp), ("b_ih", _vp), ("b_hh", _vp), ("w_peak", _vp), ("b_peak", _vp), ("w_stop", _vp),
                ("b_stop", _vp)]


class RolloutBuffers(C.Structure):
    _fields_ = [(n, _vp) for n in ("obs", "actions", "rewards", "values", "log_probs", "dones", "reached",
                                   "stop_prob", "stop_flag", "peak_pred", "trend", "info", "episode_idx",
                                   "forced_actions", "step_noise", "noise_out", "conc_window", "window_fill",
                                   "last_obs", "conc_sample", "fill_t", "src_dist", "pos_out", "src_out", "conc_out",
                                   "eval_ring", "stop_threshold", "flag_code")]


class PpoBatch(C.Structure):
    _fields_ = [("total", C.c_int64), ("obs", _vp), ("actions", _vp), ("old_log_probs", _vp),
                ("advantages", _vp), ("returns", _vp), ("old_values", _vp), ("packed", _vp)]


class TrajLog(C.Structure):
    _fields_ = [("max_episodes", C.c_int32), ("max_steps", C.c_int32), ("n_envs", C.c_int32), ("reserved", C.c_int32)] + \
               [(n, _vp) for n in ("x", "y", "conc", "steps", "source", "success", "radius", "sums", "final_conc",
                                   "c_x", "c_y", "c_conc", "c_sums", "c_len", "count")]


GAE_VARIANTS = {"quirk": 0, "bootstrap": 1, "v12": 2}
KERNEL_AUTO, KERNEL_TENSOR, KERNEL_SIMT = 0, 1, 2
KERNEL_PATHS = {"auto": KERNEL_AUTO, "tensor": KERNEL_TENSOR, "tc": KERNEL_TENSOR, "simt": KERNEL_SIMT,
                "cuda": KERNEL_SIMT}
CURRICULUM_MAX_WINDOWS = 8192
MODEL_ISOTROPIC, MODEL_DISPERSION = 0, 1
PLUME_MODELS = {"isotropic": MODEL_ISOTROPIC, "code": MODEL_ISOTROPIC, "dispersion": MODEL_DISPERSION,
                "readme": MODEL_DISPERSION}


def make_env_config(cfg, field_mode: int, seed: int, plume_model: int = 0) -> EnvConfig:
    return EnvConfig(cfg.grid_size, cfg.max_steps, cfg.grid_divisions, field_mode, cfg.conc_peak,
                     cfg.turbulence_intensity, cfg.sigma, cfg.clip_hi, cfg.conc_reward_coef,
                     cfg.tke_penalty_factor, cfg.boundary_penalty, cfg.boundary_decay_start,
                     cfg.initial_radius, seed & 0xFFFFFFFFFFFFFFFF, plume_model, 0)


_P = C.POINTER
_SIGNATURES = {
    "plume_abi_version": (C.c_int, []),
    "plume_last_error": (C.c_char_p, []),
    "plume_device_info": (C.c_int, [_P(C.c_int32)] * 3),
    "plume_env_reset": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, C.c_int32, _vp, _vp]),
    "plume_generate_fields": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, C.c_int32, _vp, _vp, _vp]),
    "plume_field_noise_at": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp]),
    "plume_field_at": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, _vp, _vp, _vp, _vp]),
    "plume_env_observe": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, _vp]),
    "plume_env_step": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, _vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp]),
    "plume_policy_forward": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "plume_policy_act": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, _vp, C.c_int32, _vp, _vp, C.c_uint32, _vp,
                                   _vp, _vp, _vp, _vp, _vp]),
    "plume_lstm_stop_head": (C.c_int, [_vp] * 8 + [C.c_int32, _vp, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "plume_stop_head_segment": (C.c_int, [_P(LstmParams), _vp, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, C.c_double,
                                          _vp, _vp, _vp, _vp, C.c_int32, _vp]),
    "plume_lstm_forward": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, _vp, _vp]),
    "plume_threshold_head": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plume_trend_features": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_double, _vp, _vp]),
    "plume_rollout": (C.c_int, [_P(EnvConfig), _P(EnvState), _vp, _P(LstmParams), _P(RolloutBuffers), C.c_int32,
                                C.c_uint32, _vp, _vp]),
    "plume_gae_scan": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_double, C.c_double, _vp, _vp, _vp]),
    "plume_gae_normalise": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "plume_gae_scan_variant": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                         _vp, _vp, _vp]),
    "plume_gae_normalise_variant": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int32, _vp, _vp]),
    "plume_ppo_grad": (C.c_int, [_vp, _P(PpoBatch), _vp, C.c_uint64, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_float, C.c_float, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, _vp]),
    "plume_ppo_update": (C.c_int, [_vp, _vp, _vp, _vp, _P(PpoBatch), _vp, C.c_uint64, C.c_int32, C.c_int64, C.c_int32,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int32, _vp]),
    "plume_ppo_workspace_bytes": (C.c_int64, [C.c_int64]),
    "plume_ppo_pack": (C.c_int, [_P(PpoBatch), _vp, _vp]),
    "plume_clip_adam": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_int32, _vp, _vp]),
    "plume_comm_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P(_vp), _vp]),
    "plume_comm_connect": (C.c_int, [_vp, _vp]),
    "plume_comm_destroy": (C.c_int, [_vp]),
    "plume_comm_error": (C.c_int, [_vp, _P(C.c_int32), _vp]),
    "plume_comm_error_async": (C.c_int, [_vp, _vp, _vp]),
    "plume_comm_reset": (C.c_int, [_vp]),
    "plume_comm_allreduce_small": (C.c_int, [_vp, _vp, C.c_int32, _vp]),
    "plume_comm_publish_codes": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "plume_curriculum_update_peer": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_double, C.c_double, C.c_double,
                                               C.c_double, C.c_int32, C.c_double, _vp, _vp]),
    "plume_allreduce_clip_adam": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_float, C.c_int32, _vp, _vp]),
    "plume_lstm_dataset": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_float, _vp,
                                     _vp, _vp]),
    "plume_lstm_train_workspace_bytes": (C.c_int64, [C.c_int32]),
    "plume_lstm_train_epoch": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                         _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "plume_permutation": (C.c_int, [C.c_int64, C.c_uint64, C.c_int32, C.c_int64, C.c_int64, _vp, _vp]),
    "plume_tc_gemm": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "plume_tc_gemm_f16": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "plume_curriculum_update": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, _vp, _vp, C.c_double, C.c_double,
                                          C.c_double, C.c_double, C.c_int32, C.c_double, _vp]),
    "plume_eval_collect": (C.c_int, [_P(RolloutBuffers), C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plume_trajectory_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "plume_trajectory_log": (C.c_int, [_P(TrajLog), _P(RolloutBuffers), C.c_int32, _vp, C.c_int32, _vp, C.c_double,
                                       _vp, C.c_int64, _vp]),
    "plume_curriculum_update_packed": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_double,
                                                 C.c_double, C.c_double, C.c_double, C.c_int32, C.c_double, _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
launch_count = 0     # kernels launched through this binding (bench.py's gpu_launches)


class PlumeLibraryError(RuntimeError):
    pass


def load():
    """Loads libplume_b200.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlumeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the plume kernels.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError = a declared symbol is missing
        fn.restype, fn.argtypes = res, args
    if lib.plume_abi_version() != 4:
        raise PlumeLibraryError("libplume_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().plume_last_error()
        raise PlumeLibraryError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
