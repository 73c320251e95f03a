"""uav-wrf-les-ppo-lstm_b200: B200-native vectorised rollout + PPO update for UAV methane-plume
tracing, behind the reference's environment / model / learner interfaces.

The package is imported as ``uav_wrf_les_ppo_lstm_b200`` (see the shim of that name).  All
arithmetic runs in ``libplume_b200.so`` (CUDA, sm_100a); there is no CPU fallback.
"""
from .config import PlumeConfig, config_for  # noqa: F401
from . import _lib  # noqa: F401
from .env import MethaneEnv, VecMethaneEnv  # noqa: F401
from .model import ConcentrationThresholdPredictor, PeakAndStopPredictor, PPOActorCritic  # noqa: F401
from .buffer import PPOBuffer  # noqa: F401
from .rollout import RolloutEngine  # noqa: F401
from .learner import FusedAdam, PPOTrainer, UpdateWorkspace, compute_advantages, update_model  # noqa: F401
from .trainer import PlumeTrainer  # noqa: F401
from .evaluate import EvalResult, ThresholdController, evaluate_policy  # noqa: F401
from .trajectory_log import TrajectoryLogger  # noqa: F401
from .lstm_train import LstmTrainer, build_dataset, eligible_episodes  # noqa: F401

_update_model = update_model   # the reference's private name (train_ppo2.0.py:14)

__all__ = ["PlumeConfig", "config_for", "MethaneEnv", "VecMethaneEnv", "PPOActorCritic", "PeakAndStopPredictor",
           "ConcentrationThresholdPredictor", "PPOBuffer", "RolloutEngine", "FusedAdam", "PPOTrainer",
           "UpdateWorkspace", "compute_advantages", "update_model", "PlumeTrainer", "EvalResult",
           "ThresholdController", "evaluate_policy", "TrajectoryLogger", "LstmTrainer", "build_dataset",
           "eligible_episodes"]
