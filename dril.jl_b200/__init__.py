"""dril_b200 — B200-native PPO rollout-and-update hot path of DRiL.jl behind the reference's API.

Everything computes in libdril_b200.so (hand-written CUDA, sm_100a). This package is only the
host-side mirror of the reference interface; importing it without the built library or using
it without a CUDA device raises (no CPU fallback)."""
from ._lib import DrilError, IterStats, NormCfg, PPOHyper, load  # noqa: F401
from .spaces import Box, Discrete  # noqa: F401
from .core import (Context, CudaBatchedEnv, DevicePolicy, NormalizeConfig, RolloutBuffer, gae_raw, set_option)  # noqa: F401
from .api import (AbstractCallback, AbstractTrainingLogger, ActorCriticLayer, Agent, BroadcastedParallelEnv,  # noqa: F401
                  ContinuousActorCriticLayer, DictLogger, DiscreteActorCriticLayer, MonitorWrapperEnv,
                  MultiThreadedParallelEnv, NeuralPolicy, NormalizeWrapperEnv, NormWrapperPolicy, NoTrainingLogger, PPO, ScalingWrapperEnv,
                  collect_rollout, evaluate_agent, extract_policy, get_action_and_values, get_hparams,
                  load_normalization_stats, load_policy_params_and_state, predict_actions, predict_values,
                  save_normalization_stats, save_policy_params_and_state, steps_taken, sync_normalization_stats, to_env, train)
