"""settlers_of_catan_rl_b200 — B200-native vectorised Catan self-play engine.

Only the env-step + PPO-rollout hot path of henrycharlesworth/settlers_of_catan_RL lives here:
``VecCatanEnv`` (vector env on CUDA tensors), ``EnvWrapper`` (single-env adapter with the reference's
surface), and the rollout kernels (``gae``, ``normalise_advantages``).  The CUDA library is loaded
lazily; there is no CPU implementation.
"""
from . import layout  # noqa: F401

__all__ = ["layout", "VecCatanEnv", "EnvWrapper", "gae", "normalise_advantages", "RolloutStorage", "PolicyInputs", "SeatPolicies",
           "CatanPolicy", "SelfPlayTrainer", "PPOConfig"]


def __getattr__(name):
    if name == "VecCatanEnv":
        from .vec_env import VecCatanEnv
        return VecCatanEnv
    if name == "EnvWrapper":
        from .env_wrapper import EnvWrapper
        return EnvWrapper
    if name in ("gae", "normalise_advantages", "RolloutStorage"):
        from . import rollout
        return getattr(rollout, name)
    if name == "PolicyInputs":
        from .policy_io import PolicyInputs
        return PolicyInputs
    if name == "SeatPolicies":
        from .self_play import SeatPolicies
        return SeatPolicies
    if name == "CatanPolicy":
        from .policy_net import CatanPolicy
        return CatanPolicy
    if name in ("SelfPlayTrainer", "PPOConfig"):
        from . import ppo
        return getattr(ppo, name)
    raise AttributeError(name)
