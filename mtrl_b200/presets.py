"""Experiment presets mirroring the reference's config scripts (experiments/mt10_mtmhsac.py:28-60,
experiments/width_scaling/mt50_mtmhsac_v2_2048.py:28-64): MultiHeadConfig actor/critic of a given
width with OptimizerConfig(max_grad_norm=1.0), two critics, Meta-World shapes."""
from __future__ import annotations

from .config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
from .config.nn import MultiHeadConfig
from .config.optim import OptimizerConfig
from .rl.algorithms.mtsac import MTSACConfig

METAWORLD_OBS = 39   # hand 3 + gripper 1 + object 14, two frames, + goal 3 (mtrl/envs/metaworld.py:34-100)
METAWORLD_ACT = 4    # mtrl/envs/metaworld.py:26-30


class _Space:
    def __init__(self, shape):
        self.shape = shape


class EnvSpec:
    """The two attributes MTSAC.initialize reads from an EnvConfig (mtsac.py:157-196)."""

    def __init__(self, obs_dim: int, action_dim: int):
        self.observation_space = _Space((obs_dim,))
        self.action_space = _Space((action_dim,))


def metaworld_mtmhsac(num_tasks: int, width: int, clip: bool = False) -> tuple[MTSACConfig, EnvSpec]:
    opt = OptimizerConfig(max_grad_norm=1.0)
    net = MultiHeadConfig(width=width, num_tasks=num_tasks, optimizer=opt)
    cfg = MTSACConfig(num_tasks=num_tasks, gamma=0.99, clip=clip,
                      actor_config=ContinuousActionPolicyConfig(network_config=net),
                      critic_config=QValueFunctionConfig(network_config=net),
                      num_critics=2, use_task_weights=False)
    return cfg, EnvSpec(METAWORLD_OBS + num_tasks, METAWORLD_ACT)
