"""Mirror of mtrl/rl/algorithms/__init__.py:9-19 for the accelerated paths."""
from .mtsac import MTSAC, MTSACConfig
from .mtppo import MTPPO, MTPPOConfig
from .sac import SAC, SACConfig


def get_algorithm_for_config(config):
    if type(config) is MTSACConfig:
        return MTSAC
    if type(config) is SACConfig:
        return SAC
    if type(config) is MTPPOConfig:
        return MTPPO
    raise ValueError(f"Unknown algorithm config type: {type(config)}")


__all__ = ["MTSAC", "MTSACConfig", "SAC", "SACConfig", "MTPPO", "MTPPOConfig", "get_algorithm_for_config"]
