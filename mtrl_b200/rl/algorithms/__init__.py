"""Mirror of mtrl/rl/algorithms/__init__.py:9-19 for the accelerated path."""
from .mtsac import MTSAC, MTSACConfig


def get_algorithm_for_config(config):
    if type(config) is MTSACConfig:
        return MTSAC
    raise ValueError(f"Unknown algorithm config type: {type(config)}")


__all__ = ["MTSAC", "MTSACConfig", "get_algorithm_for_config"]
