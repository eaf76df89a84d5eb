"""Mirror of mtrl/nn/__init__.py:18-42: config type -> network class.  Only the two architectures
on the MT-SAC hot path exist here; the dispatch keeps the reference's behaviour of raising on the
base NeuralNetworkConfig (nn/__init__.py:41-42)."""
from .. import config as _config  # noqa: F401
from ..config import nn as _nn
from .multi_head import MultiHeadNetwork
from .base import MLP, VanillaNetwork


def get_nn_arch_for_config(config):
    if type(config) is _nn.MultiHeadConfig:
        return MultiHeadNetwork
    if type(config) is _nn.VanillaNetworkConfig:
        return VanillaNetwork
    raise ValueError(
        f"Unknown config type: {type(config)}. (NeuralNetworkConfig by itself is not supported, use VanillaNeworkConfig)")


__all__ = ["VanillaNetwork", "MultiHeadNetwork", "MLP", "get_nn_arch_for_config"]
