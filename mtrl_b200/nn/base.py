"""`MLP` / `VanillaNetwork` (mtrl/nn/base.py:11-89): kept as descriptions for the single-task SAC
baseline; their fused path is a SURVEY 8(f) "next" row and is not built yet."""
from dataclasses import dataclass

from ..config.nn import VanillaNetworkConfig


@dataclass
class MLP:
    head_dim: int
    depth: int = 3
    width: int = 400
    use_skip_connections: bool = False
    use_layer_norm: bool = False


@dataclass
class VanillaNetwork:
    config: VanillaNetworkConfig
    head_dim: int
    head_kernel_init: object = None
    head_bias_init: object = None
    activate_last: bool = False
