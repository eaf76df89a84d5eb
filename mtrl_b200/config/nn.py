"""Mirror of mtrl/config/nn.py:7-65 (the network configs on the MT-SAC path)."""
from dataclasses import dataclass

from .optim import OptimizerConfig
from .utils import Activation, Initializer


@dataclass(frozen=True, kw_only=True)
class NeuralNetworkConfig:
    width: int = 400
    depth: int = 3
    kernel_init: Initializer = Initializer.HE_UNIFORM
    bias_init: Initializer = Initializer.ZEROS
    use_bias: bool = True
    activation: Activation = Activation.ReLU
    optimizer: OptimizerConfig = OptimizerConfig()


@dataclass(frozen=True, kw_only=True)
class VanillaNetworkConfig(NeuralNetworkConfig):
    use_skip_connections: bool = False
    use_layer_norm: bool = False


@dataclass(frozen=True, kw_only=True)
class MultiHeadConfig(NeuralNetworkConfig):
    num_tasks: int
