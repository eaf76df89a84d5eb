"""`MLP` / `VanillaNetwork` (mtrl/nn/base.py:11-89) as descriptions: the networks of the single-task SAC baseline
(`mtrl_b200.rl.algorithms.SAC`) and of MT-PPO on `VanillaNetworkConfig`.  Their forward / backward run inside the fused
updates (csrc/sac.cu with MTRL_VARIANT_SAC, csrc/ppo.cu): the trunk `layer_0 .. layer_{depth-1}` as tcgen05 GEMMs, the
output Dense `layer_{depth}` as the single "head"; `use_layer_norm` / `use_skip_connections` (base.py:35-53) are the junction kernels
between those GEMMs (csrc/ln_kernels.cuh), with `LayerNorm_k/{scale,bias}` in the Flax parameter tree."""
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
