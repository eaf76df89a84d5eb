"""`MultiHeadNetwork` (mtrl/nn/multi_head.py:9-68) as a shape/initialiser description.  The forward
and backward passes run inside the fused update (csrc/sac.cu); this class only owns what the
reference's module owns on the host: parameter names, shapes and initialisers."""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from ..config.nn import MultiHeadConfig
from ..config.utils import Activation, Initializer


def uniform(bound: float):
    """mtrl/nn/initializers.py:4-10."""
    def _init(gen: torch.Generator, shape):
        return (torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(bound).float()
    return _init


def _kernel_init(kind: Initializer):
    if kind == Initializer.HE_UNIFORM:   # jax variance_scaling(2.0, "fan_in", "uniform")
        return lambda gen, shape: uniform(math.sqrt(6.0 / shape[-2]))(gen, shape)
    if kind == Initializer.HE_NORMAL:
        # jax he_normal draws a truncated normal; plain normal with the same variance is used here
        return lambda gen, shape: (torch.randn(*shape, generator=gen, dtype=torch.float64) * math.sqrt(2.0 / shape[-2])).float()
    if kind == Initializer.XAVIER_UNIFORM:
        return lambda gen, shape: uniform(math.sqrt(6.0 / (shape[-2] + shape[-1])))(gen, shape)
    if kind == Initializer.ZEROS:
        return lambda gen, shape: torch.zeros(*shape)
    raise NotImplementedError(f"kernel_init {kind}")


@dataclass
class MultiHeadNetwork:
    config: MultiHeadConfig
    head_dim: int
    head_kernel_init: object = None
    head_bias_init: object = None
    normalize_layer: bool = False
    skip_connection: bool = False

    def __post_init__(self):
        assert self.config.num_tasks is not None, "Number of tasks must be provided."  # multi_head.py:25
        if self.normalize_layer or self.skip_connection:
            raise NotImplementedError("normalize_layer / skip_connection default to off in every MT-SAC experiment")
        if self.config.activation != Activation.ReLU or not self.config.use_bias:
            raise NotImplementedError("the fused path implements Dense(use_bias=True) + ReLU (the reference default)")

    def init(self, gen: torch.Generator, in_dim: int, ensemble: int | None = None) -> dict:
        """Parameter tree with the Flax names of multi_head.py:34-62."""
        c = self.config
        lead = () if ensemble is None else (ensemble,)
        kinit = _kernel_init(c.kernel_init)
        binit = _kernel_init(c.bias_init) if c.bias_init != Initializer.ZEROS else (lambda gen, shape: torch.zeros(*shape))
        hk = self.head_kernel_init or _kernel_init(Initializer.HE_NORMAL)
        hb = self.head_bias_init or (lambda gen, shape: torch.zeros(*shape))
        p = {}
        d = in_dim
        for i in range(c.depth):
            p[f"layer_{i}"] = {"kernel": kinit(gen, lead + (d, c.width)), "bias": binit(gen, lead + (c.width,))}
            d = c.width
        p["VmapDense_0"] = {"kernel": hk(gen, lead + (c.num_tasks, c.width, self.head_dim)),
                            "bias": hb(gen, lead + (c.num_tasks, self.head_dim))}
        return p
