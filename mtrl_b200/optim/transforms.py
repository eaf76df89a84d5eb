"""optax-protocol objects over `mtrl_task_combine` (see the package docstring)."""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, NamedTuple

import torch

from .. import _lib as L

_vp = C.c_void_p
L._EXTRA_DECLS.update({
    "mtrl_task_combine": ([C.c_int, _vp, C.c_longlong, C.c_int, C.c_longlong, _vp, C.c_int, _vp, _vp, _vp],),
})

KIND_PCGRAD, KIND_CAGRAD, KIND_GRADNORM, KIND_DUMMY = range(4)


class GradientTransformation(NamedTuple):
    """optax.GradientTransformation / GradientTransformationExtraArgs: a pair of pure functions."""
    init: Callable[[Any], Any]
    update: Callable[..., tuple[Any, Any]]


class PCGradState(NamedTuple):          # mtrl/optim/pcgrad.py:12-17
    n_grad_conflicts: torch.Tensor
    avg_grad_magnitude: torch.Tensor
    avg_grad_magnitude_before_surgery: torch.Tensor
    avg_cosine_similarity: torch.Tensor
    avg_cosine_similarity_diff: torch.Tensor


class CAGradState(NamedTuple):          # mtrl/optim/cagrad.py:13-18
    task_weights: torch.Tensor
    avg_grad_magnitude: torch.Tensor
    avg_grad_magnitude_before_surgery: torch.Tensor
    cagrad_objective: torch.Tensor


class GradNormState(NamedTuple):        # mtrl/optim/gradnorm.py:18-24 (the fields the reference logs)
    task_weights: torch.Tensor
    grad_magnitude: torch.Tensor
    avg_grad_magnitude_per_task: torch.Tensor


def _leaves(tree) -> list:
    if isinstance(tree, dict):
        out = []
        for k in tree:
            out += _leaves(tree[k])
        return out
    return [tree]


def _unflatten_like(flat: torch.Tensor, tree):
    """Inverse of raveling `tree` without its task axis (jax.flatten_util.ravel_pytree's unravel_fn)."""
    off = 0

    def build(t):
        nonlocal off
        if isinstance(t, dict):
            return {k: build(v) for k, v in t.items()}
        n = t[0].numel()
        leaf = flat[off:off + n].reshape(t.shape[1:])
        off += n
        return leaf
    return build(tree)


def _combine(kind: int, updates, num_tasks: int, perm=None, clip_per_task: bool = False):
    """Ravel the per-task update pytree to (T, P), run the transformation, unravel.  Returns (updates, stats[4], tw[T])."""
    leaves = _leaves(updates)
    if not leaves:
        raise ValueError("empty update pytree")
    dev = leaves[0].device
    if dev.type != "cuda":
        raise L.MtrlError("mtrl_b200.optim needs CUDA tensors; there is no CPU fallback")
    for x in leaves:
        if x.shape[0] != num_tasks:      # chex.assert_tree_shape_prefix(updates, (num_tasks,)), pcgrad.py:46
            raise ValueError(f"every leaf needs a leading task axis of {num_tasks}, got {tuple(x.shape)}")
    T = num_tasks
    if T > 64:
        raise NotImplementedError("at most 64 tasks")
    P = sum(x[0].numel() for x in leaves)
    ld = -(-P // 4) * 4
    rows = torch.zeros(T, ld, dtype=torch.float32, device=dev)
    off = 0
    for x in leaves:
        n = x[0].numel()
        rows[:, off:off + n] = x.reshape(T, n).to(torch.float32)
        off += n
    out = torch.empty(ld, dtype=torch.float32, device=dev)
    scratch = torch.zeros(T * T + 2 * T + 4, dtype=torch.float32, device=dev)
    p = None
    if perm is not None:
        p = torch.as_tensor(perm).to(device=dev, dtype=torch.int32).contiguous()
        if sorted(p.tolist()) != list(range(T)):
            raise ValueError("perm must be a permutation of range(num_tasks)")
    with torch.cuda.device(dev):
        L.check(L.lib().mtrl_task_combine(kind, _vp(rows.data_ptr()), ld, T, ld, _vp(p.data_ptr() if p is not None else None),
                                          int(clip_per_task), _vp(out.data_ptr()), _vp(scratch.data_ptr()),
                                          _vp(L.current_stream_ptr())))
    stats = scratch[T * T + T: T * T + T + 4]
    tw = scratch[T * T + T + 4: T * T + 2 * T + 4]
    return _unflatten_like(out, updates), stats, tw


def dummy_multitask_optimizer() -> GradientTransformation:
    """mtrl/optim/dummy.py:5-20: the mean of the per-task updates."""
    def init(params) -> dict:
        del params
        return {}

    def update(updates, state, params=None, **extra_args):
        del state, params, extra_args
        T = _leaves(updates)[0].shape[0]
        new, _, _ = _combine(KIND_DUMMY, updates, T)
        return new, {}

    return GradientTransformation(init=init, update=update)


def _perm_from_key(key, num_tasks: int):
    if key is None:
        raise AssertionError("RNG key must be provided")   # pcgrad.py:51-52
    gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
    return torch.randperm(num_tasks, generator=gen)


def pcgrad(num_tasks: int, cosine_sim_logs: bool = False) -> GradientTransformation:
    """mtrl/optim/pcgrad.py:20-136.  `update(updates, state, params, key=...)`; `perm=` overrides the permutation the
    reference draws from the key (:79).  cosine_sim_logs is not computed here (NaN, as the reference's default)."""
    nan = lambda dev: torch.full((), float("nan"), device=dev)  # noqa: E731

    def init(params) -> PCGradState:
        dev = _leaves(params)[0].device if _leaves(params) else "cuda"
        z = torch.zeros((), device=dev)
        return PCGradState(z, z.clone(), z.clone(), nan(dev), nan(dev))

    def update(updates, state, params=None, **extra_args):
        del state
        assert params is not None                                     # pcgrad.py:47
        perm = extra_args.get("perm")
        if perm is None:
            perm = _perm_from_key(extra_args.get("key"), num_tasks)
        new, stats, _ = _combine(KIND_PCGRAD, updates, num_tasks, perm=perm)
        dev = stats.device
        return new, PCGradState(stats[0], stats[1], stats[2], nan(dev), nan(dev))

    return GradientTransformation(init=init, update=update)


def cagrad(num_tasks: int, c: float = 0.5, num_iterations: int = 21, learning_rate: float | None = None,
           momentum: float = 0.5) -> GradientTransformation:
    """mtrl/optim/cagrad.py:20-237 with the reference's defaults (the fused kernel implements exactly those)."""
    default_lr = 25.0 if num_tasks < 50 else 50.0
    if (c, num_iterations, momentum) != (0.5, 21, 0.5) or (learning_rate is not None and learning_rate != default_lr):
        raise NotImplementedError("cagrad is implemented with the reference's defaults (c=0.5, 21 iterations, momentum 0.5)")

    def init(params) -> CAGradState:
        dev = _leaves(params)[0].device if _leaves(params) else "cuda"
        z = torch.zeros((), device=dev)
        return CAGradState(torch.full((num_tasks,), 1.0 / num_tasks, device=dev), z, z.clone(), z.clone())

    def update(updates, state, params=None, **extra_args):
        del state, params, extra_args
        new, stats, tw = _combine(KIND_CAGRAD, updates, num_tasks)
        return new, CAGradState(tw.clone(), stats[0], stats[1], stats[2])

    return GradientTransformation(init=init, update=update)


def gradnorm(optim=None, num_tasks: int = 1, asymmetry: float = 0.12, initial_weights=None,
             max_grad_norm: float | None = None) -> GradientTransformation:
    """mtrl/optim/gradnorm.py:61-163.  As written there the gradnorm loss does not depend on the task weights (:134-142), so
    their gradient is zero, the inner optimiser never moves them from their normalised initial value 1 and the
    transformation is the SUM of the per-task updates, each clipped to unit norm first when `max_grad_norm` is set."""
    del optim, asymmetry
    if initial_weights is not None:
        raise NotImplementedError("gradnorm initial_weights other than ones")

    def init(params) -> GradNormState:
        dev = _leaves(params)[0].device if _leaves(params) else "cuda"
        z = torch.zeros((), device=dev)
        return GradNormState(torch.ones(num_tasks, device=dev), z, z.clone())

    def update(updates, state, params=None, **extra_args):
        del state, params, extra_args          # task_losses only feed bookkeeping that cannot move the weights
        new, stats, _ = _combine(KIND_GRADNORM, updates, num_tasks, clip_per_task=bool(max_grad_norm))
        return new, GradNormState(torch.ones(num_tasks, device=stats.device), stats[0], stats[1])

    return GradientTransformation(init=init, update=update)
