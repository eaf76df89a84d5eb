"""Mirror of `mtrl.optim` (/root/reference/mtrl/optim/{dummy,pcgrad,gradnorm,cagrad}.py): the reference's multi-task
gradient transformations with the optax protocol they are written against,

    tx.init(params) -> state
    tx.update(updates, state, params=None, **extra_args) -> (updates, state)

where `updates` is a pytree (nested dicts) whose leaves carry a leading task axis `(num_tasks, ...)` -- the per-task
gradients `jax.vmap(jax.value_and_grad(loss))` produces (mtsac.py:568-585, 677-687) -- and the returned `updates` has the
leaves' own shapes.  Leaves are CUDA tensors; the arithmetic is the library's coefficient-space kernels (Gram matrix of
the raveled rows -> T weights -> weighted row sum, C entry `mtrl_task_combine`), i.e. the same code the fused update runs
when an `OptimizerConfig` subclass puts one of these in front of clip + adam.  There is no CPU path.

`extra_args` follow the reference: pcgrad needs `key` (the row permutation of pcgrad.py:79 is drawn from it; here a
`torch.Generator`, an int seed, or an explicit permutation under `perm`), gradnorm takes `task_losses` (only bookkeeping in
the reference, see gradnorm.py:134-142).
"""
from .transforms import (  # noqa: F401
    CAGradState,
    GradientTransformation,
    GradNormState,
    PCGradState,
    cagrad,
    dummy_multitask_optimizer,
    gradnorm,
    pcgrad,
)
