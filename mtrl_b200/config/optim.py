"""Mirror of mtrl/config/optim.py:14-124.  `spawn()` returns the description of
optax.chain(clip_by_global_norm(max_grad_norm), adam(lr, eps)) that the fused CUDA optimiser
(csrc/sac_kernels.cuh adam_kernel) executes; the multi-task configs (optim.py:46-124: DummyMultiTaskConfig, PCGradConfig,
GradNormConfig, CAGradConfig) put their transformation in front of the same chain (per-task gradients of the split losses
+ the transformation in coefficient space).  The transformations themselves are also available as optax-protocol
objects in `mtrl_b200.optim`."""
from dataclasses import dataclass

from .utils import Optimizer


@dataclass(frozen=True)
class AdamChainSpec:
    """What OptimizerConfig.spawn() would build with optax, as data."""

    lr: float
    eps: float
    b1: float = 0.9
    b2: float = 0.999
    max_grad_norm: float | None = None
    pcgrad: bool = False   # optax.chain(pcgrad(num_tasks), clip, adam): mtrl/config/optim.py:62-76
    cagrad: bool = False   # optax.chain(cagrad(num_tasks), clip, adam): mtrl/config/optim.py:104-124
    gradnorm: bool = False           # optax.chain(gradnorm(...), clip, adam): mtrl/config/optim.py:79-102
    gradnorm_clip_per_task: bool = False
    dummy: bool = False    # optax.chain(dummy_multitask_optimizer(), clip, adam): mtrl/config/optim.py:46-59


@dataclass(frozen=True, kw_only=True)
class OptimizerConfig:
    lr: float = 3e-4
    optimizer: Optimizer = Optimizer.Adam
    max_grad_norm: float | None = None
    eps: float | None = None
    weight_decay: float | None = None

    @property
    def requires_split_task_losses(self) -> bool:
        return False

    def spawn(self) -> AdamChainSpec:
        if self.optimizer != Optimizer.Adam:
            raise NotImplementedError(
                f"{self.optimizer}: only Adam (+ global-norm clip) is on the accelerated path; no experiment of the "
                "reference's MT-SAC grid uses another optimiser")
        eps = self.eps if self.eps is not None else 1e-5  # optim.py:29-32
        return AdamChainSpec(lr=self.lr, eps=eps, max_grad_norm=self.max_grad_norm)


@dataclass(frozen=True, kw_only=True)
class DummyMultiTaskConfig(OptimizerConfig):   # optim.py:46-59
    @property
    def requires_split_task_losses(self) -> bool:
        return True

    def spawn(self) -> AdamChainSpec:
        """optax.chain(dummy_multitask_optimizer(), OptimizerConfig.spawn()) as data: the dummy transformation averages
        the per-task gradients (mtrl/optim/dummy.py:18).  Those are gradients of the reference's SPLIT losses, which differ
        from the un-split loss (a' sampled on data.observations, mtsac.py:515-523; explore term in the actor loss,
        :631-637, 676-682), so the fused update takes its split path for this config as well."""
        import dataclasses

        return dataclasses.replace(OptimizerConfig.spawn(self), dummy=True)


@dataclass(frozen=True, kw_only=True)
class PCGradConfig(OptimizerConfig):
    num_tasks: int
    cosine_sim_logs: bool = False

    @property
    def requires_split_task_losses(self) -> bool:
        return True

    def spawn(self) -> AdamChainSpec:
        """optax.chain(pcgrad(num_tasks, cosine_sim_logs), OptimizerConfig.spawn()) (optim.py:71-75) as data; the fused
        update runs pcgrad in coefficient space over the per-task Gram matrix (csrc/sac_kernels.cuh)."""
        import dataclasses

        return dataclasses.replace(OptimizerConfig.spawn(self), pcgrad=True)


@dataclass(frozen=True, kw_only=True)
class CAGradConfig(OptimizerConfig):   # optim.py:104-124
    num_tasks: int
    cagrad_optimizer: OptimizerConfig | None = None   # carried but unused by the reference's spawn() (:117-123)
    initial_weights: object | None = None
    max_grad_norm: float | None = None

    @property
    def requires_split_task_losses(self) -> bool:
        return True

    def spawn(self) -> AdamChainSpec:
        """optax.chain(cagrad(num_tasks), OptimizerConfig.spawn()) as data; cagrad runs with its defaults (c = 0.5,
        21 iterations, lr 25 / 50, momentum 0.5: mtrl/optim/cagrad.py:20-41) on the per-task Gram matrix."""
        import dataclasses

        return dataclasses.replace(OptimizerConfig.spawn(self), cagrad=True)


@dataclass(frozen=True, kw_only=True)
class GradNormConfig(OptimizerConfig):   # optim.py:79-102
    num_tasks: int
    gradnorm_optimizer: OptimizerConfig | None = None
    initial_weights: object | None = None
    asymmetry: float = 0.12
    max_grad_norm: float | None = None

    @property
    def requires_split_task_losses(self) -> bool:
        return True

    def spawn(self) -> AdamChainSpec:
        """optax.chain(gradnorm(optim, num_tasks, asymmetry, initial_weights, max_grad_norm), OptimizerConfig.spawn())
        as data.  As written in the reference the task weights receive a zero gradient (gradnorm.py:134-142), stay at
        their normalised initial value, and the transformation is the weighted SUM of the per-task gradients."""
        import dataclasses

        if self.initial_weights is not None:
            raise NotImplementedError("GradNormConfig.initial_weights other than ones")
        return dataclasses.replace(OptimizerConfig.spawn(self), gradnorm=True, gradnorm_clip_per_task=bool(self.max_grad_norm))
