"""CPU: host-side logic of the mirrors (config defaults = the reference's, dispatch, task partition)."""
import pytest


def test_config_defaults_match_reference():
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig
    from mtrl_b200.config.nn import MultiHeadConfig, NeuralNetworkConfig
    from mtrl_b200.config.optim import OptimizerConfig
    from mtrl_b200.config.rl import OffPolicyTrainingConfig
    from mtrl_b200.rl.algorithms import MTSACConfig

    n = NeuralNetworkConfig()
    assert (n.width, n.depth, n.use_bias) == (400, 3, True)            # config/nn.py:9-25
    o = OptimizerConfig()
    assert (o.lr, o.max_grad_norm, o.eps) == (3e-4, None, None)         # config/optim.py:15-20
    assert o.spawn().eps == 1e-5                                        # config/optim.py:29-32
    p = ContinuousActionPolicyConfig()
    assert (p.log_std_min, p.log_std_max, p.squash_tanh) == (-20.0, 2.0, True)
    c = MTSACConfig(num_tasks=10)
    assert (c.num_critics, c.tau, c.gamma, c.initial_temperature, c.use_task_weights) == (2, 0.005, 0.99, 1.0, False)
    assert c.temperature_optimizer_config.max_grad_norm is None          # mtsac.py:120
    t = OffPolicyTrainingConfig(total_steps=1)
    assert (t.batch_size, t.buffer_size, t.warmstart_steps) == (1280, int(1e6), 4000)
    assert MultiHeadConfig(num_tasks=3).num_tasks == 3


def test_arch_dispatch_raises_on_base_config():
    from mtrl_b200.config.nn import MultiHeadConfig, NeuralNetworkConfig, VanillaNetworkConfig
    from mtrl_b200.nn import MultiHeadNetwork, VanillaNetwork, get_nn_arch_for_config

    assert get_nn_arch_for_config(MultiHeadConfig(num_tasks=2)) is MultiHeadNetwork
    assert get_nn_arch_for_config(VanillaNetworkConfig()) is VanillaNetwork
    with pytest.raises(ValueError):  # mtrl/nn/__init__.py:41-42
        get_nn_arch_for_config(NeuralNetworkConfig())


def test_task_partition_mt50():
    from mtrl_b200.rl.algorithms.mtsac import task_partition

    assert [b - a for a, b in task_partition(50, 8)] == [7, 7, 6, 6, 6, 6, 6, 6]
    assert task_partition(50, 2) == [(0, 25), (25, 50)]
    assert [b - a for a, b in task_partition(50, 4)] == [13, 13, 12, 12]
    parts = task_partition(10, 3)
    assert parts[0][0] == 0 and parts[-1][1] == 10 and all(parts[i][1] == parts[i + 1][0] for i in range(2))


def test_multihead_init_shapes_and_bounds():
    import math

    import torch

    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.nn.multi_head import MultiHeadNetwork, uniform

    net = MultiHeadNetwork(config=MultiHeadConfig(num_tasks=10, width=400), head_dim=8, head_kernel_init=uniform(1e-3),
                           head_bias_init=uniform(1e-3))
    p = net.init(torch.Generator().manual_seed(0), 49)
    assert p["layer_0"]["kernel"].shape == (49, 400) and p["layer_2"]["kernel"].shape == (400, 400)
    assert p["VmapDense_0"]["kernel"].shape == (10, 400, 8) and p["VmapDense_0"]["bias"].shape == (10, 8)
    assert float(p["layer_0"]["kernel"].abs().max()) <= math.sqrt(6 / 49)      # he_uniform
    assert float(p["layer_0"]["bias"].abs().max()) == 0.0
    assert float(p["VmapDense_0"]["kernel"].abs().max()) <= 1e-3                # networks.py:33-34
    total = sum(v.numel() for d in p.values() for v in d.values())
    assert total == 372_880
    ens = net.init(torch.Generator().manual_seed(0), 53, ensemble=2)
    assert ens["layer_0"]["kernel"].shape == (2, 53, 400)


def test_optimizer_configs_outside_the_path_fail_loudly():
    from mtrl_b200.config.optim import OptimizerConfig, PCGradConfig
    from mtrl_b200.config.utils import Optimizer

    with pytest.raises(NotImplementedError):
        OptimizerConfig(optimizer=Optimizer.SGD).spawn()
    from mtrl_b200.config.optim import DummyMultiTaskConfig

    assert PCGradConfig(num_tasks=3).requires_split_task_losses
    spec = PCGradConfig(num_tasks=3, max_grad_norm=1.0).spawn()   # chain(pcgrad, clip, adam) as data (optim.py:71-75)
    assert spec.pcgrad and spec.max_grad_norm == 1.0 and spec.eps == 1e-5
    assert not OptimizerConfig().spawn().pcgrad
    with pytest.raises(NotImplementedError):
        DummyMultiTaskConfig().spawn()
