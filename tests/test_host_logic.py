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
    from mtrl_b200.config.optim import CAGradConfig, GradNormConfig

    d = DummyMultiTaskConfig(max_grad_norm=1.0)     # chain(dummy_multitask_optimizer(), clip, adam): takes the split path
    assert d.requires_split_task_losses and d.spawn().dummy and d.spawn().max_grad_norm == 1.0
    assert CAGradConfig(num_tasks=3).spawn().cagrad and CAGradConfig(num_tasks=3).requires_split_task_losses
    g = GradNormConfig(num_tasks=3, max_grad_norm=1.0).spawn()
    assert g.gradnorm and g.gradnorm_clip_per_task and not GradNormConfig(num_tasks=3).spawn().gradnorm_clip_per_task


@pytest.mark.parametrize("T,W,depth,E", [(50, 2048, 3, 2), (10, 400, 3, 2), (10, 256, 2, 1), (7, 100, 4, 3), (50, 4096, 3, 2)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_trunk_ownership_segments_tile_the_trunk(T, W, depth, E, world):
    """csrc/sac.cu trunk_segments (host logic of the sharded exchange): every trunk element belongs to exactly one
    segment and one owner; hidden-layer kernels are cut into contiguous row blocks (the dW problems' outputs)."""
    import ctypes as C

    from mtrl_b200 import _lib as L
    from mtrl_b200.rl.algorithms.mtsac import SacConfigC, SacLayoutC

    t_local = -(-T // world)
    cfg = SacConfigC(num_tasks=T, task_begin=0, num_local_tasks=t_local, obs_dim=39 + T, action_dim=4, width=W, depth=depth,
                     num_critics=E, max_rows=128 * t_local, max_batch=128 * t_local, gamma=0.99, tau=0.005, actor_lr=3e-4,
                     critic_lr=3e-4, alpha_lr=3e-4, adam_b1=0.9, adam_b2=0.999, adam_eps=1e-5, actor_max_grad_norm=1.0,
                     critic_max_grad_norm=1.0, alpha_max_grad_norm=-1.0, log_std_min=-20.0, log_std_max=2.0,
                     target_entropy=-4.0, clip_q=0, use_task_weights=0, noise_seed=1, variant=0)
    lay = SacLayoutC()
    L.check(L.lib().mtrl_sac_query_layout(C.byref(cfg), C.byref(lay)))
    fn = L.lib().mtrl_trunk_segments
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int)]
    for net in (lay.critic, lay.actor):
        out = (C.c_longlong * (4 * 512))()
        n = C.c_int()
        L.check(fn(C.byref(net), world, out, 512, C.byref(n)))
        segs = sorted((out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]) for i in range(n.value))
        pos = 0
        for b, e, owner, pre in segs:
            assert b == pos and e > b and b % 4 == 0 and e % 4 == 0, (b, e, pos)
            assert 0 <= owner < world
            pos = e
        assert pos == net.trunk_total
        pre_floats = sum(e - b for b, e, _, pre in segs if pre)
        hidden = net.members * (depth - 1) * W * W
        assert hidden <= pre_floats <= hidden + net.members * (depth - 1) * 32   # + alignment padding of the last block
        owned = [sum(e - b for b, e, o, _ in segs if o == r) for r in range(world)]
        if depth > 1 and W >= 32 * world:
            assert max(owned) <= 1.6 * (net.trunk_total / world) + 64 * W, owned   # balanced up to the small pieces


def test_arena_plan_offsets():
    from mtrl_b200.rl.algorithms.mtsac import COMM_HEADER_BYTES, arena_plan

    plan, size = arena_plan(1000, 333)
    offs = [plan[k] for k in ("critic_grads", "actor_grads", "critic_params", "actor_params")]
    assert offs[0] == COMM_HEADER_BYTES and all(o % 4096 == 0 for o in offs) and offs == sorted(offs)
    assert offs[1] - offs[0] >= 4000 and offs[2] - offs[1] >= 1332 and size - offs[3] >= 1332 and size % 4096 == 0
