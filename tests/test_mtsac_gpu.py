"""Fused CUDA MT-SAC update vs the fp64 oracle on identical batches, weights and noise.

Tolerance (north_star): updated parameters and losses within 1e-3 relative of the oracle -- checked
against the EXACT-arithmetic fp64 oracle.

Raw gradients and the Adam step (new - old) are checked twice:
  * against the exact oracle with a loose bound.  The trunk GEMMs take tf32 operands (what XLA's
    default f32 dot precision also does on NVIDIA GPUs), so ~2e-4 of the ReLU gates sit on the other
    side of zero than in exact arithmetic; each flipped gate changes its gradient term by 100 %, which
    shows up as ~sqrt(2e-4) ~ 1-3 % in the l2 norm of deep-layer gradients although every log scalar
    agrees to 1e-5.  (fp32 vs fp64 shows the same effect at ~3e-4.)
  * against the oracle run with tf32-rounded matmul operands (OracleConfig.matmul_operands="tf32"),
    where the gates agree: this isolates the kernels' arithmetic and must hold to a few 1e-3."""
import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O

pytestmark = pytest.mark.gpu

TOL_LOG = 1e-3       # north_star tolerance on the reported losses / norms
TOL_PARAM = 1e-3     # north_star tolerance on updated parameters
TOL_GRAD_EXACT = 8e-2   # gradient leaves vs exact arithmetic (ReLU-gate flips, see module docstring)
TOL_GRAD = 1e-2         # gradient leaves vs the tf32-operand oracle (dZ is re-rounded to tf32 at every layer)
TOL_DELTA = 1.5e-2      # the Adam step (new - old) of actor / critic vs the tf32-operand oracle
TOL_DELTA_EXACT = 0.25  # the Adam step vs exact arithmetic: the first Adam step is ~lr*sign(g), so every
                        # gradient element whose sign differs moves the step by 2*lr (zero-initialised biases
                        # are pure step, which is why they are compared here and not under TOL_PARAM)


def rms(x):
    return float(x.double().pow(2).mean().sqrt())


def flat(tree):
    return torch.cat([x.detach().double().flatten().cpu() for x in O.tree_leaves(tree)])


def agent_flat(agent_tree, oracle_tree, ens):
    return torch.cat([a.detach().double().flatten().cpu() for _, _, a in SU._pairs(oracle_tree, SU._net(agent_tree, ens))])


def run_case(cfg, per_task, seed=1, shuffle=False, counts=None, steps=1, check_grads=True):
    import dataclasses

    st = O.init_state(cfg, seed=seed, dtype=torch.float32)
    agent = SU.make_agent(cfg, per_task, seed=seed)
    SU.load_oracle_state(agent, st)
    st64 = st.to(torch.float64)
    cfg_tf32 = dataclasses.replace(cfg, matmul_operands="tf32")
    worst = {}
    for step in range(steps):
        batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=100 + step, dtype=torch.float32)
        if counts is not None:  # uneven rows per task: keep the first counts[t] rows of each task
            task = batch[0][:, -cfg.num_tasks:].argmax(1)
            keep = torch.zeros(task.shape[0], dtype=torch.bool)
            for t, n in enumerate(counts):
                keep[(task == t).nonzero().flatten()[:n]] = True
            batch, ec, ea = tuple(b[keep] for b in batch), ec[keep], ea[keep]
        if shuffle:
            perm = torch.randperm(batch[0].shape[0], generator=torch.Generator().manual_seed(5))
            batch, ec, ea = tuple(b[perm] for b in batch), ec[perm], ea[perm]
        b64 = tuple(b.double() for b in batch)
        old = st64
        if step == 0:
            t_new, _, t_grads, _ = O.mtsac_update(st64, b64, ec.double(), ea.double(), cfg_tf32, return_grads=True)
        st64, logs64, grads64, _ = O.mtsac_update(st64, b64, ec.double(), ea.double(), cfg, return_grads=True)
        _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
        for k in O.LOG_KEYS:
            ref = float(logs64[k])
            got = float(logs[k])
            err = abs(got - ref) / max(abs(ref), 1e-12) if ref != 0 else abs(got)
            worst[f"log:{k}"] = max(worst.get(f"log:{k}", 0), err)
            assert err <= TOL_LOG * (1 + 2 * step), f"step {step} {k}: gpu {got} oracle {ref} rel {err}"
        if check_grads and step == 0:
            for name, tree, ens in (("actor", agent.actor.grads, False), ("critic", agent.critic.grads, True)):
                for leaf, e in SU.compare_trees(grads64[name], tree, ens).items():
                    worst[f"grad_exact:{name}/{leaf}"] = e
                    assert e <= TOL_GRAD_EXACT, f"grad vs exact oracle {name}/{leaf}: rel {e}"
                for leaf, e in SU.compare_trees(t_grads[name], tree, ens).items():
                    worst[f"grad_tf32:{name}/{leaf}"] = e
                    assert e <= TOL_GRAD, f"grad vs tf32-operand oracle {name}/{leaf}: rel {e}"
        for name, new_t, old_t, tree, ens in (("actor", st64.actor, old.actor, agent.actor.params, False),
                                              ("critic", st64.critic, old.critic, agent.critic.params, True),
                                              ("target", st64.critic_target, old.critic_target, agent.critic.target_params, True)):
            # north-star: updated parameters within 1e-3 of the exact oracle, per network ...
            e = float((agent_flat(tree, new_t, ens) - flat(new_t)).norm() / flat(new_t).norm())
            assert e <= TOL_PARAM, f"step {step} {name} parameters: rel {e}"
            # ... and per leaf wherever the Adam step (~lr per element) is small against the leaf itself; zero-
            # initialised biases and the U(+-1e-3) / U(+-3e-3) heads are dominated by the step (see TOL_DELTA_EXACT)
            old_leaves = dict((n, o) for n, o, _ in SU._pairs(old_t, SU._net(tree, ens)))
            for leaf, e in SU.compare_trees(new_t, tree, ens).items():
                if step == 0 and rms(old_leaves[leaf]) > 50 * cfg.lr:
                    assert e <= TOL_PARAM, f"step {step} param {name}/{leaf}: rel {e}"
            if step == 0 and name != "target":
                for leaf, e in SU.compare_deltas(old_t, new_t, tree, ens).items():
                    worst[f"delta_exact:{name}/{leaf}"] = e
                    assert e <= TOL_DELTA_EXACT, f"update vs exact oracle {name}/{leaf}: rel {e}"
        if step == 0:
            for name, new_t, old_t, tree, ens in (("actor", t_new.actor, old.actor, agent.actor.params, False),
                                                  ("critic", t_new.critic, old.critic, agent.critic.params, True)):
                for leaf, e in SU.compare_deltas(old_t, new_t, tree, ens).items():
                    worst[f"delta:{name}/{leaf}"] = e
                    assert e <= TOL_DELTA, f"update {name}/{leaf}: rel {e}"
        la = agent.alpha.params["params"]["log_alpha"]
        assert SU.rel(la - old.log_alpha.cuda().float(), st64.log_alpha - old.log_alpha) <= TOL_DELTA
    return worst


# ---------------------------------------------------------------------------------------------
# precision="fp32x3": every trunk contraction runs on (hi, lo) tf32 operand pairs (three tensor-core passes), which is
# the arithmetic of the reference's fp32 CPU path.  Here the north-star tolerance is asserted on EVERY leaf of every
# network -- zero-initialised biases and the U(+-1e-3) / U(+-3e-3) heads included, whose value after step 1 IS the Adam
# step -- with no magnitude filter, against the exact fp64 oracle.
# ---------------------------------------------------------------------------------------------
TOL_X3 = 1e-3


def run_case_x3(cfg, per_task, seed=1, steps=1, shuffle=False, counts=None, tol=TOL_X3):
    st = O.init_state(cfg, seed=seed, dtype=torch.float32)
    agent = SU.make_agent(cfg, per_task, seed=seed, precision="fp32x3")
    assert agent.precision == "fp32x3"
    SU.load_oracle_state(agent, st)
    st64 = st.to(torch.float64)
    worst = {}
    for step in range(steps):
        batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=100 + step, dtype=torch.float32)
        if counts is not None:
            task = batch[0][:, -cfg.num_tasks:].argmax(1)
            keep = torch.zeros(task.shape[0], dtype=torch.bool)
            for t, n in enumerate(counts[step] if isinstance(counts[0], (list, tuple)) else counts):
                keep[(task == t).nonzero().flatten()[:n]] = True
            batch, ec, ea = tuple(b[keep] for b in batch), ec[keep], ea[keep]
        if shuffle:
            perm = torch.randperm(batch[0].shape[0], generator=torch.Generator().manual_seed(5))
            batch, ec, ea = tuple(b[perm] for b in batch), ec[perm], ea[perm]
        old = st64
        st64, logs64, grads64, _ = O.mtsac_update(st64, tuple(b.double() for b in batch), ec.double(), ea.double(), cfg,
                                                  return_grads=True)
        _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
        for k in O.LOG_KEYS:
            ref, got = float(logs64[k]), float(logs[k])
            err = abs(got - ref) / max(abs(ref), 1e-12) if ref != 0 else abs(got)
            worst[f"log:{k}"] = max(worst.get(f"log:{k}", 0.0), err)
            assert err <= tol, f"step {step} {k}: gpu {got} oracle {ref} rel {err}"
        if step == 0:
            for name, tree, ens in (("actor", agent.actor.grads, False), ("critic", agent.critic.grads, True)):
                for leaf, e in SU.compare_trees(grads64[name], tree, ens).items():
                    worst[f"grad:{name}/{leaf}"] = e
                    assert e <= tol, f"grad {name}/{leaf}: rel {e}"
        for name, new_t, tree, ens in (("actor", st64.actor, agent.actor.params, False),
                                       ("critic", st64.critic, agent.critic.params, True),
                                       ("target", st64.critic_target, agent.critic.target_params, True)):
            for leaf, e in SU.compare_trees(new_t, tree, ens).items():     # EVERY leaf, no magnitude filter
                worst[f"param:{name}/{leaf}"] = max(worst.get(f"param:{name}/{leaf}", 0.0), e)
                assert e <= tol, f"step {step} param {name}/{leaf}: rel {e}"
        for name, new_t, old_t, tree, ens in (("actor", st64.actor, old.actor, agent.actor.params, False),
                                              ("critic", st64.critic, old.critic, agent.critic.params, True)):
            if step == 0:   # the Adam step itself (new - old), leaf by leaf
                for leaf, e in SU.compare_deltas(old_t, new_t, tree, ens).items():
                    worst[f"delta:{name}/{leaf}"] = e
                    assert e <= 2 * tol, f"update {name}/{leaf}: rel {e}"
        la = agent.alpha.params["params"]["log_alpha"]
        assert SU.rel(la, st64.log_alpha) <= tol and SU.rel(la - old.log_alpha.cuda().float(), st64.log_alpha - old.log_alpha) <= tol
    return worst


def test_fp32x3_small_t10_w64_every_leaf(cuda):
    cfg = O.OracleConfig(num_tasks=10, obs_dim=39 + 10, action_dim=4, width=64)
    print({k: f"{v:.1e}" for k, v in run_case_x3(cfg, per_task=8).items()})


def test_fp32x3_mt10_w400_reference_config_every_leaf(cuda):
    """BASELINE configs[0] (MT10, width 400, B = 1280) in the parity precision: 1e-3 on every leaf."""
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=400, clip=True)
    print({k: f"{v:.1e}" for k, v in run_case_x3(cfg, per_task=128).items()})


def test_fp32x3_shuffled_uneven_rows_every_leaf(cuda):
    cfg = O.OracleConfig(num_tasks=5, obs_dim=20 + 5, action_dim=3, width=96)
    run_case_x3(cfg, per_task=40, shuffle=True, counts=[40, 1, 17, 33, 8])


def test_fp32x3_spare_tiles_and_shrinking_batches_every_leaf(cuda):
    """max_rows leaves room for uneven batches, so some 128-row tiles belong to no task, and which ones changes from update
    to update (task 1 needs two tiles in the first update, one afterwards).  Their dZ rows and bias-gradient partials must
    not leak stale values into dW / the bias gradients: three updates, every leaf (biases included) to 1e-3."""
    cfg = O.OracleConfig(num_tasks=5, obs_dim=20 + 5, action_dim=3, width=96)
    run_case_x3(cfg, per_task=130, steps=3, shuffle=True, counts=[[40, 130, 17, 33, 8], [40, 1, 17, 33, 8], [3, 2, 1, 60, 8]])


def test_fp32x3_task_weights_clip_depth2_one_critic_every_leaf(cuda):
    cfg = O.OracleConfig(num_tasks=4, obs_dim=12 + 4, action_dim=2, width=128, depth=2, num_critics=1,
                         use_task_weights=True, clip=True, initial_temperature=0.7)
    run_case_x3(cfg, per_task=16)


def test_fp32x3_mt10_w1024_every_leaf(cuda):
    """BASELINE configs[1] shape (MT10, width 1024) at a reduced batch the CPU oracle finishes quickly."""
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=1024)
    run_case_x3(cfg, per_task=32)


@pytest.mark.parametrize("precision", ["fp32x3", "tf32"])
def test_drift_50_updates_mt10_w400(cuda, precision):
    """50 consecutive updates at MT10 / width 400 / B = 1280 against the fp64 oracle fed the same batches and noise.

    The update is a chaotic map at the level of single elements (Adam's first steps are ~lr * sign(g), so any element
    whose tiny gradient changes sign moves by 2 lr, and ReLU gates flip): the reference's OWN arithmetic -- the fp32
    oracle on the CPU -- is 1e-6 from fp64 after one update, 1e-4 after 10 and 3e-2 on the actor biases after 50.  So the
    yardstick for step 50 is that fp32 run, not a constant:
      fp32x3: every leaf within 1e-3 for the first 10 updates; after 50, every leaf within max(1e-3, 4 x the fp32 oracle's
              own distance from fp64 for that leaf), log scalars within 1e-3;
      tf32:   every network within 2e-2 (relative l2) after 50, log scalars within 2e-2."""
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=400, clip=True)
    st = O.init_state(cfg, seed=3, dtype=torch.float32)
    agent = SU.make_agent(cfg, 128, seed=3, precision=precision)
    SU.load_oracle_state(agent, st)
    st64, st32 = st.to(torch.float64), st
    nets = lambda s: (("actor", s.actor, agent.actor.params, False), ("critic", s.critic, agent.critic.params, True),  # noqa: E731
                      ("target", s.critic_target, agent.critic.target_params, True))
    for step in range(50):
        batch, ec, ea = O.synthetic_batch(cfg, 128, seed=500 + step, dtype=torch.float32)
        st64, logs64 = O.mtsac_update(st64, tuple(b.double() for b in batch), ec.double(), ea.double(), cfg)
        if precision == "fp32x3":
            st32, _ = O.mtsac_update(st32, batch, ec, ea, cfg)
        _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda())
        if precision == "fp32x3" and step < 10:
            for name, new_t, tree, ens in nets(st64):
                for leaf, le in SU.compare_trees(new_t, tree, ens).items():
                    assert le <= 1e-3, f"update {step}: {name}/{leaf} rel {le}"
    agent._check_status()
    errs = {}
    fp32_nets = {"actor": st32.actor, "critic": st32.critic, "target": st32.critic_target}
    for name, new_t, tree, ens in nets(st64):
        e = float((agent_flat(tree, new_t, ens) - flat(new_t)).norm() / flat(new_t).norm())
        errs[name] = e
        if precision == "tf32":
            assert e <= 2e-2, f"tf32: {name} parameters after 50 updates: rel {e}"
            continue
        yard = {n: SU.rel(a, o) for (n, o, _), (_, a, _) in zip(SU._pairs(new_t, SU._net(tree, ens)), SU._pairs(fp32_nets[name], SU._net(tree, ens)))}
        for leaf, le in SU.compare_trees(new_t, tree, ens).items():
            errs[f"{name}/{leaf}"] = le
            bound = max(1e-3, 4 * yard[leaf])
            assert le <= bound, f"fp32x3: {name}/{leaf} after 50 updates: rel {le}, fp32 oracle's own distance {yard[leaf]}"
    tol = 1e-3 if precision == "fp32x3" else 2e-2
    for k in O.LOG_KEYS:
        ref, got = float(logs64[k]), float(logs[k])
        err = abs(got - ref) / max(abs(ref), 1e-12) if ref != 0 else abs(got)
        errs[f"log:{k}"] = err
        assert err <= tol, f"{precision}: {k} after 50 updates: gpu {got} oracle {ref}"
    assert SU.rel(agent.alpha.params["params"]["log_alpha"], st64.log_alpha) <= tol
    print(precision, {k: f"{v:.1e}" for k, v in errs.items()})


def test_small_t10_w64(cuda):
    cfg = O.OracleConfig(num_tasks=10, obs_dim=39 + 10, action_dim=4, width=64)
    run_case(cfg, per_task=8)


def test_mt10_w400_reference_config(cuda):
    """BASELINE configs[0]: MT10, width 400, batch 128/task (B = 1280)."""
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=400, clip=True)
    worst = run_case(cfg, per_task=128)
    print({k: f"{v:.2e}" for k, v in worst.items()})


def test_shuffled_and_uneven_rows(cuda):
    cfg = O.OracleConfig(num_tasks=5, obs_dim=20 + 5, action_dim=3, width=96)
    run_case(cfg, per_task=40, shuffle=True, counts=[40, 1, 17, 33, 8])


def test_task_weights_clip_depth2_one_critic(cuda):
    cfg = O.OracleConfig(num_tasks=4, obs_dim=12 + 4, action_dim=2, width=128, depth=2, num_critics=1,
                         use_task_weights=True, clip=True, initial_temperature=0.7)
    run_case(cfg, per_task=16)


def test_three_consecutive_updates(cuda):
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=256)
    run_case(cfg, per_task=32, steps=3, check_grads=False)


@pytest.mark.parametrize("precision", ["tf32", "fp32x3"])
def test_fused_layer_launches_equal_layerwise(cuda, monkeypatch, precision):
    """MTRL_FUSE_LAYERS=1 runs the layers of each trunk pass as ONE phased GEMM launch (mtrl_gemm_problem_t::phase); same
    update as one launch per layer: fewer kernels, identical forward values (logs), parameters equal up to the split-K
    atomic order of the dW reductions."""
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=400)
    agents = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MTRL_FUSE_LAYERS", mode)
        st = O.init_state(cfg, seed=3, dtype=torch.float32)
        agents[mode] = SU.make_agent(cfg, 64, seed=3, precision=precision)
        SU.load_oracle_state(agents[mode], st)
    logs = {}
    for step in range(3):
        batch, ec, ea = O.synthetic_batch(cfg, 64, seed=200 + step, dtype=torch.float32)
        for mode, agent in agents.items():
            _, logs[mode] = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
        for k in O.LOG_KEYS:
            a, b = float(logs["0"][k]), float(logs["1"][k])
            assert abs(a - b) <= 2e-5 * max(abs(a), 1e-6) * (1 + step), (step, k, a, b)
    # 3 layers x 6 passes -> 6 launches
    assert agents["0"].launches_per_update() - agents["1"].launches_per_update() == 12
    for name in ("actor", "critic"):
        p0 = torch.cat([x.flatten() for x in O.tree_leaves(getattr(agents["0"], name).params)])
        p1 = torch.cat([x.flatten() for x in O.tree_leaves(getattr(agents["1"], name).params)])
        assert float((p0 - p1).norm() / p0.norm()) < 1e-5


def test_mt50_w2048_full_size_vs_oracle_on_gpu(cuda):
    """BASELINE headline config (MT50, width 2048, B = 6400): the oracle itself is run in fp64 with
    torch on the same GPU (it is device agnostic); the fused path must match it to the north-star
    tolerance at full size."""
    cfg = O.OracleConfig(num_tasks=50, obs_dim=89, action_dim=4, width=2048)
    st = O.init_state(cfg, seed=1, dtype=torch.float32)
    agent = SU.make_agent(cfg, 128, seed=1)
    SU.load_oracle_state(agent, st)
    batch, ec, ea = O.synthetic_batch(cfg, 128, seed=7, dtype=torch.float32)
    dev = torch.device("cuda")
    mv = lambda t: O.tree_map(lambda x: x.to(dev).double(), t)  # noqa: E731
    st64 = O.OracleState(mv(st.actor), mv(st.critic), mv(st.critic_target), st.log_alpha.to(dev).double(),
                         {k: {"m": mv(v["m"]), "v": mv(v["v"]), "count": v["count"]} for k, v in st.opt.items()})
    b64 = tuple(b.to(dev).double() for b in batch)
    import dataclasses

    new64, logs64, grads64, _ = O.mtsac_update(st64, b64, ec.to(dev).double(), ea.to(dev).double(), cfg, return_grads=True)
    _, _, tgrads, _ = O.mtsac_update(st64, b64, ec.to(dev).double(), ea.to(dev).double(),
                                     dataclasses.replace(cfg, matmul_operands="tf32"), return_grads=True)
    _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    for k in O.LOG_KEYS:
        ref, got = float(logs64[k]), float(logs[k])
        assert abs(got - ref) <= TOL_LOG * max(abs(ref), 1e-12) or ref == 0 == got, f"{k}: {got} vs {ref}"
    for name, tree, ens in (("critic", agent.critic.grads, True), ("actor", agent.actor.grads, False)):
        for leaf, e in SU.compare_trees(grads64[name], tree, ens).items():
            assert e <= TOL_GRAD_EXACT, f"{name} grad vs exact oracle {leaf}: {e}"
        for leaf, e in SU.compare_trees(tgrads[name], tree, ens).items():
            assert e <= TOL_GRAD, f"{name} grad vs tf32-operand oracle {leaf}: {e}"
    for name, new_t, tree, ens in (("critic", new64.critic, agent.critic.params, True), ("actor", new64.actor, agent.actor.params, False)):
        fa = torch.cat([a.detach().double().flatten() for _, _, a in SU._pairs(new_t, SU._net(tree, ens))])
        fo = torch.cat([o.flatten() for _, o, _ in SU._pairs(new_t, SU._net(tree, ens))])
        assert float((fa - fo).norm() / fo.norm()) <= TOL_PARAM, name
        for leaf, e in SU.compare_trees(new_t, tree, ens).items():
            if leaf.startswith("layer_") and leaf.endswith("kernel"):  # other leaves are dominated by the step itself
                assert e <= TOL_PARAM, f"{name} param {leaf}: {e}"


def test_row_order_invariance_full_size(cuda):
    """Size-independent property: the update is a sum over rows, so permuting the batch rows changes
    results only by fp32 summation order."""
    cfg = O.OracleConfig(num_tasks=50, obs_dim=89, action_dim=4, width=1024)
    st = O.init_state(cfg, seed=2, dtype=torch.float32)
    batch, ec, ea = O.synthetic_batch(cfg, 128, seed=9, dtype=torch.float32)
    outs = []
    for perm_seed in (None, 3):
        agent = SU.make_agent(cfg, 128, seed=2)
        SU.load_oracle_state(agent, st)
        b, c, a = batch, ec, ea
        if perm_seed is not None:
            perm = torch.randperm(b[0].shape[0], generator=torch.Generator().manual_seed(perm_seed))
            b, c, a = tuple(x[perm] for x in b), c[perm], a[perm]
        _, logs = agent.update(tuple(x.cuda() for x in b), eps_c=c.cuda(), eps_a=a.cuda(), check=True)
        outs.append((torch.stack([logs[k] for k in O.LOG_KEYS]).cpu(), agent._flat["critic_params"].clone(),
                     agent._flat["actor_grads"].clone()))
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-4, atol=1e-7)
    assert SU.rel(outs[0][1], outs[1][1]) < 1e-5
    assert SU.rel(outs[0][2], outs[1][2]) < 5e-3  # summation order moves a few ReLU gates too


def test_philox_noise_path_runs_and_differs(cuda):
    cfg = O.OracleConfig(num_tasks=4, obs_dim=16, action_dim=4, width=64)
    st = O.init_state(cfg, seed=3)
    batch, _, _ = O.synthetic_batch(cfg, 16, seed=1)
    agent = SU.make_agent(cfg, 16, seed=3)
    SU.load_oracle_state(agent, st)
    _, l1 = agent.update(tuple(b.cuda() for b in batch), check=True)
    a1 = float(l1["losses/actor_loss"])
    _, l2 = agent.update(tuple(b.cuda() for b in batch), check=True)
    assert torch.isfinite(torch.stack(list(l2.values()))).all()
    assert a1 != float(l2["losses/actor_loss"])
    assert int(agent.actor.step) == 2 and int(agent.critic.step) == 2 and int(agent.alpha.step) == 2


def test_bad_batches_are_reported(cuda):
    cfg = O.OracleConfig(num_tasks=4, obs_dim=16, action_dim=4, width=64)
    agent = SU.make_agent(cfg, 16, seed=3, max_batch=256, max_rows=512)
    batch, ec, ea = O.synthetic_batch(cfg, 64, seed=1)
    # 64 rows per task x 4 fits 512 padded rows; 130 rows of one task would need 256 padded rows for it
    ok = tuple(b.cuda() for b in batch)
    agent.update(ok, eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    obs = ok[0].clone()
    obs[:, -4:] = 0
    obs[:, -4] = 1  # every row claims task 0: 256 rows -> fits (256 padded), still fine
    agent.update((obs,) + ok[1:], eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    with pytest.raises(Exception):
        small = SU.make_agent(cfg, 16, seed=3)  # max_batch 64
        small.update(ok, eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)  # 256 rows > max_batch
