"""CPU: the C-ABI library builds, loads and exports every symbol include/mtrl_b200.h declares; host-side
argument validation works without a GPU (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mtrl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtrl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_groups():
    names = declared_functions()
    for prefix in ("mtrl_gemm_plan_", "mtrl_sampler_", "mtrl_sac_"):
        assert any(n.startswith(prefix) for n in names), prefix
    assert "mtrl_last_error" in names


def test_library_exports_every_declared_symbol():
    from mtrl_b200 import _lib

    l = _lib.lib()
    missing = [n for n in declared_functions() if not hasattr(l, n)]
    assert not missing, missing
    assert l.mtrl_abi_version() >= 1


def test_declarations_registered_after_load_are_applied():
    from mtrl_b200 import _lib

    l = _lib.lib()
    import mtrl_b200.rl.algorithms  # noqa: F401
    import mtrl_b200.rl.buffers  # noqa: F401

    assert l.mtrl_sampler_create.argtypes is not None
    assert l.mtrl_sac_update.argtypes is not None


def test_layout_query_and_validation_run_on_host():
    from mtrl_b200 import _lib
    from mtrl_b200.rl.algorithms.mtsac import SacConfigC, SacLayoutC

    l = _lib.lib()
    cfg = SacConfigC(num_tasks=10, task_begin=0, num_local_tasks=10, obs_dim=49, action_dim=4, width=400, depth=3,
                     num_critics=2, max_rows=1280, max_batch=1280)
    lay = SacLayoutC()
    assert l.mtrl_sac_query_layout(C.byref(cfg), C.byref(lay)) == 0
    # logical parameter counts of SURVEY 8(d): 372 880 actor, 692 820 critic (two members); the flat layout
    # only adds alignment padding and 32 reduction slots
    assert lay.actor.total >= 372_880 and lay.actor.total - 372_880 < 4096
    assert lay.critic.total >= 692_820 and lay.critic.total - 692_820 < 4096
    assert lay.actor.trunk_total % 32 == 0 and lay.k_actor == 64 and lay.k_critic == 64
    bad = SacConfigC(num_tasks=10, task_begin=0, num_local_tasks=10, obs_dim=49, action_dim=4, width=402, depth=3,
                     num_critics=2, max_rows=1280, max_batch=1280)
    assert l.mtrl_sac_query_layout(C.byref(bad), C.byref(lay)) != 0
    assert b"width" in l.mtrl_last_error()


def test_num_params_matches_reference_kat():
    """372 880 is the exact count behind the 370_000 hard-coded in plots/get_data.py:51-53."""
    from mtrl_b200.rl.algorithms.mtsac import MTSAC, SacConfigC

    obj = object.__new__(MTSAC)
    obj._cfg = SacConfigC(obs_dim=49, action_dim=4, width=400, depth=3, num_critics=2)
    obj.num_tasks = 10
    assert obj.get_num_params() == {"actor_num_params": 372_880, "critic_num_params": 692_820}


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mtrl_b200 import _lib
    from mtrl_b200.presets import _Space, metaworld_mtmhsac
    from mtrl_b200.rl.algorithms import MTSAC
    from mtrl_b200.rl.buffers import MultiTaskReplayBuffer

    with pytest.raises(_lib.MtrlError):
        MultiTaskReplayBuffer(100, 10, _Space((49,)), _Space((4,)), seed=0)
    cfg, env = metaworld_mtmhsac(10, 400)
    with pytest.raises(_lib.MtrlError):
        MTSAC.initialize(cfg, env)
