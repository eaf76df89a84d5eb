"""bench.py contract (CPU): the reference arm prints ONE JSON line with the fields the driver reads; the B200 arm's
helpers compute the algorithmic FLOPs of SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, MTRL_REF_BUDGET_S="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "mt10_w400",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MT-SAC gradient updates/sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["config"]["workload"] == "mt10_w400"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # both arms print the SAME config object: it is a function of the command line only
    import argparse

    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.workload_config(argparse.Namespace(workload="mt10_w400", gpus=1, capacity=100000))
    assert d["steps"] == 2 and d["warmup"] == 3   # (warm-up is raised to the timing rules' minimum of 3 on both arms)


def test_non_zero_rank_of_reference_arm_prints_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", MTRL_REF_BUDGET_S="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "mt10_w400"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_flops_formula():
    sys.path.insert(0, ROOT)
    import bench

    # MT50 / W2048 / B6400 -> 1757.7 GFLOP with the one-hot columns counted as dense K; the official
    # formula (d = 39 / 43, heads included) is <= 2.5 % below
    f = bench.algorithmic_flops(50, 2048, 6400)
    assert 0.975 * 1757.7e9 <= f <= 1757.7e9
    trunk = bench.algorithmic_flops(50, 2048, 6400, trunk_only=True)
    assert trunk < f and (f - trunk) / f < 0.01
    assert set(bench.WORKLOADS) >= {"mt10_w400", "mt10_w1024", "mt50_w2048", "mt50_w4096"}
