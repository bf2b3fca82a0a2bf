"""bench.py contract checks that need no GPU: byte formulas of SURVEY.md §8(d) and the JSON line of the reference arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
    assert bench.algorithmic_bytes_per_asset_step(50, 50, 5, 0.0, False) == pytest.approx(8.5)          # Mode S, A=50
    assert bench.algorithmic_bytes_per_asset_step(100, 50, 5, 0.0, False) == pytest.approx(8.25)
    assert bench.algorithmic_bytes_per_asset_step(500, 50, 5, 0.0025, False) == pytest.approx(12.05)    # commission: + w_last
    assert bench.algorithmic_bytes_per_asset_step(100, 50, 5, 0.0, True) == pytest.approx(1208.25)      # Mode O
    assert bench.algorithmic_bytes_per_asset_step(50, 50, 5, 0.0, True) == pytest.approx(1208.5)


def test_workloads_are_the_baseline_configs():
    w = bench.WORKLOADS
    assert w["c2"][:3] == (4096, 50, 50) and w["c3"][:3] == (65536, 100, 50) and w["c5"][:3] == (262144, 500, 50)
    assert w["c4_shard"][0] * 8 == 1048576 and w["c4_shard"][1:3] == (100, 50)
    assert w["c5"][4] == 0.0025 and not w["c5"][5]                     # commission, state-only
    assert w["c4"][:3] == (1048576, 100, 50)                           # config 4 whole (divided over the ranks: strong scaling)


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, PMRL_BENCH_REF_SECONDS="0.2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--workload", "c2"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["config"]["workload"] == "c2"
    from oracle import live_reference as live
    want_kind = "reference" if live.locate() else "port"                # the live tree first, else the port
    assert line["cpu_baseline"]["kind"] == want_kind and line["cpu_baseline"]["cores"] >= 1
    cfg = bench.shared_config(type("A", (), dict(workload="c2", envs=0, tune=[], graph=0, burst=0))(), 1)
    assert line["config"] == cfg                                        # both arms print the same config object
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_falls_back_to_the_port_without_the_tree():
    env = dict(os.environ, PMRL_BENCH_REF_SECONDS="0.2", PMRL_REFERENCE_ROOT="/nonexistent")
    code = ("import sys; sys.path.insert(0, %r); from oracle import live_reference as live; live.locate = lambda: None; import bench; "
            "print(bench.cpu_baseline(11, 8, 0.0, seconds=0.2, context=False)['kind'])" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1] == "port", out.stderr[-2000:]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""
