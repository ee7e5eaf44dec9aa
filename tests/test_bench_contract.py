"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm's JSON line and the
refusal of the product arm to run (or fall back to the CPU) when no CUDA device is present."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    env = dict(os.environ)
    env.pop("MCTSB200_LIB", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MCTS simulations/sec" and d["unit"] == "simulations/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    # the headline is BASELINE config 2 on the DEFAULT SimpleGridWorld with the deterministic rollout policy, and the reference arm
    # names exactly the configuration the product arm names (same keys, same values, roots_per_gpu included)
    assert d["config"]["workload"] == "config2-uct-gridworld-default" and d["config"]["n_iterations"] == 1000
    assert d["config"]["mdp_params"] == {"tprob": 0.7} and d["config"]["rollout_policy"] == "up" and d["config"]["roots_per_gpu"] == 65536
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.config_of("config2-uct-gridworld-default", 65536, 1)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "roots" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0
    assert not any(l.startswith("{") for l in r.stdout.splitlines())  # no number without the CUDA path
