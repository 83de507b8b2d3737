"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys (one
process per unit at --gpus N, non-zero ranks print nothing), and the B200 arm refuses to run without a GPU."""
import json
import os
import subprocess
import sys

import pytest

import oracle.ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(argv), capture_output=True, text=True, env=e, timeout=600)


@pytest.mark.parametrize("gpus", [1, 2])
def test_reference_arm_line(gpus):
    if not R.available():
        pytest.skip("oracle/_ref not built")
    out = run("--impl", "reference", "--workload", "tiny", "--gpus", str(gpus), "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr
    lines = [x for x in out.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["n_gpus"] == gpus and d["steps"] == 2
    assert d["unit"] == "bases/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == gpus and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config keys as the B200 arm prints (the driver compares the two objects key by key): unit 0 of the workload
    assert set(d["config"]) == {"workload", "bases_per_step_per_gpu", "mums_per_step", "minl", "minn", "units", "l2", "sharding"}
    assert d["config"]["units"] == gpus and 390000 < d["config"]["bases_per_step_per_gpu"] < 410000


def test_reference_arm_other_ranks_are_silent():
    out = run("--impl", "reference", "--workload", "tiny", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_bounded_sample():
    if not R.available():
        pytest.skip("oracle/_ref not built")
    out = run("--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "1", env={"RV_REF_BUDGET_S": "0.08"})
    d = json.loads(out.stdout.strip())
    # the config still names the full workload; the sample that was timed is described in cpu_baseline.sample
    assert "of every genome per step" in d["cpu_baseline"]["sample"] and 390000 < d["config"]["bases_per_step_per_gpu"] < 410000


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = run("--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CPU fallback" in out.stderr
