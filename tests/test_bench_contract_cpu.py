"""CPU checks of bench.py's contract with the driver that need no GPU: the reference arm prints ONE JSON line with
the agreed keys (and nothing else on stdout), other ranks of a torchrun launch stay silent, and the GPU arm fails
loudly -- not with a CPU fallback -- when there is no device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-size", "24"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_are_silent():
    out = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-size", "24", "--gpus", "2"],
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    out = _run(["--steps", "1", "--warmup", "3", "--no-extras", "--no-e2e"])
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")], "no bench line may be printed without a GPU"
