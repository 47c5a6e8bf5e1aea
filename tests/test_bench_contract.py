"""bench.py contract that can be checked without a GPU: the reference arm prints ONE JSON line with the keys the driver
reads (impl, metric, value, unit, config.workload, cpu_baseline{kind,cores,sample,value}, e2e{...}), and the B200 arm
refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def _run(*args):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=900, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encoded audio-sec/sec" and d["unit"] == "audio-s/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("c1") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run("--workload", "c1", "--steps", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
