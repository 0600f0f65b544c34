"""bench.py contract on a CPU-only box: the reference arm prints exactly one JSON line with the contract keys, and the
GPU arm refuses to run without a CUDA device (no CPU fallback behind the headline number)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_cuda

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-sample", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"].startswith("ModernBERT-v2 span extraction")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # both arms print the same `config` (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config()
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"].startswith(d["metric"])


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_gpu_arm_fails_loudly_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr
