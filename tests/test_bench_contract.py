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


def test_secondary_blocks_with_collectives_run_on_every_rank():
    """Multi-rank safety of bench.py: a secondary block that contains a collective (all-reduce of a time, barrier,
    sharded search, peer exchange) must be flagged `collective` in the ONE table both rank branches iterate, so that
    every rank enters it; a block rank 0 alone entered would leave the job hanging in the final barrier (this happened
    once with a block added to the rank-0 list only)."""
    import inspect
    import re

    import bench
    src = inspect.getsource(bench.run_gpu_arm)
    table = re.findall(r'\("([\w]+)",\s*(True|False),\s*lambda:\s*(sec_\w+)\(', src)
    assert len(table) >= 5, table
    assert src.count("secondary_blocks") >= 3          # defined once, iterated by the rank != 0 and the rank 0 branch
    assert "if collective:" in src
    for name, flag, fn_name in table:
        body = inspect.getsource(getattr(bench, fn_name))
        has_collective = bool(re.search(r"\bdist\.|_all_max\(|barrier\(|sharded_search_|make_peer_exchange|rag_query_batch|"
                                        r"ShardedB200VectorStore", body))
        assert has_collective == (flag == "True"), (name, fn_name, flag, has_collective)
