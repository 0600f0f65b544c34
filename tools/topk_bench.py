"""Dense top-k timings on the 1 M x 768 corpus (BASELINE configs[3] shape) + a cross-check of the tensor-core scan
(> 8 queries per call) against the FMA scan (<= 8 queries per call).  Development aid; bench.py reports the numbers."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402


def main():
    ctx = _native.default_context(0)
    n, dim, k = int(os.environ.get("TOPK_N", "1000000")), 768, 10
    g = torch.Generator(device="cuda").manual_seed(0)
    corpus = torch.randn(n, dim, device="cuda", generator=g)
    ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    ix.add_dense(corpus)
    del corpus
    out = {"n": n, "dim": dim}
    q = torch.randn(1000, dim, device="cuda", generator=g)
    st = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))

    def search(qq):
        nq = qq.shape[0]
        ids = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        sc = torch.empty(nq, k, dtype=torch.float32, device="cuda")
        ix.search_dense_device(qq, nq, k, ids, sc)
        ctx.sync()
        return ids, sc

    # cross-check: 64 queries in one call (tensor-core scan) vs 16 calls of 4 (FMA scan)
    ids_tc, sc_tc = search(q[:64].contiguous())
    ids_fma = torch.cat([search(q[i:i + 4].contiguous())[0] for i in range(0, 64, 4)])
    sc_fma = torch.cat([search(q[i:i + 4].contiguous())[1] for i in range(0, 64, 4)])
    out["tc_vs_fma_ids_equal"] = bool((ids_tc == ids_fma).all().item())
    out["tc_vs_fma_score_maxdiff"] = float((sc_tc - sc_fma).abs().max().item())
    for nq in (1, 2, 4, 8, 16, 64, 256, 1000):
        qq = q[:nq].contiguous()
        search(qq)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ids = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        sc = torch.empty(nq, k, dtype=torch.float32, device="cuda")
        e0.record(st)
        for _ in range(3):
            ix.search_dense_device(qq, nq, k, ids, sc)
        e1.record(st)
        ctx.sync()
        wall = e0.elapsed_time(e1) / 3
        ctx.profile(True)
        for _ in range(3):
            search(qq)
        prof = ctx.profile_read()
        ctx.profile(False)
        scan_ms = prof["scan"]["ms"] / 3
        passes = prof["scan"]["launches"] / 3
        tot = sum(v["ms"] for v in prof.values()) / 3
        out[f"q{nq}"] = {"scan_ms": scan_ms, "passes": passes, "scan_GBps": passes * n * dim * 4 / scan_ms / 1e6,
                         "device_ms_sum_of_kernels": tot, "search_ms": wall, "qps": nq / wall * 1e3,
                         "search_GBps": passes * n * dim * 4 / wall / 1e6}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
