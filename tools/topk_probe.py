"""Dense top-10 over 1 M x 768 on the GPU box (development aid): the batched GEMM search vs the 16-query scan, per-class
CUDA-event times.  Writes gpurun_out/topk_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    ctx = _native.default_context(0)
    dim, k = 768, 10
    g = torch.Generator(device="cuda").manual_seed(1004)
    ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    ix.add_dense(torch.randn(n, dim, device="cuda", generator=g))
    st = torch.cuda.ExternalStream(ctx.stream)
    out = []
    for nq in (16, 64, 256, 1000, 4000):
        q = torch.randn(nq, dim, device="cuda", generator=g)
        ids = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        sc = torch.empty(nq, k, dtype=torch.float32, device="cuda")
        for env in (None, "0"):
            if env is None:
                os.environ.pop("VRAG_SCAN_BIG_MIN", None)
            else:
                os.environ["VRAG_SCAN_BIG_MIN"] = env
            if nq > 1000 and env == "0":
                continue
            for _ in range(2):
                ix.search_dense_device(q, nq, k, ids, sc)
            ctx.sync()
            ctx.profile(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(3):
                ix.search_dense_device(q, nq, k, ids, sc)
            e1.record(st)
            ctx.sync()
            pr = ctx.profile_read()
            ctx.profile(False)
            ms = e0.elapsed_time(e1) / 3
            rec = {"n": n, "nq": nq, "path": "gemm" if env is None and nq >= 64 else "scan", "ms": round(ms, 3),
                   "queries_per_s": round(nq / ms * 1e3), "scan_ms": round(pr["scan"]["ms"] / 3, 3),
                   "select_ms": round(pr["select"]["ms"] / 3, 3), "finish_ms": round(pr["other"]["ms"] / 3, 3), "tflops_algorithmic": round(2 * nq * n * dim / ms / 1e9, 1)}
            out.append(rec)
            print(rec, flush=True)
    os.environ.pop("VRAG_SCAN_BIG_MIN", None)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "topk_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
