"""GEMM diagnostics on the GPU box (development aid): self tests of the deferred-LayerNorm epilogues, then a per-shape
timing table (operand ring depth, mainloop-only mode) written to gpurun_out/gemm_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402


def main():
    ctx = _native.default_context(0)
    out = {"selftest": [], "bench": []}
    for epi, M, N, K in [(11, 300, 768, 768), (11, 5000, 768, 1152), (12, 300, 2304, 768), (13, 300, 2304, 768),
                         (11, 40000, 768, 768)]:
        try:
            d, m = ctx.selftest_gemm(M, N, K, epi)
            out["selftest"].append({"epi": epi, "M": M, "N": N, "K": K, "diff": d, "ref_max": m})
        except Exception as e:  # noqa: BLE001
            out["selftest"].append({"epi": epi, "M": M, "N": N, "K": K, "error": str(e)})
        print(out["selftest"][-1], flush=True)
    M = int(os.environ.get("PROBE_M", "131072"))
    plan = [("wqkv", 1, 2304, 768, (0, 3)), ("wqkv_norm", 12, 2304, 768, (0, 3, 4)),
            ("wi", 3, 2304, 768, (0,)), ("wi_norm", 13, 2304, 768, (0, 3, 4)),
            ("wo", 2, 768, 768, (0, 3)), ("wo_stats", 11, 768, 768, (0, 4)),
            ("wo2", 2, 768, 1152, (0, 3)), ("wo2_stats", 11, 768, 1152, (0, 4))]
    for name, epi, N, K, modes in plan:
        for dbg in modes:
            try:
                ms = ctx.bench_gemm(M, N, K, epi, stages=0, debug_mode=dbg, iters=10)
                rec = {"name": name, "epi": epi, "N": N, "K": K, "debug": dbg, "ms": round(ms, 4),
                       "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
            except Exception as e:  # noqa: BLE001
                rec = {"name": name, "epi": epi, "debug": dbg, "error": str(e)}
            out["bench"].append(rec)
            print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
