"""Precise-mode timing on the GPU box (development aid): split-precision GEMM shapes, split attention, and a short
whole-model step with the per-class CUDA-event profile; fast mode beside it.  Writes gpurun_out/precise_probe.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402
from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights  # noqa: E402


def main():
    import torch
    ctx = _native.default_context(0)
    out = {"gemm": [], "attention": [], "model": {}}
    M = 131072
    for name, N, K, epi, fast_epi in (("Wqkv+RoPE", 2304, 768, 1, 12), ("Wi+GeGLU", 2304, 768, 3, 13),
                                      ("Wo resid", 768, 768, 2, 11), ("mlp.Wo resid", 768, 1152, 2, 11)):
        ms_s = ctx.bench_gemm_split(M, N, K, epi, iters=10)
        ms_f = ctx.bench_gemm(M, N, K, fast_epi, iters=10)
        rec = {"gemm": name, "split_ms": round(ms_s, 4), "split_tflops_3x": round(3 * 2 * M * N * K / ms_s / 1e9, 1),
               "fast_ms": round(ms_f, 4), "ratio": round(ms_s / ms_f, 2)}
        out["gemm"].append(rec)
        print(rec, flush=True)
    for window in (-1, 64):
        ms_s = ctx.bench_attention_split(256, 512, window, iters=10)
        ms_f = ctx.bench_attention(256, 512, window, iters=10)
        rec = {"window": window, "split_ms": round(ms_s, 4), "fast_ms": round(ms_f, 4), "ratio": round(ms_s / ms_f, 2)}
        out["attention"].append(rec)
        print(rec, flush=True)
    spec = ModernBertSpec(layers=22)
    w = make_modernbert_weights(1001, spec)
    nseq, L = 1024, 512
    rng = np.random.default_rng(3)
    ids = rng.integers(5, 50279, size=(nseq, L), dtype=np.int32)
    ids[:, 0], ids[:, 30], ids[:, -1] = spec.cls_id, spec.sep_id, spec.sep_id
    cu = (np.arange(nseq + 1) * L).astype(np.int32)
    ids_d = torch.from_numpy(ids.reshape(-1)).cuda()
    probs_d = torch.empty(nseq * L, dtype=torch.float32, device="cuda")
    st = torch.cuda.ExternalStream(ctx.stream)
    for prec in ("fast", "precise"):
        enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=131072,
                              precision=prec)
        for _ in range(2):
            enc.span_forward_device(ids_d, cu, probs_d)
        ctx.sync()
        ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            enc.span_forward_device(ids_d, cu, probs_d)
        e1.record(st)
        ctx.sync()
        ms = e0.elapsed_time(e1) / 3
        pr = ctx.profile_read()
        ctx.profile(False)
        rec = {"ms_per_1024_seqs": round(ms, 2), "extractions_per_s": round(nseq / ms * 1e3, 1),
               "class_ms": {k: round(v["ms"] / 3, 2) for k, v in pr.items() if v["launches"]}}
        out["model"][prec] = rec
        print(prec, rec, flush=True)
        enc.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precise_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
