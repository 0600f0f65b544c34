"""Quick device-side timings (CUDA events on the library's stream) written to gpurun_out/quickbench.json.
Development aid; bench.py is the contract benchmark."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402
from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights  # noqa: E402


def timed(ctx, fn, iters=3, warmup=1):
    st = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warmup):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    ctx.sync()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    ctx = _native.default_context(0)
    layers = int(os.environ.get("QB_LAYERS", "22"))
    nseq = int(os.environ.get("QB_NSEQ", "256"))
    L = 512
    spec = ModernBertSpec(layers=layers)
    t0 = time.time()
    w = make_modernbert_weights(1001, spec)
    out["weights_s"] = time.time() - t0
    for max_tokens in (32768, 65536, 131072):
        if max_tokens > nseq * L:
            continue
        enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=max_tokens)
        rng = np.random.default_rng(0)
        ids = torch.from_numpy(rng.integers(5, 50279, size=nseq * L).astype(np.int32)).cuda()
        cu = np.arange(nseq + 1, dtype=np.int32) * L
        probs = torch.empty(nseq * L, dtype=torch.float32, device="cuda")
        l0 = ctx.launches
        ms = timed(ctx, lambda: enc.span_forward_device(ids, cu, probs), iters=2, warmup=1)
        flops = 122.65e9 * nseq * layers / 22.0
        out[f"span_forward_mt{max_tokens}"] = {"ms": ms, "seq_per_s": nseq / ms * 1e3, "tflops": flops / ms / 1e9,
                                               "launches_per_call": (ctx.launches - l0) / 3}
        enc.close()
        del enc
    # GEMM-only sanity: selftest shapes time (includes SIMT reference, so not a perf number) -- skip.
    # dense scan
    n, dim = 1_000_000, 768
    g = torch.Generator(device="cuda").manual_seed(0)
    corpus = torch.randn(n, dim, device="cuda", generator=g)
    ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    ix.add_dense(corpus)
    del corpus
    for nq in (1, 8, 64):
        q = torch.randn(nq, dim, device="cuda", generator=g)
        ids_o = torch.empty(nq, 10, dtype=torch.int64, device="cuda")
        s_o = torch.empty(nq, 10, dtype=torch.float32, device="cuda")
        ms = timed(ctx, lambda: ix.search_dense_device(q, nq, 10, ids_o, s_o), iters=3, warmup=1)
        passes = (nq + 7) // 8
        out[f"dense_search_q{nq}"] = {"ms": ms, "GBps_corpus": passes * n * dim * 4 / ms / 1e6, "qps": nq / ms * 1e3}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quickbench.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
