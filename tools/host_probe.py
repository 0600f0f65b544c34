"""Strings-in throughput of B200SpanExtractor.extract_spans_batch on the GPU box for several tokenizer-worker counts
(development aid).  Prints extractions/s; writes gpurun_out/host_probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import CHUNKS_PER_Q, make_text_batch
    from verbatim_rag_b200 import B200SpanExtractor
    from verbatim_rag_b200.synthetic import ModernBertSpec, SyntheticTokenizer, make_modernbert_weights
    spec = ModernBertSpec(layers=22)
    w = make_modernbert_weights(1001, spec)
    tk = SyntheticTokenizer("modernbert")
    qs, rs = make_text_batch(tk, 192, 7)
    out = []
    for workers, step in ((0, 512), (2, 512), (3, 512), (4, 512), (4, 1024), (6, 1024)):
        ext = B200SpanExtractor(weights=w, tokenizer=tk, num_layers=22, vocab_size=spec.vocab_size, max_tokens=131072,
                                tokenizer_workers=workers)
        ext.pipeline_pairs = step
        ext.extract_spans_batch(qs[:40], rs[:40])
        t0 = time.perf_counter()
        ext.extract_spans_batch(qs, rs)
        dt = time.perf_counter() - t0
        rec = {"workers": workers, "slice_pairs": step, "extractions_per_s": round(len(qs) * CHUNKS_PER_Q / dt, 1)}
        out.append(rec)
        print(rec, flush=True)
        ext._workers.close()
        ext._enc.close()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "host_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
