"""Strings in -> spans out through B200SpanExtractor with the tokenizer worker pool on (a one-minute GPU check of the
host pipeline: worker processes, device span runs, per-question results).  Usage: python tools/plugin_smoke.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import CHUNKS_PER_Q, make_text_batch
    from verbatim_rag_b200 import B200SpanExtractor
    from verbatim_rag_b200.synthetic import ModernBertSpec, SyntheticTokenizer, make_modernbert_weights
    spec = ModernBertSpec(layers=22)
    w = make_modernbert_weights(1001, spec)
    tk = SyntheticTokenizer("modernbert")
    qs, rs = make_text_batch(tk, 48, 7)
    outs = {}
    for workers in (2, 0):
        ext = B200SpanExtractor(weights=w, tokenizer=tk, num_layers=22, vocab_size=spec.vocab_size, max_tokens=131072,
                                tokenizer_workers=workers)
        ext.extract_spans_batch(qs[:8], rs[:8])
        t0 = time.perf_counter()
        outs[workers] = ext.extract_spans_batch(qs, rs)
        dt = time.perf_counter() - t0
        print("workers %d: %.0f extractions/s, %d spans, pool %s" % (
            workers, len(qs) * CHUNKS_PER_Q / dt, sum(len(v) for d in outs[workers] for v in d.values()),
            "on" if ext._workers._pool is not None else "off"), flush=True)
        ext._workers.close()
        ext._enc.close()
    assert outs[2] == outs[0], "worker pool changes the result"
    print("PLUGIN_OK")


if __name__ == "__main__":
    main()
