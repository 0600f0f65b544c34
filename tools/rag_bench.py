"""BASELINE configs[4] shape, bounded: SPLADE retrieve top-20 + span extraction through the plugin surfaces
(B200SpladeProvider -> B200VectorStore -> B200SpanExtractor, batched entry points of verbatim_rag_b200.pipeline) on
synthetic texts, with the stage breakdown SURVEY.md 8(d) asks for.  Strings in, verbatim span strings out: host
tokenisation (tokenizers library) is inside every stage.  bench.py embeds the result as secondary["rag_e2e"].

    python tools/rag_bench.py [n_chunks] [n_queries]
"""
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n_chunks: int = 4096, n_queries: int = 128, k: int = 20, chunk_tokens: int = 256, device: str = "cuda:0"):
    from verbatim_rag_b200 import B200SpanExtractor, B200SpladeProvider, B200VectorStore
    from verbatim_rag_b200.pipeline import index_query_batch

    t0 = time.perf_counter()
    prov = B200SpladeProvider("synthetic:1002", device=device)                  # BERT-base MLM, 12 layers
    ext = B200SpanExtractor("synthetic:1001", device=device, max_tokens=131072)  # ModernBERT-base token classifier
    btok = prov._te.tokenizer
    load_s = time.perf_counter() - t0
    rng = np.random.default_rng(1005)
    chunks = [btok.make_text(rng, chunk_tokens) for _ in range(n_chunks)]
    questions = [btok.make_question(rng, int(rng.integers(12, 21))) for _ in range(n_queries)]
    store = B200VectorStore(enable_dense=False, enable_sparse=True, device=device)
    ids = [f"c{i:07d}" for i in range(n_chunks)]

    t0 = time.perf_counter()
    csr = prov.embed_batch_csr(chunks)
    t1 = time.perf_counter()
    store.add_csr(ids, *csr, chunks, chunks, [{} for _ in chunks])
    t2 = time.perf_counter()
    index = types.SimpleNamespace(vector_store=store, sparse_provider=prov, dense_provider=None)
    index_query_batch(index, questions[:4], k=k)          # warm-up
    ext.extract_spans_batch(questions[:2], index_query_batch(index, questions[:2], k=k))
    first = None
    for _ in range(2):   # the first full-size batch grows pinned / device staging buffers; report the steady state
        t3 = time.perf_counter()
        found = index_query_batch(index, questions, k=k)      # SPLADE query encode + sparse-dot top-k
        t4 = time.perf_counter()
        spans = ext.extract_spans_batch(questions, found)     # tokenise pairs + 22-layer forward + span post-processing
        t5 = time.perf_counter()
        if first is None:
            first = (t5 - t3) * 1e3
    n_ext = sum(len(r) for r in found)
    return {
        "chunks": n_chunks, "chunk_tokens": chunk_tokens, "queries": n_queries, "k": k, "extractions": n_ext,
        "model_load_s": load_s,
        "index_encode_chunks_per_s": n_chunks / (t1 - t0), "index_add_s": t2 - t1,
        "first_batch_ms": first, "retrieve_ms": (t4 - t3) * 1e3, "extract_ms": (t5 - t4) * 1e3,
        "queries_per_s": n_queries / (t5 - t3), "extractions_per_s": n_ext / (t5 - t4),
        "spans": int(sum(len(v) for d in spans for v in d.values())),
        "mean_nnz_doc": float(np.diff(csr[0]).mean()),
        "note": "strings in -> verbatim spans out through the plugin classes; host tokenisation included in every stage",
    }


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:3]]
    print(json.dumps(run(*a), indent=1))
