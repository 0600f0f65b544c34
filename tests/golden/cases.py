"""Seeded inputs of the golden cases (shared by make_golden.py, the CPU oracle tests and the GPU parity tests)."""
import functools
import hashlib

import numpy as np

from verbatim_rag_b200.synthetic import (BertSpec, ModernBertSpec, SyntheticTokenizer, csr_to_dicts, make_bert_mlm_weights,
                                         make_dense_corpus, make_dense_queries, make_modernbert_weights, make_sparse_rows)


@functools.lru_cache(maxsize=None)
def tokenizer(kind):
    return SyntheticTokenizer(kind)


@functools.lru_cache(maxsize=None)
def span_cfg1(seed: int = 1001):
    """BASELINE config 1: 4 questions x 8 chunks of exactly 128 tokens (questions 12-20 words)."""
    tk = tokenizer("modernbert")
    spec = ModernBertSpec()
    rng = np.random.default_rng(seed)
    pairs = []
    for _ in range(4):
        q = tk.make_question(rng, int(rng.integers(12, 21)))
        for _ in range(8):
            pairs.append((q, tk.make_text(rng, 128)))
    seqs, n_q, offs = [], [], []
    for q, c in pairs:
        qe = tk.tok.encode(q, add_special_tokens=False)
        ce = tk.tok.encode(c, add_special_tokens=False)
        assert len(ce.ids) == 128
        seqs.append(np.asarray([tk.cls_id] + qe.ids + [tk.sep_id] + ce.ids + [tk.sep_id], dtype=np.int64))
        n_q.append(len(qe.ids))
        offs.append(list(ce.offsets))
    h = hashlib.sha256(("|".join(q + "#" + c for q, c in pairs)).encode()).hexdigest()
    return {"pairs": pairs, "seqs": seqs, "n_q": n_q, "ctx_offsets": offs, "spec": spec, "hash": h,
            "weights": make_modernbert_weights(seed, spec), "tokenizer": tk}


@functools.lru_cache(maxsize=None)
def splade_small(seed: int = 1002):
    tk = tokenizer("bert")
    spec = BertSpec()
    rng = np.random.default_rng(seed)
    texts = [tk.make_text(rng, n) for n in (64, 64, 17, 120, 254, 254, 8, 33, 300, 64, 200, 1)]
    seqs = []
    for t in texts:
        ids = tk.tok.encode(t, add_special_tokens=False).ids[:510]
        seqs.append(np.asarray([tk.cls_id] + ids + [tk.sep_id], dtype=np.int64))
    h = hashlib.sha256("|".join(texts).encode()).hexdigest()
    return {"texts": texts, "seqs": seqs, "spec": spec, "hash": h, "weights": make_bert_mlm_weights(seed, spec),
            "tokenizer": tk}


@functools.lru_cache(maxsize=None)
def topk_dense():
    return {"corpus": make_dense_corpus(20000, 768, seed=1004), "queries": make_dense_queries(16, 768, seed=2004), "k": 10}


@functools.lru_cache(maxsize=None)
def topk_sparse():
    corpus = make_sparse_rows(5000, seed=1002)
    q = make_sparse_rows(16, seed=2002, query=True)
    return {"corpus": corpus, "queries": q, "query_dicts": csr_to_dicts(*q), "k": 10}
