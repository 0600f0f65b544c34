"""Seeded inputs of the golden cases (shared by make_golden.py, the CPU oracle tests and the GPU parity tests)."""
import functools
import hashlib

import numpy as np

from verbatim_rag_b200.synthetic import (BertSpec, ModernBertSpec, SyntheticTokenizer, csr_to_dicts, make_bert_mlm_weights,
                                         make_dense_corpus, make_dense_queries, make_modernbert_weights, make_sparse_rows)


@functools.lru_cache(maxsize=None)
def tokenizer(kind):
    return SyntheticTokenizer(kind)


SPAN_CFG1_MARGIN = 2.5e-3   # every golden context token keeps at least this distance from the 0.2 threshold
SPAN_CFG1_CANDIDATES = 24   # candidate chunks generated per question; make_golden.py keeps the first 8 that qualify


def _span_case(pairs, tk, spec, seed):
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
def span_cfg1_candidates(seed: int = 1001):
    """4 questions (12-20 words) x SPAN_CFG1_CANDIDATES candidate chunks of exactly 128 tokens."""
    tk = tokenizer("modernbert")
    rng = np.random.default_rng(seed)
    pairs = []
    for _ in range(4):
        q = tk.make_question(rng, int(rng.integers(12, 21)))
        for _ in range(SPAN_CFG1_CANDIDATES):
            pairs.append((q, tk.make_text(rng, 128)))
    return _span_case(pairs, tk, ModernBertSpec(), seed)


@functools.lru_cache(maxsize=None)
def span_cfg1(seed: int = 1001):
    """BASELINE config 1: 4 questions x 8 chunks of exactly 128 tokens (questions 12-20 words).

    The 8 chunks of a question are the first 8 of its candidates whose oracle probabilities all stay at least
    SPAN_CFG1_MARGIN away from the threshold (chosen by make_golden.py, stored in span_cfg1_select.json), so that the
    span comparison of the parity tests is decided by the arithmetic and not by which side of 0.2 a rounding error
    lands on: with that margin every pair must match exactly."""
    import json
    import os
    sel = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "span_cfg1_select.json")))
    cand = span_cfg1_candidates(seed)
    assert sel["candidates_hash"] == cand["hash"], "candidate generator drifted: re-run make_golden.py"
    pairs = [cand["pairs"][qi * SPAN_CFG1_CANDIDATES + j] for qi, row in enumerate(sel["chunks"]) for j in row]
    return _span_case(pairs, cand["tokenizer"], cand["spec"], seed)


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
