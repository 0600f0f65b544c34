"""
Oracle: exact (brute-force) top-k, dense COSINE and sparse IP.  TEST INFRASTRUCTURE.

Restates what backs ``LocalMilvusStore`` (verbatim_rag/vector_stores/milvus_local.py:111-125:
dense metric COSINE, sparse SPARSE_INVERTED_INDEX / IP) as searched by
``BaseMilvusStore.query`` (milvus_base.py:239-259).  milvus-lite 2.5.1 (third-party, not
installed) runs a FLAT exact scan for dense and an exact inverted-index IP for sparse
(SURVEY.md App. B.3); both return the k best by similarity, larger = better, descending.

Definition used by this oracle and by the GPU path:
* score is evaluated in float64 from the stored fp32 values, reported as float32;
* order is (score descending, insertion index ascending) -- milvus leaves ties unspecified;
* a zero-norm dense vector has cosine 0.
"""

from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def _order(scores64: np.ndarray, k: int) -> np.ndarray:
    """indices of the k best: score desc, index asc (stable)."""
    n = scores64.shape[0]
    k = min(k, n)
    if k <= 0:
        return np.zeros((0,), dtype=np.int64)
    if n > 4 * k + 64:
        # everything >= the k-th best value (keeps all ties of the k-th value), then exact ordering
        kth = np.partition(scores64, n - k)[n - k]
        cand = np.nonzero(scores64 >= kth)[0]
    else:
        cand = np.arange(n)
    o = np.lexsort((cand, -scores64[cand]))
    return cand[o[:k]].astype(np.int64)


def dense_cosine_scores(corpus: np.ndarray, queries: np.ndarray, block: int = 65536) -> np.ndarray:
    """[Q, N] float64 cosine similarities of fp32 rows."""
    q64 = np.asarray(queries, dtype=np.float64)
    qn = np.sqrt((q64 * q64).sum(axis=1))
    N = corpus.shape[0]
    out = np.empty((q64.shape[0], N), dtype=np.float64)
    for s in range(0, N, block):
        c64 = np.asarray(corpus[s:s + block], dtype=np.float64)
        cn = np.sqrt((c64 * c64).sum(axis=1))
        dots = q64 @ c64.T
        den = qn[:, None] * cn[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            sc = np.where(den > 0, dots / den, 0.0)
        out[:, s:s + block] = sc
    return out


def dense_cosine_topk(corpus: np.ndarray, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """ids [Q, k] int64 (row indices), scores [Q, k] float32."""
    sc = dense_cosine_scores(corpus, queries)
    kk = min(k, corpus.shape[0])
    ids = np.zeros((sc.shape[0], kk), dtype=np.int64)
    out = np.zeros((sc.shape[0], kk), dtype=np.float32)
    for qi in range(sc.shape[0]):
        o = _order(sc[qi], k)
        ids[qi] = o
        out[qi] = sc[qi, o].astype(np.float32)
    return ids, out


def sparse_ip_scores(indptr: np.ndarray, indices: np.ndarray, values: np.ndarray, dim: int,
                     queries: Sequence[Dict[int, float]]) -> np.ndarray:
    """[Q, N] float64 inner products of CSR fp32 rows with dict queries (values rounded to fp32 first)."""
    import scipy.sparse as sp

    n = len(indptr) - 1
    m = sp.csr_matrix((np.asarray(values, dtype=np.float64), np.asarray(indices), np.asarray(indptr)), shape=(n, dim))
    out = np.empty((len(queries), n), dtype=np.float64)
    for qi, q in enumerate(queries):
        dense = np.zeros(dim, dtype=np.float64)
        for t, wv in q.items():
            if 0 <= int(t) < dim:
                dense[int(t)] = np.float64(np.float32(wv))
        out[qi] = m @ dense
    return out


def sparse_ip_topk(indptr, indices, values, dim: int, queries: Sequence[Dict[int, float]], k: int):
    sc = sparse_ip_scores(indptr, indices, values, dim, queries)
    kk = min(k, sc.shape[1])
    ids = np.zeros((sc.shape[0], kk), dtype=np.int64)
    out = np.zeros((sc.shape[0], kk), dtype=np.float32)
    for qi in range(sc.shape[0]):
        o = _order(sc[qi], k)
        ids[qi] = o
        out[qi] = sc[qi, o].astype(np.float32)
    return ids, out


def dicts_to_csr(rows: Sequence[Dict[int, float]]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    idx: List[int] = []
    val: List[float] = []
    for i, r in enumerate(rows):
        ks = sorted(int(k) for k in r.keys())
        idx.extend(ks)
        val.extend(float(r[k]) for k in ks)
        indptr[i + 1] = len(idx)
    return indptr, np.asarray(idx, dtype=np.int32), np.asarray(val, dtype=np.float32)
