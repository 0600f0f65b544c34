"""
B200VectorStore -- drop-in for ``LocalMilvusStore`` (verbatim_rag/vector_stores/milvus_local.py:12-167 on top of
``BaseMilvusStore``, milvus_base.py:28-500): same constructor flags, same ``add_vectors`` / ``query`` / ``delete``
behaviour and the duck-typed extras ``VerbatimIndex`` probes (``enable_full_text``, ``add_documents``,
``get_document``; verbatim_rag/index.py:60, 299-316, 667-669).

Vectors live in HBM (row-major fp32 dense block, CSR sparse block); exact top-k (dense COSINE, sparse IP -- the
metrics of milvus_local.py:111-125) runs through ``vrag_index_search_*``.  Payload (ids, texts, metadata) stays in
host Python lists, row-aligned with the device blocks.  The hybrid branches reuse the reference's own weighted-RRF
merge (vector_stores/hybrid_search.py:73-129) on top of the GPU top-k lists.
"""
from __future__ import annotations

import json
import logging
import threading
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _native
from .interfaces import (SearchResult, VectorStore, json_serialize_safe, merge_hybrid_results, promote_metadata,
                         sanitize_hybrid_weights)
from .models import parse_device

logger = logging.getLogger(__name__)

MAX_TEXT_LENGTH = 60000  # bytes; the reference truncates to the Milvus VARCHAR limit (milvus_base.py:20, 77-88)
_DYNAMIC_FIELDS = ["document_id", "user_id", "dataset_id"]


def _truncate(text: str, field: str, chunk_id: str) -> str:
    enc = text.encode("utf-8")
    if len(enc) <= MAX_TEXT_LENGTH:
        return text
    out = enc[:MAX_TEXT_LENGTH].decode("utf-8", errors="ignore")
    logger.warning("Truncating %s for chunk %s: %d bytes -> %d bytes", field, chunk_id, len(enc), len(out.encode("utf-8")))
    return out


class B200VectorStore(VectorStore):
    def __init__(
        self,
        db_path: Optional[str] = None,
        collection_name: str = "verbatim_rag",
        dense_dim: int = 384,
        enable_dense: bool = True,
        enable_sparse: bool = True,
        enable_full_text: bool = False,
        index_type: str = "FLAT",
        nlist: int = 8192,
        *,
        sparse_dim: int = 30522,
        device=None,
        id_base: int = 0,
    ):
        if enable_full_text:
            logger.warning("Full text search (BM25) is not supported by this store; it will be disabled.")
            enable_full_text = False
        if not enable_dense and not enable_sparse:
            raise ValueError("At least one of enable_dense or enable_sparse must be True")
        self.db_path = db_path  # accepted for signature compatibility; the index is memory (HBM) resident
        self.collection_name = collection_name
        self.documents_collection_name = f"{collection_name}_documents"
        self.dense_dim = dense_dim
        self.enable_dense = enable_dense
        self.enable_sparse = enable_sparse
        self.enable_full_text = enable_full_text
        self.index_type = index_type
        self.nlist = nlist
        self.sparse_dim = sparse_dim
        self._ctx = _native.default_context(parse_device(device))
        self._dense = _native.Index(self._ctx, _native.INDEX_DENSE_COSINE, dense_dim) if enable_dense else None
        self._sparse = _native.Index(self._ctx, _native.INDEX_SPARSE_IP, sparse_dim) if enable_sparse else None
        self._id_base = int(id_base)
        for ix in (self._dense, self._sparse):
            if ix is not None:
                ix.set_id_base(self._id_base)
        self._ids: List[str] = []
        self._texts: List[str] = []
        self._enh: List[str] = []
        self._meta: List[Dict[str, Any]] = []
        self._promoted: List[Dict[str, Any]] = []
        self._alive: List[bool] = []
        self._row_of: Dict[str, int] = {}
        self._documents: Dict[str, Dict[str, Any]] = {}
        self._lock = threading.RLock()

    # ------------------------------------------------------------------------------------------ insert
    def add_vectors(self, ids, dense_vectors, sparse_vectors, texts, enhanced_texts, metadatas):
        if self.enable_dense and (dense_vectors is None or len(dense_vectors) == 0):
            raise ValueError("Dense vectors required but not provided")
        if self.enable_sparse and (sparse_vectors is None or len(sparse_vectors) == 0):
            raise ValueError("Sparse vectors required but not provided")
        csr = None
        if self.enable_sparse:
            indptr = np.zeros(len(ids) + 1, dtype=np.int64)
            idx: List[int] = []
            val: List[float] = []
            for i, sv in enumerate(sparse_vectors):
                ks = sorted(sv.keys())
                idx.extend(int(k) for k in ks)
                val.extend(float(sv[k]) for k in ks)
                indptr[i + 1] = len(idx)
            csr = (indptr, np.asarray(idx, dtype=np.int32), np.asarray(val, dtype=np.float32))
        dense = np.asarray(dense_vectors, dtype=np.float32) if self.enable_dense else None
        self._insert(ids, dense, csr, texts, enhanced_texts, metadatas)

    def add_csr(self, ids, indptr, indices, values, texts, enhanced_texts, metadatas, dense=None):
        """Bulk insert of sparse rows already in CSR form (``B200SpladeProvider.embed_batch_csr``)."""
        self._insert(ids, dense, (indptr, indices, values), texts, enhanced_texts, metadatas)

    def _insert(self, ids, dense, csr, texts, enhanced_texts, metadatas):
        n = len(ids)
        with self._lock:
            if self._dense is not None:
                if dense is None or dense.shape != (n, self.dense_dim):
                    raise ValueError(f"dense vectors must be [{n}, {self.dense_dim}]")
                self._dense.add_dense(dense)
            if self._sparse is not None:
                if csr is None or len(csr[0]) != n + 1:
                    raise ValueError("sparse vectors missing / wrong row count")
                self._sparse.add_sparse(*csr)
            for i in range(n):
                promoted, cleaned = promote_metadata(metadatas[i])
                cid = ids[i]
                old = self._row_of.get(cid)
                if old is not None:   # primary-key upsert: the newest row wins
                    self._kill_rows([old])
                self._row_of[cid] = len(self._ids)
                self._ids.append(cid)
                self._texts.append(_truncate(texts[i], "text", cid))
                self._enh.append(_truncate(enhanced_texts[i], "enhanced_text", cid))
                self._meta.append(json_serialize_safe(cleaned))
                self._promoted.append(promoted)
                self._alive.append(True)
        logger.info("Added %d vectors to B200VectorStore", n)

    def _kill_rows(self, rows: Sequence[int]):
        rows = [r for r in rows if self._alive[r]]
        if not rows:
            return
        for ix in (self._dense, self._sparse):
            if ix is not None:
                ix.mark_deleted(rows)
        for r in rows:
            self._alive[r] = False

    # ------------------------------------------------------------------------------------------ documents
    def add_documents(self, documents: List[Dict[str, Any]]):
        for doc in documents or []:
            md = doc.get("metadata", {})
            row = {
                "id": doc.get("id", ""),
                "title": doc.get("title") or "",
                "source": doc.get("source") or "",
                "content_type": doc.get("doc_type") or doc.get("content_type") or "",
                "raw_content": doc.get("raw_content", ""),
                "metadata": json_serialize_safe(md) if isinstance(md, dict) else md,
            }
            with self._lock:
                self._documents[row["id"]] = row

    def add_document_schema(self, document_dict: Dict[str, Any], doc_id: str = None):
        if doc_id:
            document_dict["id"] = doc_id
        self.add_documents([document_dict])

    def get_document(self, document_id: str) -> Optional[Dict[str, Any]]:
        return self._documents.get(document_id)

    def get_all_documents(self) -> List[Dict[str, Any]]:
        return list(self._documents.values())[:1000]

    # ------------------------------------------------------------------------------------------ search
    def _hits(self, ids: np.ndarray, scores: np.ndarray, drop_zero: bool) -> List[Dict[str, Any]]:
        out = []
        for gid, sc in zip(ids.tolist(), scores.tolist()):
            if gid < 0:
                continue
            if drop_zero and sc == 0.0:   # an inverted index never visits rows that share no term with the query
                continue
            r = gid - self._id_base
            ent = {"text": self._texts[r], "enhanced_text": self._enh[r], "metadata": dict(self._meta[r])}
            ent.update(self._promoted[r])
            out.append({"id": self._ids[r], "distance": sc, "entity": ent})
        return out

    def _search_dense(self, q: Sequence[float], limit: int) -> List[Dict[str, Any]]:
        ids, sc = self._dense.search_dense(np.asarray(q, dtype=np.float32)[None, :], limit)
        return self._hits(ids[0], sc[0], drop_zero=False)

    def _search_sparse(self, q: Dict[int, float], limit: int) -> List[Dict[str, Any]]:
        ks = sorted(int(k) for k in q.keys() if 0 <= int(k) < self.sparse_dim)
        indptr = np.asarray([0, len(ks)], dtype=np.int64)
        ids, sc = self._sparse.search_sparse(indptr, np.asarray(ks, np.int32),
                                             np.asarray([q[k] for k in ks], np.float32), limit)
        return self._hits(ids[0], sc[0], drop_zero=True)

    def _to_results(self, hits: List[Dict[str, Any]]) -> List[SearchResult]:
        out = []
        for h in hits:
            ent = h.get("entity", {})
            md = ent.get("metadata", {}) or {}
            if isinstance(md, str):
                try:
                    md = json.loads(md)
                except Exception:
                    md = {"raw": md}
            for f in _DYNAMIC_FIELDS:
                if ent.get(f) is not None:
                    md[f] = ent[f]
            out.append(SearchResult(id=h.get("id"), score=h.get("distance", 0.0), text=ent.get("text", ""),
                                    enhanced_text=ent.get("enhanced_text", ""), metadata=md))
        return out

    def query(
        self,
        dense_query: Optional[List[float]] = None,
        sparse_query: Optional[Dict[int, float]] = None,
        text_query: Optional[str] = None,
        top_k: int = 5,
        search_type: str = "hybrid",
        filter: Optional[str] = None,
        search_params: Optional[Dict[str, Any]] = None,
        hybrid_weights: Optional[Dict[str, float]] = None,
        rrf_k: int = 60,
    ) -> List[SearchResult]:
        """Same branch structure as BaseMilvusStore.query (milvus_base.py:189-313)."""
        has_dense = dense_query is not None and len(dense_query) > 0
        has_sparse = sparse_query is not None and len(sparse_query) > 0
        if filter and (has_dense or has_sparse):
            raise NotImplementedError("metadata filter pushdown into the GPU scan is not implemented (filter-only "
                                      "browsing is); pass filter=None for vector search")
        with self._lock:
            if hybrid_weights is not None:
                weights = sanitize_hybrid_weights(hybrid_weights)
                weights = {k: v for k, v in weights.items() if k != "full_text"}
                if not weights:
                    raise ValueError("No valid search methods in hybrid_weights")
                by_method = {}
                if "dense" in weights and dense_query is not None and self._dense is not None:
                    by_method["dense"] = self._search_dense(dense_query, top_k * 2)
                if "sparse" in weights and sparse_query is not None and self._sparse is not None:
                    by_method["sparse"] = self._search_sparse(sparse_query, top_k * 2)
                if not by_method:
                    logger.warning("Hybrid search: no valid methods executed after validation")
                    return []
                if len(by_method) == 1:
                    return self._to_results(list(by_method.values())[0][:top_k])
                return self._to_results(merge_hybrid_results(by_method, top_k, weights, rrf_k,
                                                             log_label=self.__class__.__name__))
            if not has_dense and not has_sparse:
                return self._filter_only_query(filter, top_k)
            if search_type == "dense" and has_dense:
                hits = self._search_dense(dense_query, top_k)
            elif search_type == "sparse" and has_sparse:
                hits = self._search_sparse(sparse_query, top_k)
            elif search_type == "hybrid" and has_dense and has_sparse:
                by_method = {"dense": self._search_dense(dense_query, top_k * 2),
                             "sparse": self._search_sparse(sparse_query, top_k * 2)}
                hits = merge_hybrid_results(by_method, top_k, {"dense": 0.5, "sparse": 0.5}, rrf_k=rrf_k,
                                            log_label=self.__class__.__name__)
            else:
                raise ValueError(f"Invalid search configuration: type={search_type}, "
                                 f"dense={dense_query is not None}, sparse={sparse_query is not None}")
            return self._to_results(hits)

    # -- batched search: many queries per corpus pass ---------------------------------
    def query_batch_dense(self, queries: np.ndarray, top_k: int = 5) -> List[List[SearchResult]]:
        with self._lock:
            ids, sc = self._dense.search_dense(np.asarray(queries, np.float32), top_k)
            return [self._to_results(self._hits(ids[i], sc[i], False)) for i in range(ids.shape[0])]

    def query_batch_sparse(self, queries: Sequence[Dict[int, float]], top_k: int = 5) -> List[List[SearchResult]]:
        indptr = np.zeros(len(queries) + 1, np.int64)
        idx: List[int] = []
        val: List[float] = []
        for i, q in enumerate(queries):
            ks = sorted(int(k) for k in q.keys() if 0 <= int(k) < self.sparse_dim)
            idx.extend(ks)
            val.extend(float(q[k]) for k in ks)
            indptr[i + 1] = len(idx)
        with self._lock:
            ids, sc = self._sparse.search_sparse(indptr, np.asarray(idx, np.int32), np.asarray(val, np.float32), top_k)
            return [self._to_results(self._hits(ids[i], sc[i], True)) for i in range(ids.shape[0])]

    # ------------------------------------------------------------------------------------------ browse / delete
    def _filter_only_query(self, filter: Optional[str], limit: int) -> List[SearchResult]:
        """No vectors given: browse up to ``limit`` live rows, score 1.0 (milvus_base.py:315-353).
        Supported filter forms: None/"", ``field == "v"``, ``id in ["a", ...]`` on id / promoted fields."""
        try:
            pred = _compile_filter(filter)
            out = []
            for r in range(len(self._ids)):
                if not self._alive[r]:
                    continue
                row = {"id": self._ids[r], **self._promoted[r]}
                if not pred(row, self._meta[r]):
                    continue
                md = dict(self._meta[r])
                for f in _DYNAMIC_FIELDS:
                    if self._promoted[r].get(f) is not None:
                        md[f] = self._promoted[r][f]
                out.append(SearchResult(id=self._ids[r], score=1.0, text=self._texts[r], enhanced_text=self._enh[r],
                                        metadata=md))
                if len(out) >= limit:
                    break
            return out
        except Exception as e:  # reference convention (milvus_base.py:351-353)
            logger.error("Failed to query chunks: %s", e)
            return []

    def delete(self, ids: List[str]):
        if not ids:
            return
        with self._lock:
            rows = [self._row_of.pop(i) for i in ids if i in self._row_of]
            self._kill_rows(rows)

    def delete_document(self, document_id: str):
        with self._lock:
            rows = [r for r in range(len(self._ids)) if self._alive[r] and self._promoted[r].get("document_id") == document_id]
            for r in rows:
                self._row_of.pop(self._ids[r], None)
            self._kill_rows(rows)
            self._documents.pop(document_id, None)

    def __len__(self) -> int:
        return sum(self._alive)


def _compile_filter(expr: Optional[str]):
    """Tiny subset of the Milvus boolean expression language used by the reference's own callers
    (verbatim_rag/index.py:735-738): ``field == "value"`` / ``field == number`` / ``field in [..]``,
    ``metadata["key"] <op> value``, joined by ``and``."""
    if not expr or not expr.strip():
        return lambda row, md: True
    import ast
    import re

    clauses = [c.strip() for c in re.split(r"\s+and\s+|\s*&&\s*", expr.strip()) if c.strip()]
    tests = []
    for c in clauses:
        m = re.match(r'^(metadata\[\s*["\'](?P<mk>[^"\']+)["\']\s*\]|(?P<f>\w+))\s*(?P<op>==|!=|>=|<=|>|<|in)\s*(?P<v>.+)$', c)
        if not m:
            raise ValueError(f"unsupported filter clause: {c!r}")
        val = ast.literal_eval(m.group("v"))
        tests.append((m.group("mk"), m.group("f"), m.group("op"), val))

    def pred(row, md):
        for mk, f, op, val in tests:
            x = md.get(mk) if mk is not None else row.get(f, md.get(f))
            try:
                ok = {"==": lambda: x == val, "!=": lambda: x != val, ">=": lambda: x >= val, "<=": lambda: x <= val,
                      ">": lambda: x > val, "<": lambda: x < val, "in": lambda: x in val}[op]()
            except TypeError:
                ok = False
            if not ok:
                return False
        return True

    return pred
