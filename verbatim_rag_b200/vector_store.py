"""
B200VectorStore -- drop-in for ``LocalMilvusStore`` (verbatim_rag/vector_stores/milvus_local.py:12-167 on top of
``BaseMilvusStore``, milvus_base.py:28-500): same constructor flags, same ``add_vectors`` / ``query`` / ``delete``
behaviour and the duck-typed extras ``VerbatimIndex`` probes (``enable_full_text``, ``add_documents``,
``get_document``; verbatim_rag/index.py:60, 299-316, 667-669).

Vectors live in HBM (row-major fp32 dense block, CSR sparse block); exact top-k (dense COSINE, sparse IP -- the
metrics of milvus_local.py:111-125) runs through ``vrag_index_search_*``.  Payload (ids, texts, metadata) stays in
host Python lists, row-aligned with the device blocks.  The hybrid branches reuse the reference's own weighted-RRF
merge (vector_stores/hybrid_search.py:73-129) on top of the GPU top-k lists.

``filter=`` (a Milvus boolean expression, milvus_base.py:189-259) is pushed down into the scan: the expression is
evaluated once over the host payload columns, the resulting row mask goes to ``vrag_index_set_filter`` and the scan
kernels skip masked rows like deleted ones.

``db_path`` (the milvus-lite ``.db`` file of milvus_local.py:61-63) names a DIRECTORY holding an append-only store
(SURVEY.md 8f-2): ``dense.f32`` (raw row-major fp32, mmap-able, loads straight into HBM), ``sparse.indptr.i64`` /
``sparse.indices.i32`` / ``sparse.values.f32`` (CSR), ``payload.jsonl`` (id, texts, metadata per row),
``tombstones.i64`` (deleted row numbers), ``documents.jsonl`` and ``manifest.json``.  Every ``add_vectors`` /
``delete`` appends before it returns, so re-opening the same path restores the collection, as with the reference.
"""
from __future__ import annotations

import contextlib
import json
import logging
import os
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


def dicts_to_csr(sparse_vectors: Sequence[Dict[int, float]]):
    """The reference's ``List[Dict[int, float]]`` sparse rows (embedding_providers.py:161-163) -> CSR (indptr int64,
    indices int64 ascending per row, values fp32)."""
    indptr = np.zeros(len(sparse_vectors) + 1, dtype=np.int64)
    idx: List[int] = []
    val: List[float] = []
    for i, sv in enumerate(sparse_vectors):
        ks = sorted(sv.keys())
        idx.extend(int(k) for k in ks)
        val.extend(float(sv[k]) for k in ks)
        indptr[i + 1] = len(idx)
    return indptr, np.asarray(idx, dtype=np.int64), np.asarray(val, dtype=np.float32)


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
        self.db_path = db_path  # None: memory (HBM) only; else the directory of the append-only store (see module doc)
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
        self._mask_cache: Optional[tuple] = None   # (filter, rows, deletions) -> exclude mask
        self._n_deleted = 0
        self._broken: Optional[str] = None       # set when a failed insert left the store in a state it cannot repair
        self._columns: Dict[Any, Any] = {}       # filter pushdown: payload columns as numpy arrays (see _column)
        self._store: Optional[_DiskStore] = None
        if db_path:
            self._store = _DiskStore(db_path, collection_name, dense_dim if enable_dense else 0,
                                     sparse_dim if enable_sparse else 0, self._id_base)
            self._restore()

    # ------------------------------------------------------------------------------------------ persistence
    def _restore(self):
        st = self._store
        n = st.rows
        if n == 0:
            return
        payload = st.read_payload(n)
        if self._dense is not None:
            for a, b in st.dense_chunks(n):
                self._dense.add_dense(st.dense_rows(a, b))
        if self._sparse is not None:
            self._sparse.add_sparse(*st.sparse_csr(n))
        for i, row in enumerate(payload):
            cid = row["id"]
            self._row_of[cid] = i            # later rows with the same id win; the older ones are tombstoned on disk
            self._ids.append(cid)
            self._texts.append(row["text"])
            self._enh.append(row["enhanced_text"])
            self._meta.append(row["metadata"])
            self._promoted.append(row["promoted"])
            self._alive.append(True)
        all_dead = st.read_tombstones()
        dead = [r for r in all_dead if 0 <= r < n]
        if len(dead) != len(all_dead):   # tombstones past the recovered rows (crash mid-insert): new rows will reuse those
            st.rewrite_tombstones(dead)  # numbers, so the stale entries must not survive
        self._kill_rows(sorted(set(dead)), persist=False)
        for r in dead:
            if self._row_of.get(self._ids[r]) == r:
                del self._row_of[self._ids[r]]
        for doc in st.read_documents():
            self._documents[doc["id"]] = doc
        logger.info("Restored %d rows (%d live) from %s", n, sum(self._alive), self.db_path)

    # ------------------------------------------------------------------------------------------ insert
    def add_vectors(self, ids, dense_vectors, sparse_vectors, texts, enhanced_texts, metadatas):
        if self.enable_dense and (dense_vectors is None or len(dense_vectors) == 0):
            raise ValueError("Dense vectors required but not provided")
        if self.enable_sparse and (sparse_vectors is None or len(sparse_vectors) == 0):
            raise ValueError("Sparse vectors required but not provided")
        csr = None
        if self.enable_sparse:
            if len(sparse_vectors) != len(ids):
                raise ValueError(f"{len(sparse_vectors)} sparse vectors for {len(ids)} ids")
            csr = dicts_to_csr(sparse_vectors)
        dense = np.asarray(dense_vectors, dtype=np.float32) if self.enable_dense else None
        self._insert(ids, dense, csr, texts, enhanced_texts, metadatas)

    def add_csr(self, ids, indptr, indices, values, texts, enhanced_texts, metadatas, dense=None):
        """Bulk insert of sparse rows already in CSR form (``B200SpladeProvider.embed_batch_csr``)."""
        self._insert(ids, dense, (indptr, indices, values), texts, enhanced_texts, metadatas)

    def _insert(self, ids, dense, csr, texts, enhanced_texts, metadatas, local=None):
        """All-or-nothing insert.  ``local`` (sharded store only): ``(lo, hi)`` -- the vectors given are those of rows
        [lo, hi) of this batch, the payload is complete.  Everything that can be wrong with the INPUT is checked before the device is touched
        (list lengths, dense shape, CSR shape / monotonicity, term ids inside [0, sparse_dim)), and the host payload is
        prepared first.  A failure after that point can only come from the device (out of memory) or the disk; the
        rows that were already appended to one index are then tombstoned and padded so that the dense block, the
        sparse block and the payload lists keep the same row numbering."""
        n = len(ids)
        nv = n if local is None else local[1] - local[0]   # rows the vectors cover
        if not (len(texts) == len(enhanced_texts) == len(metadatas) == n):
            raise ValueError(f"ids / texts / enhanced_texts / metadatas must have the same length ({n}, {len(texts)}, "
                             f"{len(enhanced_texts)}, {len(metadatas)})")
        if self._broken:
            raise RuntimeError(f"B200VectorStore is unusable after a failed insert: {self._broken}")
        if self._dense is not None:
            if dense is None:
                raise ValueError(f"dense vectors must be [{nv}, {self.dense_dim}]")
            dense = np.ascontiguousarray(dense, dtype=np.float32)
            if dense.shape != (nv, self.dense_dim):
                raise ValueError(f"dense vectors must be [{nv}, {self.dense_dim}]")
        if self._sparse is not None:
            if csr is None or len(csr[0]) != nv + 1:
                raise ValueError("sparse vectors missing / wrong row count")
            indptr = np.ascontiguousarray(csr[0], dtype=np.int64)
            indices, values = np.asarray(csr[1]), np.ascontiguousarray(csr[2], dtype=np.float32)
            a, b = int(indptr[0]), int(indptr[-1])
            if a < 0 or b > len(indices) or len(indices) != len(values) or (nv and bool((np.diff(indptr) < 0).any())):
                raise ValueError("sparse vectors: indptr must be non-decreasing and stay inside indices / values")
            if b > a and (int(indices[a:b].min()) < 0 or int(indices[a:b].max()) >= self.sparse_dim):
                raise ValueError(f"sparse vectors: term ids must lie in [0, {self.sparse_dim})")
            csr = (indptr, np.ascontiguousarray(indices, dtype=np.int32), values)
        rows = []
        for i in range(n):   # host payload first: metadata that cannot be handled fails before anything is stored
            promoted, cleaned = promote_metadata(metadatas[i])
            rows.append((ids[i], _truncate(texts[i], "text", ids[i]), _truncate(enhanced_texts[i], "enhanced_text", ids[i]),
                         json_serialize_safe(cleaned), promoted))
        with self._lock:
            first_row = len(self._ids)
            try:
                self._device_add(dense, csr, first_row, n, local)
            except Exception as exc:
                self._rollback(first_row, n, exc)
                raise
            dead: List[int] = []
            for cid, text, enh, meta, promoted in rows:
                old = self._row_of.get(cid)
                if old is not None:   # primary-key upsert: the newest row wins
                    dead.append(old)
                self._row_of[cid] = len(self._ids)
                self._ids.append(cid)
                self._texts.append(text)
                self._enh.append(enh)
                self._meta.append(meta)
                self._promoted.append(promoted)
                self._alive.append(True)
            self._kill_rows(dead, persist=False)
            if self._store is not None:
                try:
                    self._store.append_rows(
                        dense if self._dense is not None else None, csr if self._sparse is not None else None,
                        [{"id": r[0], "text": r[1], "enhanced_text": r[2], "metadata": r[3], "promoted": r[4]} for r in rows])
                    if dead:   # tombstones only after the rows they may refer to are durable
                        self._store.append_tombstones(dead)
                except Exception as exc:
                    # memory now holds rows the disk does not: searches keep working, persistence cannot be trusted
                    self._broken = f"disk append failed ({exc}); reopen the store from {self.db_path}"
                    raise
        logger.info("Added %d vectors to B200VectorStore", n)

    def _device_add(self, dense, csr, first_row: int, n: int, local=None):
        """Append the vectors of one insert batch to the device indexes (the sharded store keeps its slice only)."""
        if self._dense is not None:
            self._dense.add_dense(dense)
        if self._sparse is not None:
            self._sparse.add_sparse(*csr)

    def _rollback(self, first_row: int, n: int, exc: Exception):
        """Bring both indexes to first_row + n rows, all n of them tombstoned, with dead placeholder payload, so that
        row numbers stay aligned after a device-side failure in the middle of an insert."""
        try:
            for ix, kind in ((self._dense, "dense"), (self._sparse, "sparse")):
                if ix is None:
                    continue
                missing = first_row + n - len(ix)
                if missing > 0:
                    if kind == "dense":
                        ix.add_dense(np.zeros((missing, self.dense_dim), np.float32))
                    else:
                        ix.add_sparse(np.zeros(missing + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32))
                ix.mark_deleted(list(range(first_row, first_row + n)))
            for i in range(n):
                self._ids.append(f"__failed_insert_{first_row + i}")
                self._texts.append("")
                self._enh.append("")
                self._meta.append({})
                self._promoted.append({})
                self._alive.append(False)
            self._n_deleted += n
            if self._store is not None:
                self._broken = f"insert failed on the device ({exc}); the on-disk store was left untouched, reopen it"
        except Exception as exc2:  # the padding failed too (e.g. still out of memory): refuse further use
            self._broken = f"insert failed ({exc}) and the indexes could not be re-aligned ({exc2})"

    def _kill_rows(self, rows: Sequence[int], persist: bool = True):
        rows = [r for r in rows if self._alive[r]]
        if not rows:
            return
        for ix in (self._dense, self._sparse):
            if ix is not None:
                ix.mark_deleted(rows)
        for r in rows:
            self._alive[r] = False
        self._n_deleted += len(rows)
        if persist and self._store is not None:
            self._store.append_tombstones(rows)

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
                if self._store is not None:
                    self._store.append_document(row)

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
        with self._lock, self._filtered(filter if (has_dense or has_sparse) else None):
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

    # -- metadata filter pushdown -------------------------------------------------------
    def _column(self, meta_key: Optional[str], field: Optional[str]) -> np.ndarray:
        """Payload column as an object array, row-aligned (built once per field and extended as rows arrive), so that a
        filter expression costs a few vectorised numpy passes instead of a Python call per row."""
        key = (meta_key, field)
        have = self._columns.get(key)
        n = len(self._ids)
        start = 0 if have is None else len(have)
        if start < n:
            if meta_key is not None:
                new = [self._meta[r].get(meta_key) for r in range(start, n)]
            elif field == "id":
                new = self._ids[start:n]
            else:
                new = [self._promoted[r].get(field, self._meta[r].get(field)) for r in range(start, n)]
            col = np.empty(n - start, dtype=object)
            col[:] = new
            have = col if have is None else np.concatenate([have, col])
            self._columns[key] = have
        return have

    def _exclude_mask(self, filter: Optional[str]) -> Optional[np.ndarray]:
        """Row mask for the scan kernels: 1 = skip.  Evaluated on the host payload columns, cached per
        (expression, store state)."""
        if not filter or not filter.strip():
            return None
        key = (filter, len(self._ids), self._n_deleted)
        if self._mask_cache is not None and self._mask_cache[0] == key:
            return self._mask_cache[1]
        keep = np.asarray(self._alive, dtype=bool)
        for mk, f, op, val in _parse_filter(filter):
            col = self._column(mk, f)
            if op == "in":
                members = set(val) if _hashable_all(val) else list(val)
                ok = np.fromiter((x in members for x in col), bool, count=len(col))
            elif op in ("==", "!="):
                ok = col == val
                if not isinstance(ok, np.ndarray):   # numpy falls back to a scalar for exotic operands
                    ok = np.fromiter((x == val for x in col), bool, count=len(col))
                if op == "!=":
                    ok = ~ok
            else:   # ordering: rows whose value cannot be compared with the literal do not match (TypeError -> False)
                num = isinstance(val, (int, float))   # Python semantics, as in the row predicate: bool is an int
                comparable = np.fromiter((isinstance(x, (int, float)) if num else isinstance(x, type(val)) for x in col),
                                         bool, count=len(col))
                ok = np.zeros(len(col), bool)
                if comparable.any():
                    sub = col[comparable]
                    arr = sub.astype(np.float64) if num else sub
                    ok[comparable] = {">=": arr >= val, "<=": arr <= val, ">": arr > val, "<": arr < val}[op]
            keep &= ok
        mask = (~keep).astype(np.uint8)
        self._mask_cache = (key, mask)
        return mask

    @contextlib.contextmanager
    def _filtered(self, filter: Optional[str]):
        mask = self._exclude_mask(filter)
        live = [ix for ix in (self._dense, self._sparse) if ix is not None] if mask is not None else []
        for ix in live:
            ix.set_filter(mask)
        try:
            yield
        finally:
            for ix in live:
                ix.set_filter(None)

    # -- batched search: many queries per corpus pass ---------------------------------
    def _hits_dense_batch(self, queries, limit: int) -> List[List[Dict[str, Any]]]:
        ids, sc = self._dense.search_dense(np.asarray(queries, np.float32).reshape(-1, self.dense_dim), limit)
        return [self._hits(ids[i], sc[i], drop_zero=False) for i in range(ids.shape[0])]

    def _hits_sparse_batch(self, queries: Sequence[Dict[int, float]], limit: int) -> List[List[Dict[str, Any]]]:
        indptr = np.zeros(len(queries) + 1, np.int64)
        idx: List[int] = []
        val: List[float] = []
        for i, q in enumerate(queries):
            ks = sorted(int(k) for k in q.keys() if 0 <= int(k) < self.sparse_dim)
            idx.extend(ks)
            val.extend(float(q[k]) for k in ks)
            indptr[i + 1] = len(idx)
        ids, sc = self._sparse.search_sparse(indptr, np.asarray(idx, np.int32), np.asarray(val, np.float32), limit)
        return [self._hits(ids[i], sc[i], drop_zero=True) for i in range(ids.shape[0])]

    def query_batch_dense(self, queries: np.ndarray, top_k: int = 5, filter: Optional[str] = None
                          ) -> List[List[SearchResult]]:
        with self._lock, self._filtered(filter):
            return [self._to_results(h) for h in self._hits_dense_batch(queries, top_k)]

    def query_batch_sparse(self, queries: Sequence[Dict[int, float]], top_k: int = 5, filter: Optional[str] = None
                           ) -> List[List[SearchResult]]:
        with self._lock, self._filtered(filter):
            return [self._to_results(h) for h in self._hits_sparse_batch(queries, top_k)]

    def query_batch(
        self,
        dense_queries=None,
        sparse_queries: Optional[Sequence[Dict[int, float]]] = None,
        text_queries: Optional[Sequence[str]] = None,
        top_k: int = 5,
        search_type: str = "hybrid",
        filter: Optional[str] = None,
        search_params: Optional[Dict[str, Any]] = None,
        hybrid_weights: Optional[Dict[str, float]] = None,
        rrf_k: int = 60,
    ) -> List[List[SearchResult]]:
        """``[self.query(dense_queries[i], sparse_queries[i], ...) for i]`` with one corpus scan per 16 queries instead
        of one per query: same branch structure as ``query`` / BaseMilvusStore.query (milvus_base.py:189-313)."""
        nq = len(dense_queries) if dense_queries is not None else (len(sparse_queries) if sparse_queries is not None else 0)
        has_dense = dense_queries is not None and nq > 0
        has_sparse = sparse_queries is not None and nq > 0
        if not has_dense and not has_sparse:
            n = len(text_queries) if text_queries is not None else 0
            return [self.query(top_k=top_k, filter=filter) for _ in range(n)]
        with self._lock, self._filtered(filter):
            if hybrid_weights is not None:
                weights = sanitize_hybrid_weights(hybrid_weights)
                weights = {k: v for k, v in weights.items() if k != "full_text"}
                if not weights:
                    raise ValueError("No valid search methods in hybrid_weights")
                by_method = {}
                if "dense" in weights and has_dense and self._dense is not None:
                    by_method["dense"] = self._hits_dense_batch(dense_queries, top_k * 2)
                if "sparse" in weights and has_sparse and self._sparse is not None:
                    by_method["sparse"] = self._hits_sparse_batch(sparse_queries, top_k * 2)
                if not by_method:
                    logger.warning("Hybrid search: no valid methods executed after validation")
                    return [[] for _ in range(nq)]
                if len(by_method) == 1:
                    return [self._to_results(h[:top_k]) for h in list(by_method.values())[0]]
                return [self._to_results(merge_hybrid_results({m: by_method[m][i] for m in by_method}, top_k, weights,
                                                              rrf_k, log_label=self.__class__.__name__))
                        for i in range(nq)]
            if search_type == "dense" and has_dense:
                hits = self._hits_dense_batch(dense_queries, top_k)
            elif search_type == "sparse" and has_sparse:
                hits = self._hits_sparse_batch(sparse_queries, top_k)
            elif search_type == "hybrid" and has_dense and has_sparse:
                d = self._hits_dense_batch(dense_queries, top_k * 2)
                sp = self._hits_sparse_batch(sparse_queries, top_k * 2)
                hits = [merge_hybrid_results({"dense": d[i], "sparse": sp[i]}, top_k, {"dense": 0.5, "sparse": 0.5},
                                             rrf_k=rrf_k, log_label=self.__class__.__name__) for i in range(nq)]
            else:
                raise ValueError(f"Invalid search configuration: type={search_type}, "
                                 f"dense={dense_queries is not None}, sparse={sparse_queries is not None}")
            return [self._to_results(h) for h in hits]

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


class _DiskStore:
    """Append-only on-disk form of one collection (module docstring).  All files only ever grow; a torn tail (crash
    in the middle of an append) is cut back to the last complete row on open."""

    FORMAT = "vrag-b200-store"
    VERSION = 1
    CHUNK_ROWS = 1 << 16   # dense rows per host->device copy on restore

    def __init__(self, path: str, collection: str, dense_dim: int, sparse_dim: int, id_base: int):
        self.dir = path
        self.dense_dim, self.sparse_dim = int(dense_dim), int(sparse_dim)
        os.makedirs(path, exist_ok=True)
        man = os.path.join(path, "manifest.json")
        want = {"format": self.FORMAT, "version": self.VERSION, "collection_name": collection,
                "dense_dim": self.dense_dim, "sparse_dim": self.sparse_dim, "id_base": int(id_base)}
        if os.path.exists(man):
            with open(man) as f:
                have = json.load(f)
            for k in ("format", "version", "dense_dim", "sparse_dim"):
                if have.get(k) != want[k]:
                    raise ValueError(f"{path}: stored {k}={have.get(k)!r} does not match this store's {want[k]!r}")
        else:
            with open(man, "w") as f:
                json.dump(want, f, indent=1)
        self.rows = self._count_rows()

    def _p(self, name: str) -> str:
        return os.path.join(self.dir, name)

    def _size(self, name: str) -> int:
        return os.path.getsize(self._p(name)) if os.path.exists(self._p(name)) else 0

    def _count_rows(self) -> int:
        n = 0
        if os.path.exists(self._p("payload.jsonl")):
            with open(self._p("payload.jsonl"), "rb") as f:
                data = f.read()
            n = data.count(b"\n")
        if self.dense_dim:
            n = min(n, self._size("dense.f32") // (4 * self.dense_dim))
        if self.sparse_dim:
            n = min(n, self._size("sparse.indptr.i64") // 8)
            if n:
                ip = np.memmap(self._p("sparse.indptr.i64"), np.int64, "r", shape=(n,))
                nnz_have = min(self._size("sparse.indices.i32"), self._size("sparse.values.f32")) // 4
                while n and int(ip[n - 1]) > nnz_have:
                    n -= 1
        return n

    # -- read
    def read_payload(self, n: int) -> List[Dict[str, Any]]:
        out = []
        with open(self._p("payload.jsonl"), encoding="utf-8") as f:
            for line in f:
                if len(out) == n:
                    break
                out.append(json.loads(line))
        return out

    def dense_chunks(self, n: int):
        for a in range(0, n, self.CHUNK_ROWS):
            yield a, min(n, a + self.CHUNK_ROWS)

    def dense_rows(self, a: int, b: int) -> np.ndarray:
        mm = np.memmap(self._p("dense.f32"), np.float32, "r", offset=a * self.dense_dim * 4,
                       shape=(b - a, self.dense_dim))
        return np.ascontiguousarray(mm)

    def sparse_csr(self, n: int):
        ends = np.fromfile(self._p("sparse.indptr.i64"), np.int64, count=n)
        indptr = np.concatenate([np.zeros(1, np.int64), ends])
        nnz = int(indptr[-1])
        idx = np.fromfile(self._p("sparse.indices.i32"), np.int32, count=nnz) if nnz else np.zeros(0, np.int32)
        val = np.fromfile(self._p("sparse.values.f32"), np.float32, count=nnz) if nnz else np.zeros(0, np.float32)
        return indptr, idx, val

    def read_tombstones(self) -> List[int]:
        if not os.path.exists(self._p("tombstones.i64")):
            return []
        return np.fromfile(self._p("tombstones.i64"), np.int64).tolist()

    def read_documents(self) -> List[Dict[str, Any]]:
        if not os.path.exists(self._p("documents.jsonl")):
            return []
        with open(self._p("documents.jsonl"), encoding="utf-8") as f:
            return [json.loads(line) for line in f if line.strip()]

    # -- append
    def _truncate_to_rows(self):
        """Cut every file back to `self.rows` complete rows before appending (drops a torn tail)."""
        n = self.rows
        if self.dense_dim and self._size("dense.f32") != n * 4 * self.dense_dim:
            with open(self._p("dense.f32"), "ab") as f:
                f.truncate(n * 4 * self.dense_dim)
        if self.sparse_dim:
            # indices / values are written BEFORE indptr: a crash in between leaves an orphan tail on them even though
            # indptr has its expected size, so each file is checked against its own expected length
            nnz = 0
            if n:
                nnz = int(np.memmap(self._p("sparse.indptr.i64"), np.int64, "r", shape=(n,))[n - 1])
            for name, item in (("sparse.indptr.i64", n * 8), ("sparse.indices.i32", nnz * 4), ("sparse.values.f32", nnz * 4)):
                if self._size(name) != item:
                    with open(self._p(name), "ab") as f:
                        f.truncate(item)
        if os.path.exists(self._p("payload.jsonl")):
            with open(self._p("payload.jsonl"), "rb") as f:
                data = f.read()
            pos, seen = 0, 0
            while seen < n:
                pos = data.index(b"\n", pos) + 1
                seen += 1
            if pos != len(data):
                with open(self._p("payload.jsonl"), "ab") as f:
                    f.truncate(pos)

    def append_rows(self, dense: Optional[np.ndarray], csr, payload: List[Dict[str, Any]]):
        self._truncate_to_rows()
        if self.dense_dim:
            with open(self._p("dense.f32"), "ab") as f:
                np.ascontiguousarray(dense, np.float32).tofile(f)
        if self.sparse_dim:
            indptr, idx, val = (np.asarray(x) for x in csr)
            a, b = int(indptr[0]), int(indptr[-1])
            base = self._size("sparse.indices.i32") // 4
            with open(self._p("sparse.indices.i32"), "ab") as f:
                np.ascontiguousarray(idx[a:b], np.int32).tofile(f)
            with open(self._p("sparse.values.f32"), "ab") as f:
                np.ascontiguousarray(val[a:b], np.float32).tofile(f)
            with open(self._p("sparse.indptr.i64"), "ab") as f:
                (np.asarray(indptr[1:], np.int64) - a + base).tofile(f)
        with open(self._p("payload.jsonl"), "a", encoding="utf-8") as f:   # last: a row counts once its payload line is complete
            for row in payload:
                f.write(json.dumps(row, ensure_ascii=False, default=str) + "\n")
        self.rows += len(payload)

    def append_tombstones(self, rows: Sequence[int]):
        with open(self._p("tombstones.i64"), "ab") as f:
            np.asarray(list(rows), np.int64).tofile(f)

    def rewrite_tombstones(self, rows: Sequence[int]):
        tmp = self._p("tombstones.i64.tmp")
        with open(tmp, "wb") as f:
            np.asarray(list(rows), np.int64).tofile(f)
        os.replace(tmp, self._p("tombstones.i64"))

    def append_document(self, doc: Dict[str, Any]):
        with open(self._p("documents.jsonl"), "a", encoding="utf-8") as f:
            f.write(json.dumps(doc, ensure_ascii=False, default=str) + "\n")


def _hashable_all(vals) -> bool:
    try:
        return all(isinstance(v, (str, int, float, bool)) for v in vals)
    except TypeError:
        return False


def _parse_filter(expr: str):
    """-> [(metadata key | None, field | None, op, literal)] for the clauses of ``expr`` (joined by ``and``)."""
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
    return tests


def _compile_filter(expr: Optional[str]):
    """Tiny subset of the Milvus boolean expression language used by the reference's own callers
    (verbatim_rag/index.py:735-738): ``field == "value"`` / ``field == number`` / ``field in [..]``,
    ``metadata["key"] <op> value``, joined by ``and``."""
    if not expr or not expr.strip():
        return lambda row, md: True
    tests = _parse_filter(expr)

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
