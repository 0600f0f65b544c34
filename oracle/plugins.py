"""
Reference-shaped CPU plugins built on the oracle.  TEST / CPU-BASELINE INFRASTRUCTURE ONLY.

These restate, on the CPU and with the reference's own control flow, what the reference's plugins do once their
third-party dependencies are replaced by the oracle forwards:

* ``OracleSpanExtractor``  -- ModelSpanExtractor._extract_highlighter (extractors.py:203-228): ONE batch-1 forward per
  chunk, sequentially, then the ``process`` post-processing (oracle/highlighter.py);
* ``OracleSpladeProvider`` -- SpladeProvider (embedding_providers.py:117-169): batches of 32 sorted by length, dense
  ``[N, V]`` intermediate, then the reference's own ``np.nonzero`` / ``abs > 1e-6`` dict loops;
* ``OracleFlatStore``      -- LocalMilvusStore/BaseMilvusStore.query (milvus_base.py:189-313): exact scan per query.

They are what ``bench.py`` times as ``cpu_baseline`` / ``--impl reference`` and what
tests/test_reference_conformance.py drives through the reference's real VerbatimIndex / VerbatimRAG.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from verbatim_rag_b200.interfaces import (SearchResult, SpanExtractor, SparseEmbeddingProvider, VectorStore,
                                          merge_hybrid_results, promote_metadata)

from . import bert_splade, flat_topk, highlighter, modernbert


class OracleSpanExtractor(SpanExtractor):
    def __init__(self, weights, tokenizer, spec=None, threshold: float = 0.2, min_span_chars: int = 30,
                 merge_gap_chars: int = 20, max_length: int = 8192, doc_stride: int = 256):
        self.weights, self.tokenizer, self.spec = weights, tokenizer, spec
        self.threshold, self.min_span_chars, self.merge_gap_chars = threshold, min_span_chars, merge_gap_chars
        self.max_length, self.doc_stride = max_length, doc_stride

    def _forward(self, seqs):
        return modernbert.modernbert_forward_varlen(self.weights, seqs, self.spec, batch=1)

    def process(self, question: str, context: str) -> Dict[str, Any]:
        return highlighter.process(question, context, tokenizer=self.tokenizer, forward=self._forward,
                                   threshold=self.threshold, min_span_chars=self.min_span_chars,
                                   merge_gap_chars=self.merge_gap_chars, max_length=self.max_length,
                                   doc_stride=self.doc_stride)

    def extract_spans(self, question: str, search_results: List[Any]) -> Dict[str, List[str]]:
        relevant: Dict[str, List[str]] = {}
        for result in search_results:           # sequential, batch size 1: extractors.py:207
            context = getattr(result, "text", "")
            if not context.strip():
                relevant[context] = []
                continue
            out = self.process(question, context)
            relevant[context] = [sp["text"] for sp in out.get("spans", []) if sp.get("text", "").strip()]
        return relevant


class OracleSpladeProvider(SparseEmbeddingProvider):
    def __init__(self, weights, tokenizer, spec=None, max_seq_length: int = 512):
        self.weights, self.tokenizer, self.spec, self.max_seq_length = weights, tokenizer, spec, max_seq_length

    def tokenize(self, texts):
        tk = self.tokenizer
        enc = tk.tok.encode_batch(list(texts), add_special_tokens=False)
        body = self.max_seq_length - 2
        return [np.asarray([tk.cls_id] + list(e.ids[:body]) + [tk.sep_id], dtype=np.int64) for e in enc]

    def encode_dense(self, texts) -> np.ndarray:
        return bert_splade.splade_encode(self.weights, self.tokenize(texts), self.spec, batch_size=32)

    def embed_text(self, text: str) -> Dict[int, float]:
        return bert_splade.to_dict_embed_text(self.encode_dense([text])[0])

    def embed_batch(self, texts: List[str]) -> List[Dict[int, float]]:
        return bert_splade.to_dicts_embed_batch(self.encode_dense(texts))

    def get_dimension(self) -> int:
        return 30522 if self.spec is None else self.spec.vocab_size


class OracleFlatStore(VectorStore):
    """Exact scan, one call per query, like the single-query Milvus RPC of milvus_base.py:239-259."""

    def __init__(self, dense_dim: int = 768, enable_dense: bool = True, enable_sparse: bool = True,
                 sparse_dim: int = 30522):
        self.enable_dense, self.enable_sparse, self.enable_full_text = enable_dense, enable_sparse, False
        self.dense_dim, self.sparse_dim = dense_dim, sparse_dim
        self.ids: List[str] = []
        self.texts: List[str] = []
        self.enh: List[str] = []
        self.meta: List[Dict[str, Any]] = []
        self.dense = np.zeros((0, dense_dim), np.float32)
        self.sparse_rows: List[Dict[int, float]] = []
        self._csr = None
        self.documents: Dict[str, Dict[str, Any]] = {}

    def add_vectors(self, ids, dense_vectors, sparse_vectors, texts, enhanced_texts, metadatas):
        if self.enable_dense and (dense_vectors is None or len(dense_vectors) == 0):
            raise ValueError("Dense vectors required but not provided")
        if self.enable_sparse and (sparse_vectors is None or len(sparse_vectors) == 0):
            raise ValueError("Sparse vectors required but not provided")
        self.ids += list(ids)
        self.texts += list(texts)
        self.enh += list(enhanced_texts)
        for m in metadatas:
            promoted, cleaned = promote_metadata(m)
            self.meta.append({**cleaned, **promoted})
        if self.enable_dense:
            self.dense = np.concatenate([self.dense, np.asarray(dense_vectors, np.float32)], axis=0)
        if self.enable_sparse:
            self.sparse_rows += list(sparse_vectors)
            self._csr = None

    def add_documents(self, documents):
        for d in documents or []:
            self.documents[d.get("id", "")] = d

    def get_document(self, document_id):
        return self.documents.get(document_id)

    def _hits(self, ids, scores, drop_zero):
        return [{"id": self.ids[i], "distance": float(s),
                 "entity": {"text": self.texts[i], "enhanced_text": self.enh[i], "metadata": dict(self.meta[i])}}
                for i, s in zip(ids.tolist(), scores.tolist()) if not (drop_zero and s == 0.0)]

    def _dense_hits(self, q, limit):
        ids, sc = flat_topk.dense_cosine_topk(self.dense, np.asarray(q, np.float32)[None], limit)
        return self._hits(ids[0], sc[0], False)

    def _sparse_hits(self, q, limit):
        if self._csr is None:
            self._csr = flat_topk.dicts_to_csr(self.sparse_rows)
        ids, sc = flat_topk.sparse_ip_topk(*self._csr, self.sparse_dim, [q], limit)
        return self._hits(ids[0], sc[0], True)

    def query(self, dense_query=None, sparse_query=None, text_query=None, top_k: int = 5, search_type: str = "hybrid",
              filter: Optional[str] = None, search_params=None, hybrid_weights=None, rrf_k: int = 60):
        if not dense_query and not sparse_query:
            return [SearchResult(id=i, score=1.0, text=t, enhanced_text=e, metadata=dict(m))
                    for i, t, e, m in list(zip(self.ids, self.texts, self.enh, self.meta))[:top_k]]
        if search_type == "dense" and dense_query:
            hits = self._dense_hits(dense_query, top_k)
        elif search_type == "sparse" and sparse_query:
            hits = self._sparse_hits(sparse_query, top_k)
        elif search_type == "hybrid" and dense_query and sparse_query:
            hits = merge_hybrid_results({"dense": self._dense_hits(dense_query, top_k * 2),
                                         "sparse": self._sparse_hits(sparse_query, top_k * 2)},
                                        top_k, {"dense": 0.5, "sparse": 0.5}, rrf_k=rrf_k)
        else:
            raise ValueError(f"Invalid search configuration: type={search_type}")
        return [SearchResult(id=h["id"], score=h["distance"], text=h["entity"]["text"],
                             enhanced_text=h["entity"]["enhanced_text"], metadata=h["entity"]["metadata"]) for h in hits]

    def delete(self, ids):
        raise NotImplementedError("oracle store is append-only")
