"""
The reference's plugin interfaces for the hot path.

When the reference packages are importable (``verbatim_rag`` / ``verbatim_core`` on sys.path -- e.g. the
user has ``pip install verbatim-rag``), the B200 plugins subclass the reference's OWN ABCs, so
``isinstance`` checks and ``VerbatimIndex`` / ``VerbatimRAG`` accept them unchanged.  When they are not
(the GPU box has no /root/reference), structurally identical mirrors are defined here so the plugins still
load and the C-ABI tests run.  Signatures mirror:

* ``SpanExtractor``            packages/core/verbatim_core/extractors.py:34-54
* ``DenseEmbeddingProvider``   verbatim_rag/embedding_providers.py:14-30
* ``SparseEmbeddingProvider``  verbatim_rag/embedding_providers.py:33-49
* ``VectorStore``/``SearchResult``  verbatim_rag/vector_stores/base.py:10-74
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

USING_REFERENCE_ABCS = {"extractor": False, "providers": False, "vector_store": False}

try:  # reference's own classes
    from verbatim_core.extractors import SpanExtractor  # type: ignore
    USING_REFERENCE_ABCS["extractor"] = True
except Exception:  # pragma: no cover - exercised on machines without the reference

    class SpanExtractor(ABC):  # type: ignore[no-redef]
        @abstractmethod
        def extract_spans(self, question: str, search_results: List[Any]) -> Dict[str, List[str]]:
            raise NotImplementedError

        async def extract_spans_async(self, question: str, search_results: List[Any]) -> Dict[str, List[str]]:
            import asyncio

            return await asyncio.to_thread(self.extract_spans, question, search_results)


try:
    from verbatim_rag.embedding_providers import DenseEmbeddingProvider, SparseEmbeddingProvider  # type: ignore
    USING_REFERENCE_ABCS["providers"] = True
except Exception:  # pragma: no cover

    class DenseEmbeddingProvider(ABC):  # type: ignore[no-redef]
        @abstractmethod
        def embed_text(self, text: str) -> List[float]: ...

        @abstractmethod
        def embed_batch(self, texts: List[str]) -> List[List[float]]: ...

        @abstractmethod
        def get_dimension(self) -> int: ...

    class SparseEmbeddingProvider(ABC):  # type: ignore[no-redef]
        @abstractmethod
        def embed_text(self, text: str) -> Dict[int, float]: ...

        @abstractmethod
        def embed_batch(self, texts: List[str]) -> List[Dict[int, float]]: ...

        @abstractmethod
        def get_dimension(self) -> int: ...


try:
    from verbatim_rag.vector_stores.base import SearchResult, VectorStore  # type: ignore
    USING_REFERENCE_ABCS["vector_store"] = True
except Exception:  # pragma: no cover

    @dataclass
    class SearchResult:  # type: ignore[no-redef]
        id: str
        score: float
        metadata: Dict[str, Any]
        text: str
        enhanced_text: str = ""

        def __gt__(self, other):
            return self.score > other.score

        def __lt__(self, other):
            return self.score < other.score

        def __eq__(self, other):
            return self.score == other.score

        def __hash__(self):
            return hash((self.id, self.score, self.text, self.enhanced_text))

    class VectorStore(ABC):  # type: ignore[no-redef]
        @abstractmethod
        def add_vectors(self, ids, dense_vectors, sparse_vectors, texts, enhanced_texts, metadatas): ...

        @abstractmethod
        def query(self, dense_query=None, sparse_query=None, text_query=None, top_k: int = 5,
                  search_type: str = "hybrid", filter: Optional[str] = None) -> List["SearchResult"]: ...

        @abstractmethod
        def delete(self, ids: List[str]): ...


try:
    from verbatim_rag.rerankers import BaseReranker  # type: ignore
    USING_REFERENCE_ABCS["reranker"] = True
except Exception:  # pragma: no cover
    USING_REFERENCE_ABCS["reranker"] = False

    class BaseReranker(ABC):  # type: ignore[no-redef]
        """Structural mirror of verbatim_rag/rerankers.py:14-41 (Reranker + BaseReranker)."""

        def __init__(self, rerank_k: int = 50, text_field: str = "text"):
            self.rerank_k = rerank_k
            self.text_field = text_field

        @abstractmethod
        def rerank(self, question: str, results: List[Any]) -> List[Any]: ...

        async def rerank_async(self, question: str, results: List[Any]) -> List[Any]:
            import asyncio
            return await asyncio.to_thread(self.rerank, question, results)

        def _split_results(self, results):
            return results[: self.rerank_k], results[self.rerank_k:]

        def _get_texts(self, results) -> List[str]:
            if self.text_field == "enhanced_text":
                return [r.enhanced_text or r.text for r in results]
            return [r.text for r in results]


try:  # the reference's own hybrid merge / metadata helpers are reused verbatim when present
    from verbatim_rag.vector_stores.hybrid_search import merge_hybrid_results, sanitize_hybrid_weights  # type: ignore
    from verbatim_rag.vector_stores.utils import json_serialize_safe, promote_metadata  # type: ignore
    USING_REFERENCE_ABCS["hybrid"] = True
except Exception:  # pragma: no cover
    USING_REFERENCE_ABCS["hybrid"] = False

    def sanitize_hybrid_weights(hybrid_weights: Dict[str, float]) -> Dict[str, float]:
        if not hybrid_weights:
            raise ValueError("hybrid_weights must be a non-empty dict")
        ok = {m: float(w) for m, w in hybrid_weights.items()
              if m in ("dense", "sparse", "full_text") and isinstance(w, (int, float)) and w > 0}
        if not ok:
            raise ValueError("No valid hybrid_weights after validation")
        return ok

    def merge_hybrid_results(results_by_method, top_k, weights, rrf_k: int = 60, log_label: str = ""):
        """Weighted reciprocal-rank fusion: score(id) = sum_m w_m / (rrf_k + rank_m + 1), weights normalised to 1,
        distance = 1 - score (behaviour of verbatim_rag/vector_stores/hybrid_search.py:73-129)."""
        avail = {m: weights.get(m, 0.0) for m in results_by_method}
        tot = sum(avail.values())
        norm = ({m: 1.0 / len(avail) for m in avail} if tot == 0 else {m: v / tot for m, v in avail.items()})
        score, first = {}, {}
        for m, hits in results_by_method.items():
            for rank, hit in enumerate(hits):
                hid = hit.get("id")
                if not hid:
                    continue
                if hid not in score:
                    score[hid] = 0.0
                    first[hid] = hit
                score[hid] += norm.get(m, 0.0) * (1.0 / (rrf_k + rank + 1))
        out = []
        for hid in sorted(score, key=lambda i: score[i], reverse=True)[:top_k]:
            h = dict(first[hid])
            h["distance"] = 1.0 - score[hid]
            out.append(h)
        return out

    def json_serialize_safe(obj):
        from datetime import datetime
        from enum import Enum
        if isinstance(obj, datetime):
            return obj.isoformat()
        if isinstance(obj, Enum):
            return getattr(obj, "value", str(obj))
        if isinstance(obj, dict):
            return {str(k): json_serialize_safe(v) for k, v in obj.items()}
        if isinstance(obj, list):
            return [json_serialize_safe(v) for v in obj]
        return obj

    def promote_metadata(metadata):
        md = dict(metadata or {})
        promoted = {k: md.pop(k) for k in list(md) if k in ("user_id", "document_id", "dataset_id")}
        return promoted, md
