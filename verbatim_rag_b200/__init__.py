"""
verbatim_rag_b200 -- the B200-native hot path of KRLabsOrg/verbatim-rag behind the reference's own plugin surfaces.

    from verbatim_rag_b200 import B200SpanExtractor, B200SpladeProvider, B200DenseProvider, B200VectorStore

Drop these into ``VerbatimIndex(vector_store=..., sparse_provider=...)`` / ``VerbatimRAG(index, extractor=...)``
(INTEGRATION.md).  All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
include/vrag_b200.h (libvrag_b200.so, bound with ctypes in ``_native``).  There is no CPU fallback: constructing a
plugin without a CUDA device raises ``_native.NativeError``.
"""
__version__ = "0.1.0"

_LAZY = {
    "B200SpanExtractor": ("extractor", "B200SpanExtractor"),
    "B200SpladeProvider": ("providers", "B200SpladeProvider"),
    "B200DenseProvider": ("providers", "B200DenseProvider"),
    "B200VectorStore": ("vector_store", "B200VectorStore"),
    "ShardedB200VectorStore": ("sharded_store", "ShardedB200VectorStore"),
    "B200Reranker": ("rerank", "B200Reranker"),
    "B200QAExtractor": ("qa_extractor", "B200QAExtractor"),
    "sharded_search_dense": ("distributed", "sharded_search_dense"),
    "sharded_search_sparse": ("distributed", "sharded_search_sparse"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
