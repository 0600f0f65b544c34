"""
Batched entry points of the callers either side of the hot path.

The reference's orchestration is one query at a time: ``VerbatimIndex.query`` (verbatim_rag/index.py:552-655) embeds
one text and issues one vector-store call, ``VerbatimRAG.query`` (verbatim_rag/core.py:210-277) retrieves, extracts
spans and fills a template for one question.  On a GPU that leaves every kernel launch-bound.  The functions here
run the SAME steps for a list of queries with GPU-sized batches:

* ``index_query_batch(index, texts, k, ...)``  == ``[index.query(t, k, ...) for t in texts]``
* ``rag_query_batch(rag, questions, ...)``     == ``[rag.query(q, ...) for q in questions]``

They take the reference's own ``VerbatimIndex`` / ``VerbatimRAG`` objects (nothing is subclassed or patched) and use a
plugin's batch method when it has one -- ``embed_batch`` / ``embed_batch_csr``, ``B200VectorStore.query_batch``,
``B200SpanExtractor.extract_spans_batch`` -- and the reference's per-item method otherwise, so the result is the same
list whichever plugins are installed.  ``tests/test_reference_conformance.py`` checks the equality against the real
reference classes.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence


def _embed_queries_sparse(provider, texts: Sequence[str]) -> List[Dict[int, float]]:
    """Query-side SPLADE vectors.  The reference embeds queries with ``embed_text`` whose filter is ``abs(w) > 1e-6``
    (embedding_providers.py:138-146) -- ``embed_batch`` uses ``!= 0`` (:161-163) -- so the batch path asks for the
    query filter explicitly when the provider can take it."""
    if hasattr(provider, "embed_batch_csr"):
        ip, idx, val = provider.embed_batch_csr(list(texts), min_abs=1e-6)
        idx_l, val_l = idx.tolist(), val.tolist()
        return [dict(zip(idx_l[ip[i]:ip[i + 1]], val_l[ip[i]:ip[i + 1]])) for i in range(len(texts))]
    return [provider.embed_text(t) for t in texts]


def _embed_queries_dense(provider, texts: Sequence[str]):
    if hasattr(provider, "embed_array"):
        return provider.embed_array(list(texts))
    return [provider.embed_text(t) for t in texts]


def index_query_batch(
    index,
    texts: Sequence[Optional[str]],
    k: int = 5,
    search_type: str = "auto",
    filter: Optional[str] = None,
    search_params: Optional[Dict[str, Any]] = None,
    hybrid_weights: Optional[Dict[str, float]] = None,
    rrf_k: int = 60,
) -> List[List[Any]]:
    """Batched ``VerbatimIndex.query`` (index.py:552-655): same branch order, one embedding pass and one vector-store
    pass for all texts."""
    store = index.vector_store
    dense_p = getattr(index, "dense_provider", None)
    sparse_p = getattr(index, "sparse_provider", None)
    texts = list(texts)
    out: List[Optional[List[Any]]] = [None] * len(texts)
    live = [i for i, t in enumerate(texts) if t]
    for i, t in enumerate(texts):
        if not t:   # 1. filter-only query
            out[i] = store.query(dense_query=None, sparse_query=None, text_query=None, top_k=k, filter=filter,
                                 search_params=search_params)
    if not live:
        return out  # type: ignore[return-value]
    q_texts = [texts[i] for i in live]

    want_dense = want_sparse = False
    if hybrid_weights is not None:   # 2. hybrid_weights drive everything
        want_dense = "dense" in hybrid_weights and dense_p is not None
        want_sparse = "sparse" in hybrid_weights and sparse_p is not None
    else:
        if search_type == "auto":    # 3. resolve
            if dense_p and sparse_p:
                search_type = "hybrid"
            elif dense_p:
                search_type = "dense"
            elif sparse_p:
                search_type = "sparse"
            elif getattr(store, "enable_full_text", False):
                search_type = "full_text"
            else:
                raise ValueError("No search method available")
        if search_type == "full_text":   # 4. no embeddings
            for i in live:
                out[i] = store.query(dense_query=None, sparse_query=None, text_query=texts[i], top_k=k,
                                     search_type="full_text", filter=filter, search_params=search_params)
            return out  # type: ignore[return-value]
        want_dense = search_type in ("dense", "hybrid") and dense_p is not None
        want_sparse = search_type in ("sparse", "hybrid") and sparse_p is not None

    dq = _embed_queries_dense(dense_p, q_texts) if want_dense else None
    sq = _embed_queries_sparse(sparse_p, q_texts) if want_sparse else None

    if hasattr(store, "query_batch"):
        res = store.query_batch(dense_queries=dq, sparse_queries=sq, text_queries=q_texts, top_k=k,
                                search_type=search_type, filter=filter, search_params=search_params,
                                hybrid_weights=hybrid_weights, rrf_k=rrf_k)
    else:
        res = []
        for j, t in enumerate(q_texts):
            kw = dict(dense_query=(list(dq[j]) if dq is not None else None),
                      sparse_query=(sq[j] if sq is not None else None), text_query=t, top_k=k, filter=filter,
                      search_params=search_params, rrf_k=rrf_k)
            if hybrid_weights is not None:
                kw["hybrid_weights"] = hybrid_weights
            else:
                kw["search_type"] = search_type
            res.append(store.query(**kw))
    for i, r in zip(live, res):
        out[i] = r
    return out  # type: ignore[return-value]


def _rag_route(rag, questions: Sequence[str]):
    """step 0 of VerbatimRAG.query (core.py:210-230): optional intent detection; -> (responses, indices still to do)"""
    responses: List[Any] = [None] * len(questions)
    todo: List[int] = []
    for i, q in enumerate(questions):
        decision = rag._detect_intent(q)
        route = rag._decision_field(decision, "route")
        if decision and route and route != "continue":
            answer = rag._decision_field(decision, "answer", "") or ""
            responses[i] = rag._build_short_circuit_response(q, answer)
        else:
            todo.append(i)
    return responses, todo


def _rag_finish(rag, questions: Sequence[str], found: Sequence[List[Any]]):
    """steps 2-5 of VerbatimRAG.query (core.py:240-277) for questions whose retrieval is done: rerank, span extraction
    (one batched call), template, response.  -> [(response, search_results)]"""
    results = [rag._apply_reranker(q, r) for q, r in zip(questions, found)]
    structured = rag.template_manager.current_mode == "structured"
    spans: List[Dict[str, List[str]]] = []
    if not structured and questions:
        if hasattr(rag.extractor, "extract_spans_batch"):
            spans = rag.extractor.extract_spans_batch(list(questions), results)
        else:
            spans = [rag.extractor.extract_spans(q, r) for q, r in zip(questions, results)]
    out = []
    for j, q in enumerate(questions):
        if structured:
            answer, all_spans = rag._process_structured(q, results[j])
        else:
            all_spans = spans[j]
            display_spans, citation_spans = rag._rank_and_split_spans(all_spans)
            answer = rag.template_manager.process(q, display_spans, citation_spans)
        answer = rag.response_builder.clean_answer(answer)
        out.append((rag.response_builder.build_response(question=q, answer=answer, search_results=results[j],
                                                        relevant_spans=all_spans, display_span_count=len(all_spans)),
                    results[j]))
    return out


def rag_query_batch(
    rag,
    questions: Sequence[str],
    k: Optional[int] = None,
    filter: Optional[str] = None,
    hybrid_weights: Optional[Dict[str, float]] = None,
    rrf_k: int = 60,
    search_params: Optional[Dict[str, Any]] = None,
    return_search_results: bool = False,
    group=None,
) -> List[Any]:
    """Batched ``VerbatimRAG.query`` (core.py:210-277): intent routing and templates stay per question (host string
    work), retrieval and span extraction run once for the whole batch.

    Under ``torch.distributed`` (one process per GPU, every rank calling this with the same questions; BASELINE
    configs[4]) the retrieval is the sharded store's collective search over all questions, then the questions are split
    contiguously over the ranks for span extraction + response building (independent units, no collective on that
    path) and the responses are exchanged with one ``all_gather_object``: every rank returns the full list."""
    import torch.distributed as dist
    from .distributed import shard_bounds
    questions = list(questions)
    n = len(questions)
    responses, todo = _rag_route(rag, questions)
    results: List[List[Any]] = [[] for _ in range(n)]
    if todo:
        kk = k or rag.k
        found = index_query_batch(rag.index, [questions[i] for i in todo], k=kk, filter=filter,
                                  hybrid_weights=hybrid_weights, rrf_k=rrf_k, search_params=search_params)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            lo, hi = shard_bounds(len(todo), dist.get_rank(group), world)
            mine = _rag_finish(rag, [questions[i] for i in todo[lo:hi]], found[lo:hi])
            parts: List[Any] = [None] * world
            dist.all_gather_object(parts, mine, group=group)
            done = [x for part in parts for x in part]
        else:
            done = _rag_finish(rag, [questions[i] for i in todo], found)
        for i, (resp, res) in zip(todo, done):
            responses[i] = resp
            results[i] = res
    if return_search_results:
        return [(responses[i], results[i]) for i in range(n)]
    return responses
