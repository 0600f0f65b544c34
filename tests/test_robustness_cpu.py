"""Failure-path behaviour of the host code (round-2 review findings): all-or-nothing inserts, crash consistency of
the on-disk store, the vectorised filter mask, per-chunk error isolation of the extractor.  CPU: native handles are the
oracle-backed fakes of tests/fake_native.py."""
import logging
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture()
def store_cls(monkeypatch):
    import fake_native
    fake_native.install(monkeypatch)
    from verbatim_rag_b200.vector_store import B200VectorStore
    return B200VectorStore


def _rows(n, dim=8, vocab=40, seed=0, start=0):
    rng = np.random.default_rng(seed)
    ids = [f"r{start + i:04d}" for i in range(n)]
    dense = rng.standard_normal((n, dim)).astype(np.float32)
    sparse = [{int(k): float(abs(rng.standard_normal()) + 0.1) for k in rng.choice(vocab, 4, replace=False)} for _ in range(n)]
    metas = [{"year": 2000 + (start + i) % 7, "document_id": f"d{(start + i) % 3}"} for i in range(n)]
    return ids, dense, sparse, [f"t{start + i}" for i in range(n)], metas


def _consistent(store):
    n = len(store._ids)
    assert len(store._dense) == n and len(store._sparse) == n
    assert len(store._texts) == len(store._enh) == len(store._meta) == len(store._promoted) == len(store._alive) == n


def test_invalid_input_is_rejected_before_the_device_is_touched(store_cls):
    store = store_cls(dense_dim=8, sparse_dim=40)
    ids, dense, sparse, texts, metas = _rows(6)
    store.add_vectors(ids, dense.tolist(), sparse, texts, texts, metas)
    bad_sparse = [dict(s) for s in sparse]
    bad_sparse[3][40] = 1.0                                   # term id == sparse_dim
    ids2 = [i + "x" for i in ids]
    for kwargs in (dict(sparse=bad_sparse), dict(texts=texts[:-1]), dict(dense=dense[:, :7]), dict(sparse=sparse[:-1])):
        with pytest.raises(ValueError):
            store.add_vectors(ids2, kwargs.get("dense", dense).tolist(), kwargs.get("sparse", sparse),
                              kwargs.get("texts", texts), texts, metas)
        _consistent(store)
        assert len(store._ids) == 6
    with pytest.raises(ValueError):                           # non-monotone CSR
        store.add_csr(ids2, np.array([0, 4, 2, 6, 8, 10, 12]), np.zeros(12, np.int32), np.ones(12, np.float32), texts, texts,
                      metas, dense=dense)
    _consistent(store)
    res = store.query(dense_query=dense[2].tolist(), top_k=3, search_type="dense")
    assert res[0].id == ids[2] and res[0].text == texts[2]


def test_device_failure_in_the_second_add_keeps_rows_aligned(store_cls):
    store = store_cls(dense_dim=8, sparse_dim=40)
    a = _rows(5)
    store.add_vectors(a[0], a[1].tolist(), a[2], a[3], a[3], a[4])
    real = store._sparse.add_sparse
    calls = {"n": 0}

    def failing(*args):
        calls["n"] += 1
        if calls["n"] == 1:
            raise RuntimeError("cudaMalloc: out of memory")
        return real(*args)

    store._sparse.add_sparse = failing
    b = _rows(4, seed=1, start=5)
    with pytest.raises(RuntimeError):
        store.add_vectors(b[0], b[1].tolist(), b[2], b[3], b[3], b[4])
    _consistent(store)                                          # 9 rows everywhere, the 4 failed ones tombstoned
    assert len(store) == 5 and sum(store._alive) == 5
    c = _rows(3, seed=2, start=9)
    store.add_vectors(c[0], c[1].tolist(), c[2], c[3], c[3], c[4])   # a memory-only store stays usable
    _consistent(store)
    for ids, dense, sparse, texts, _ in (a, c):                 # every hit still maps to its own payload
        for i in range(len(ids)):
            r = store.query(dense_query=dense[i].tolist(), top_k=1, search_type="dense")[0]
            assert (r.id, r.text) == (ids[i], texts[i])
            hit = {r.id: r for r in store.query(sparse_query=sparse[i], top_k=20, search_type="sparse")}[ids[i]]
            assert hit.text == texts[i] and abs(hit.score - sum(v * v for v in sparse[i].values())) < 1e-5
    assert [r.id for r in store.query(dense_query=a[1][0].tolist(), top_k=20, search_type="dense",
                                      filter='document_id == "d1"')] != []


def test_disk_store_orphan_sparse_tail_is_cut_on_the_next_append(store_cls, tmp_path):
    """indices / values are written before indptr: a crash between them leaves bytes no row refers to."""
    path = str(tmp_path / "s")
    store = store_cls(db_path=path, dense_dim=8, sparse_dim=40)
    a = _rows(5)
    store.add_vectors(a[0], a[1].tolist(), a[2], a[3], a[3], a[4])
    with open(os.path.join(path, "sparse.indices.i32"), "ab") as f:      # torn append: indices landed, values half, no indptr
        np.arange(7, dtype=np.int32).tofile(f)
    with open(os.path.join(path, "sparse.values.f32"), "ab") as f:
        np.ones(3, dtype=np.float32).tofile(f)
    with open(os.path.join(path, "tombstones.i64"), "ab") as f:          # tombstone of a row that never became durable
        np.asarray([6], np.int64).tofile(f)
    store2 = store_cls(db_path=path, dense_dim=8, sparse_dim=40)
    assert len(store2) == 5
    b = _rows(4, seed=1, start=5)
    store2.add_vectors(b[0], b[1].tolist(), b[2], b[3], b[3], b[4])
    store3 = store_cls(db_path=path, dense_dim=8, sparse_dim=40)
    assert len(store3) == 9                                              # row 6 is alive: the stale tombstone is gone
    for ids, dense, sparse, texts, _ in (a, b):
        for i in range(len(ids)):
            hit = {r.id: r for r in store3.query(sparse_query=sparse[i], top_k=20, search_type="sparse")}[ids[i]]
            assert hit.text == texts[i] and abs(hit.score - sum(v * v for v in sparse[i].values())) < 1e-5


def test_upsert_tombstones_follow_the_rows_on_disk(store_cls, tmp_path):
    path = str(tmp_path / "s")
    store = store_cls(db_path=path, dense_dim=8, sparse_dim=40)
    a = _rows(4)
    store.add_vectors(a[0], a[1].tolist(), a[2], a[3], a[3], a[4])
    b = _rows(3, seed=3, start=10)
    ids_b = [a[0][1], "dup", "dup"]                       # upsert of an old id + a duplicate inside the batch
    store.add_vectors(ids_b, b[1].tolist(), b[2], ["new1", "dupA", "dupB"], b[3], b[4])
    again = store_cls(db_path=path, dense_dim=8, sparse_dim=40)
    for s in (store, again):
        assert len(s) == 5
        got = {r.id: r.text for r in s.query(dense_query=None, sparse_query=None, top_k=50)}
        assert got[a[0][1]] == "new1" and got["dup"] == "dupB" and len(got) == 5


def test_vectorised_filter_mask_equals_the_row_predicate(store_cls):
    from verbatim_rag_b200.vector_store import _compile_filter
    store = store_cls(dense_dim=8, sparse_dim=40)
    rng = np.random.default_rng(4)
    n = 300
    ids, dense, sparse, texts, _ = _rows(n)
    pool = [2019, 2021.5, "2020", None, True, 2023]
    metas = [{"year": pool[int(rng.integers(len(pool)))], "document_id": f"d{i % 4}", "tag": ["a", "b", None][i % 3]}
             for i in range(n)]
    for m in metas[::17]:
        del m["year"]
    store.add_vectors(ids, dense.tolist(), sparse, texts, texts, metas)
    store.delete(ids[5:9])
    for expr in ('metadata["year"] >= 2020', 'metadata["year"] < 2022 and document_id != "d1"', 'metadata["year"] == "2020"',
                 'document_id in ["d0", "d3"]', 'metadata["tag"] == "a" and metadata["year"] > 2019',
                 'id in ["r0003", "r0100"]', 'metadata["tag"] != "b"', 'metadata["year"] >= "2020"'):
        pred = _compile_filter(expr)
        want = np.array([0 if (store._alive[r] and pred({"id": store._ids[r], **store._promoted[r]}, store._meta[r])) else 1
                         for r in range(n)], np.uint8)
        assert np.array_equal(store._exclude_mask(expr), want), expr
    more = _rows(10, seed=9, start=n)                     # columns extend with the rows
    store.add_vectors(more[0], more[1].tolist(), more[2], more[3], more[3], more[4])
    assert len(store._exclude_mask('metadata["year"] >= 2003')) == n + 10


def test_one_bad_pair_does_not_blank_the_batch(monkeypatch, caplog):
    """Reference convention (extractors.py:225-227): a failing chunk yields [] for that chunk only."""
    import cases
    import fake_native
    fake_native.install(monkeypatch)
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.extractor import B200SpanExtractor
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    monkeypatch.setattr(_native, "spans_from_probs", _host_spans)
    spec = ModernBertSpec(layers=1)
    tok = cases.tokenizer("modernbert")
    ext = B200SpanExtractor(weights=make_modernbert_weights(3, spec), tokenizer=tok, num_layers=1,
                            vocab_size=spec.vocab_size, max_length=64, doc_stride=8, tokenizer_workers=0)
    rng = np.random.default_rng(0)

    class R:
        def __init__(self, t):
            self.text = t
    good_q, long_q = tok.make_question(rng, 10), tok.make_question(rng, 70)      # 70 tokens > max_length - 3
    docs = [R(tok.make_text(rng, 40)) for _ in range(3)]
    with caplog.at_level(logging.ERROR):
        out = ext.extract_spans_batch([good_q, long_q, good_q], [docs, docs[:2], docs[1:]])
    alone = ext.extract_spans(good_q, docs)
    assert out[0] == alone and any(alone.values())          # the good questions are unaffected ...
    assert out[2] == {d.text: alone[d.text] for d in docs[1:]}
    assert out[1] == {d.text: [] for d in docs[:2]}        # ... the unplannable one comes back empty, with a log line
    assert any("cannot be windowed" in r.message for r in caplog.records)
    # a forward failure on one question: the batch is retried per question, only that question is empty
    real = ext._enc.span_forward
    poison = tok.tok.encode(good_q, add_special_tokens=False).ids[0]

    def flaky(ids, cu, want_logits=False):
        if len(cu) > 2 and np.count_nonzero(np.asarray(ids) == poison) > 3:
            raise RuntimeError("injected device failure")
        return real(ids, cu, want_logits)
    other_q = tok.make_question(rng, 12)
    ext._enc.span_forward = flaky
    mixed = ext.extract_spans_batch([good_q, other_q], [docs + [R(tok.make_text(rng, 30))], docs])
    ext._enc.span_forward = real
    assert mixed[1] == ext.extract_spans(other_q, docs)
    assert all(v == [] for v in mixed[0].values())


def _host_spans(probs, tok_cs, tok_ce, ctx_indptr, threshold, min_span_chars, merge_gap_chars):
    """oracle restatement of vrag_spans_from_probs for the CPU tests (the C function needs no GPU, but these tests must
    not depend on the built library)."""
    from oracle.highlighter import spans_from_token_probs
    out = {k: [] for k in ("ctx", "start", "end", "score", "tok_start", "tok_end")}
    for c in range(len(ctx_indptr) - 1):
        a, b = int(ctx_indptr[c]), int(ctx_indptr[c + 1])
        offs = list(zip(tok_cs[a:b].tolist(), tok_ce[a:b].tolist()))
        for sp in spans_from_token_probs("x" * (max(tok_ce[a:b].tolist() + [0])), probs[a:b], offs, threshold,
                                         min_span_chars, merge_gap_chars):
            out["ctx"].append(c)
            for k in ("start", "end", "score", "tok_start", "tok_end"):
                out[k].append(sp[k])
    return {k: np.asarray(v) for k, v in out.items()}


def test_tokenizer_workers_match_in_process_and_survive_a_broken_pool(caplog):
    """The worker pool is an optimisation: same arrays as in-process tokenisation, and a pool that cannot start (e.g. a
    caller script without the ``if __name__ == "__main__"`` guard, which makes ``spawn`` raise) costs ONE warning, after
    which everything is tokenised in-process -- no batch is lost."""
    from verbatim_rag_b200._tokworker import TokenizerWorkers, encode_with
    from verbatim_rag_b200.synthetic import SyntheticTokenizer
    tk = SyntheticTokenizer("modernbert")
    rng = np.random.default_rng(3)
    texts = [tk.make_text(rng, int(n)) for n in rng.integers(1, 60, size=23)]
    ref = encode_with(tk.tok, texts)
    w = TokenizerWorkers(tk.tok, 2)
    try:
        got = w.encode(texts)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
    finally:
        w.close()
    broken = TokenizerWorkers(tk.tok, 2)

    def boom():
        raise RuntimeError("An attempt has been made to start a new process before the current process has finished "
                           "its bootstrapping phase.")
    broken._ensure = boom
    with caplog.at_level(logging.WARNING):
        got = broken.encode(texts)
        again = broken.encode(texts)
    for a, b, c in zip(got, again, ref):
        assert np.array_equal(a, c) and np.array_equal(b, c)
    assert broken.n == 0
    assert sum("tokenising in-process" in r.message for r in caplog.records) == 1
