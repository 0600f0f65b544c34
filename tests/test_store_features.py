"""Host logic of B200VectorStore beyond the single-query path: metadata-filter pushdown,
batched queries, the append-only on-disk store.  CPU: the native handles are the oracle-backed fakes of
tests/fake_native.py (the GPU versions of the same checks live in tests/test_gpu_golden.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


@pytest.fixture()
def store_cls(monkeypatch):
    import fake_native
    fake_native.install(monkeypatch)
    from verbatim_rag_b200.vector_store import B200VectorStore
    return B200VectorStore


def _corpus(n=60, dim=16, vocab=50, seed=0):
    rng = np.random.default_rng(seed)
    dense = rng.standard_normal((n, dim)).astype(np.float32)
    sparse = []
    for _ in range(n):
        ks = rng.choice(vocab, size=int(rng.integers(3, 9)), replace=False)
        sparse.append({int(k): float(abs(rng.standard_normal()) + 0.05) for k in ks})
    ids = [f"c{i:04d}" for i in range(n)]
    texts = [f"text {i}\nline two" for i in range(n)]
    metas = [{"year": 2000 + i % 30, "document_id": f"doc{i % 5}", "tag": "a" if i % 3 else "b"} for i in range(n)]
    return ids, dense, sparse, texts, metas


def _fill(store, c, a=0, b=None):
    ids, dense, sparse, texts, metas = c
    b = len(ids) if b is None else b
    store.add_vectors(ids[a:b], dense[a:b].tolist() if store.enable_dense else None,
                      sparse[a:b] if store.enable_sparse else None, texts[a:b], [t + " +" for t in texts[a:b]],
                      metas[a:b])


def _brute_dense(c, q, k, keep):
    ids, dense, _, _, _ = c
    d = dense.astype(np.float64)
    sc = d @ q.astype(np.float64) / (np.linalg.norm(d, axis=1) * np.linalg.norm(q.astype(np.float64)))
    order = [i for i in np.lexsort((np.arange(len(ids)), -sc)) if keep(i)][:k]
    return [ids[i] for i in order], [sc[i] for i in order]


@pytest.mark.parametrize("expr,keep", [
    ('metadata["year"] >= 2020', lambda m: m["year"] >= 2020),
    ('document_id == "doc3"', lambda m: m["document_id"] == "doc3"),
    ('metadata["tag"] == "b" and metadata["year"] < 2010', lambda m: m["tag"] == "b" and m["year"] < 2010),
    ('document_id in ["doc1", "doc4"]', lambda m: m["document_id"] in ("doc1", "doc4")),
])
def test_filter_pushdown_equals_post_filtered_exact_search(store_cls, expr, keep):
    c = _corpus()
    store = store_cls(dense_dim=16, sparse_dim=50)
    _fill(store, c)
    rng = np.random.default_rng(9)
    q = rng.standard_normal(16).astype(np.float32)
    got = store.query(dense_query=q.tolist(), top_k=7, search_type="dense", filter=expr)
    exp_ids, exp_sc = _brute_dense(c, q, 7, lambda i: keep(c[4][i]))
    assert [r.id for r in got] == exp_ids
    assert np.allclose([r.score for r in got], exp_sc, atol=1e-6)
    # the filter is gone afterwards: an unfiltered query sees every row again
    free = store.query(dense_query=q.tolist(), top_k=7, search_type="dense")
    assert [r.id for r in free] == _brute_dense(c, q, 7, lambda i: True)[0]
    # sparse branch honours the same mask
    sq = c[2][5]
    sp = store.query(sparse_query=sq, top_k=60, search_type="sparse", filter=expr)
    assert sp and all(keep(c[4][int(r.id[1:])]) for r in sp)
    # deletes compose with the filter
    victim = exp_ids[0]
    store.delete([victim])
    again = store.query(dense_query=q.tolist(), top_k=7, search_type="dense", filter=expr)
    assert victim not in [r.id for r in again]
    assert [r.id for r in again][:len(exp_ids) - 1] == exp_ids[1:]


def test_unsupported_filter_raises_and_leaves_no_mask(store_cls):
    c = _corpus(20)
    store = store_cls(dense_dim=16, sparse_dim=50)
    _fill(store, c)
    q = np.ones(16, np.float32).tolist()
    with pytest.raises(ValueError):
        store.query(dense_query=q, top_k=3, search_type="dense", filter="year LIKE 'x%'")
    assert len(store.query(dense_query=q, top_k=3, search_type="dense")) == 3


@pytest.mark.parametrize("mode", ["dense", "sparse", "hybrid", "weights", "weights_one"])
def test_query_batch_equals_single_queries(store_cls, mode):
    c = _corpus()
    store = store_cls(dense_dim=16, sparse_dim=50)
    _fill(store, c)
    rng = np.random.default_rng(4)
    dq = rng.standard_normal((9, 16)).astype(np.float32)
    sq = [c[2][i] for i in range(9)]
    kw = {"top_k": 5, "filter": 'metadata["year"] >= 2005'}
    if mode == "dense":
        batch = store.query_batch(dense_queries=dq, search_type="dense", **kw)
        single = [store.query(dense_query=dq[i].tolist(), search_type="dense", **kw) for i in range(9)]
    elif mode == "sparse":
        batch = store.query_batch(sparse_queries=sq, search_type="sparse", **kw)
        single = [store.query(sparse_query=sq[i], search_type="sparse", **kw) for i in range(9)]
    elif mode == "hybrid":
        batch = store.query_batch(dense_queries=dq, sparse_queries=sq, search_type="hybrid", rrf_k=30, **kw)
        single = [store.query(dense_query=dq[i].tolist(), sparse_query=sq[i], search_type="hybrid", rrf_k=30, **kw)
                  for i in range(9)]
    else:
        w = {"dense": 0.7, "sparse": 0.3} if mode == "weights" else {"sparse": 1.0}
        batch = store.query_batch(dense_queries=dq, sparse_queries=sq, hybrid_weights=w, **kw)
        single = [store.query(dense_query=dq[i].tolist(), sparse_query=sq[i], hybrid_weights=w, **kw) for i in range(9)]
    assert len(batch) == 9
    for b, s in zip(batch, single):
        assert [r.id for r in b] == [r.id for r in s]
        assert np.allclose([r.score for r in b], [r.score for r in s], atol=1e-7)
        assert [r.metadata for r in b] == [r.metadata for r in s]


def test_on_disk_store_round_trip(store_cls, tmp_path):
    c = _corpus()
    path = str(tmp_path / "collection")
    store = store_cls(db_path=path, dense_dim=16, sparse_dim=50)
    _fill(store, c, 0, 40)
    _fill(store, c, 40, 60)
    store.delete(["c0003", "c0041"])
    # upsert: same primary key again, new vector + text
    ids, dense, sparse, texts, metas = c
    store.add_vectors(["c0007"], [(-dense[7]).tolist()], [sparse[8]], ["replaced"], ["replaced +"], [{"year": 1999}])
    store.add_documents([{"id": "doc1", "title": "T", "source": "s", "raw_content": "raw", "metadata": {"a": 1}}])
    q = dense[7] * 0.9 + 0.01
    before_d = store.query(dense_query=q.tolist(), top_k=10, search_type="dense")
    before_s = store.query(sparse_query=sparse[41], top_k=10, search_type="sparse")
    before_browse = store.query(top_k=100)
    assert sorted(os.listdir(path)) == ["dense.f32", "documents.jsonl", "manifest.json", "payload.jsonl",
                                        "sparse.indices.i32", "sparse.indptr.i64", "sparse.values.f32",
                                        "tombstones.i64"]
    assert os.path.getsize(os.path.join(path, "dense.f32")) == 61 * 16 * 4   # raw row-major fp32, append-only

    again = store_cls(db_path=path, dense_dim=16, sparse_dim=50)
    assert len(again) == len(store) == 58
    for a, b in ((before_d, again.query(dense_query=q.tolist(), top_k=10, search_type="dense")),
                 (before_s, again.query(sparse_query=sparse[41], top_k=10, search_type="sparse")),
                 (before_browse, again.query(top_k=100))):
        assert [(r.id, r.text, r.enhanced_text, r.metadata) for r in a] == \
               [(r.id, r.text, r.enhanced_text, r.metadata) for r in b]
        assert np.allclose([r.score for r in a], [r.score for r in b], atol=0)
    assert "c0003" not in [r.id for r in again.query(top_k=100)]
    assert [r.text for r in again.query(top_k=100) if r.id == "c0007"] == ["replaced"]
    assert again.get_document("doc1")["title"] == "T"
    # the reopened store keeps appending
    again.add_vectors(["new"], [dense[0].tolist()], [sparse[0]], ["n"], ["n +"], [{}])
    third = store_cls(db_path=path, dense_dim=16, sparse_dim=50)
    assert len(third) == 59 and third.query(dense_query=dense[0].tolist(), top_k=2, search_type="dense")[0].score > 0.999

    with pytest.raises(ValueError):
        store_cls(db_path=path, dense_dim=32, sparse_dim=50)   # a different schema must not silently reuse the files


def test_on_disk_store_survives_a_torn_append(store_cls, tmp_path):
    c = _corpus(30)
    path = str(tmp_path / "torn")
    store = store_cls(db_path=path, dense_dim=16, sparse_dim=50)
    _fill(store, c, 0, 20)
    # simulate a crash in the middle of the next append: vectors on disk, payload line incomplete
    with open(os.path.join(path, "dense.f32"), "ab") as f:
        c[1][20:25].tofile(f)
    with open(os.path.join(path, "payload.jsonl"), "a") as f:
        f.write('{"id": "c0020", "text": "tr')
    again = store_cls(db_path=path, dense_dim=16, sparse_dim=50)
    assert len(again) == 20
    _fill(again, c, 20, 30)
    final = store_cls(db_path=path, dense_dim=16, sparse_dim=50)
    assert len(final) == 30
    q = c[1][25]
    assert final.query(dense_query=q.tolist(), top_k=1, search_type="dense")[0].id == "c0025"
