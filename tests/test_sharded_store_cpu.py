"""N>1 host logic of BASELINE configs[4] on CPU (world_size 2, gloo): ShardedB200VectorStore (vectors split over the
ranks, payload replicated, one all-gather per search) must answer exactly like the unsharded B200VectorStore, and the
data-parallel ``rag_query_batch`` must return the same responses on every rank as the single-process call.  Native
handles are the oracle-backed fakes of tests/fake_native.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _case():
    rng = np.random.default_rng(3)
    n, dim, vocab = 53, 16, 60
    ids = [f"c{i:03d}" for i in range(n)]
    dense = rng.standard_normal((n, dim)).astype(np.float32)
    dense[40] = dense[2]                                                     # an exact tie across the shards
    sparse = [{int(t): float(abs(rng.standard_normal()) + 0.05) for t in rng.choice(vocab, 5, replace=False)}
              for _ in range(n)]
    texts = [f"chunk {i}" for i in range(n)]
    metas = [{"year": 2000 + i % 9, "document_id": f"d{i % 4}"} for i in range(n)]
    dq = rng.standard_normal((7, dim)).astype(np.float32)
    sq = [{int(t): 1.0 + 0.1 * j for j, t in enumerate(rng.choice(vocab, 6, replace=False))} for _ in range(7)]
    return ids, dense, sparse, texts, metas, dq, sq


def _answers(store):
    ids, dense, sparse, texts, metas, dq, sq = _case()
    for a, b in ((0, 20), (20, 21), (21, 53)):                              # several insert batches, one of a single row
        store.add_vectors(ids[a:b], dense[a:b].tolist(), sparse[a:b], texts[a:b], texts[a:b], metas[a:b])
    store.delete([ids[5], ids[33]])
    store.add_vectors([ids[7]], dense[8:9].tolist(), sparse[8:9], ["upserted"], ["upserted"], [metas[7]])   # upsert
    out = {}
    out["dense"] = [[(r.id, r.score, r.text) for r in rs] for rs in store.query_batch(dense_queries=dq, top_k=6, search_type="dense")]
    out["sparse"] = [[(r.id, r.score) for r in rs] for rs in store.query_batch(sparse_queries=sq, top_k=6, search_type="sparse")]
    out["hybrid"] = [[(r.id, round(r.score, 12)) for r in rs]
                     for rs in store.query_batch(dense_queries=dq, sparse_queries=sq, top_k=5, search_type="hybrid")]
    out["filtered"] = [[r.id for r in rs] for rs in store.query_batch(dense_queries=dq, top_k=6, search_type="dense",
                                                                        filter='metadata["year"] >= 2004 and document_id != "d1"')]
    out["single"] = [(r.id, r.score) for r in store.query(dense_query=dq[0].tolist(), top_k=4, search_type="dense")]
    out["browse"] = [r.id for r in store.query(top_k=100, filter='document_id == "d2"')]
    out["len"] = len(store)
    return out


def _worker(rank, world, port, q_out):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_native
    from verbatim_rag_b200 import _native
    _native.default_context = lambda device=0: fake_native.FakeContext()
    _native.Encoder, _native.Index = fake_native.FakeEncoder, fake_native.FakeIndex
    from verbatim_rag_b200.sharded_store import ShardedB200VectorStore
    store = ShardedB200VectorStore(dense_dim=16, sparse_dim=60)
    out = _answers(store)
    out["local_rows"] = (len(store._dense), len(store._sparse))
    q_out.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_store_equals_unsharded_store(monkeypatch):
    sys.path.insert(0, HERE)
    import fake_native
    fake_native.install(monkeypatch)
    from verbatim_rag_b200.vector_store import B200VectorStore
    want = _answers(B200VectorStore(dense_dim=16, sparse_dim=60))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = [got[r].pop("local_rows") for r in range(2)]
    assert sum(r[0] for r in rows) == 54 and all(r[0] == r[1] and r[0] >= 26 for r in rows)   # vectors really are split
    for r in range(2):
        for key in want:
            assert got[r][key] == want[key], (r, key)
