"""torchrun worker: row-sharded dense + sparse top-k over NCCL must equal the unsharded search bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from verbatim_rag_b200 import _native  # noqa: E402
from verbatim_rag_b200.distributed import (make_peer_exchange, shard_bounds, sharded_search_dense,  # noqa: E402
                                           sharded_search_sparse)
from verbatim_rag_b200.synthetic import make_dense_corpus, make_dense_queries, make_sparse_rows  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _native.default_context(local)
    n, k = 60001, 10
    corpus = make_dense_corpus(n, 768, seed=1004)
    corpus[5] = corpus[n - 7]                      # an exact tie across shards
    queries = make_dense_queries(24, 768, seed=2004)
    lo, hi = shard_bounds(n, rank, world)
    shard = _native.Index(ctx, _native.INDEX_DENSE_COSINE, 768)
    shard.set_id_base(lo)
    shard.add_dense(corpus[lo:hi])
    ids, s32, s64 = sharded_search_dense(shard, torch.from_numpy(queries).to(dev), k)
    full = _native.Index(ctx, _native.INDEX_DENSE_COSINE, 768)
    full.add_dense(corpus)
    fid, fs32, fs64 = full.search_dense(queries, k, want64=True)
    assert np.array_equal(ids.cpu().numpy(), fid), "dense ids differ"
    assert np.array_equal(s64.cpu().numpy(), fs64), "dense fp64 scores differ"
    # the same exchange as stores into NVLink peer memory (symmetric buffer + device barrier) instead of NCCL
    ex = make_peer_exchange(ctx, dev)
    if ex is not None:
        qd_dev = torch.from_numpy(queries).to(dev)
        for rep in range(5):   # several calls: the two slots of the exchange buffer alternate
            pid, _, ps64 = sharded_search_dense(shard, qd_dev if rep % 2 == 0 else qd_dev[:7], k, exchange=ex)
            nq = pid.shape[0]
            assert np.array_equal(pid.cpu().numpy(), fid[:nq]), "dense ids differ (peer exchange)"
            assert np.array_equal(ps64.cpu().numpy(), fs64[:nq]), "dense fp64 scores differ (peer exchange)"
    if rank == 0:
        print("PEER_EXCHANGE", "on" if ex is not None else "unavailable", flush=True)

    ip, ix, vl = make_sparse_rows(8000, seed=1002)
    qip, qix, qvl = make_sparse_rows(16, seed=2002, query=True)
    lo, hi = shard_bounds(8000, rank, world)
    ss = _native.Index(ctx, _native.INDEX_SPARSE_IP, 30522)
    ss.set_id_base(lo)
    ss.add_sparse(ip[lo:hi + 1], ix, vl)
    sid, _, sd = sharded_search_sparse(ss, qip, qix, qvl, k, dev)
    fsx = _native.Index(ctx, _native.INDEX_SPARSE_IP, 30522)
    fsx.add_sparse(ip, ix, vl)
    rid, _, rd = fsx.search_sparse(qip, qix, qvl, k, want64=True)
    assert np.array_equal(sid.cpu().numpy(), rid), "sparse ids differ"
    assert np.array_equal(sd.cpu().numpy(), rd), "sparse fp64 scores differ"
    if ex is not None:
        sid, _, sd = sharded_search_sparse(ss, qip, qix, qvl, k, dev, exchange=ex)
        assert np.array_equal(sid.cpu().numpy(), rid), "sparse ids differ (peer exchange)"
        assert np.array_equal(sd.cpu().numpy(), rd), "sparse fp64 scores differ (peer exchange)"
    # the plugin-level store: vectors split over the ranks, payload replicated, one all-gather per search
    from verbatim_rag_b200 import B200VectorStore
    from verbatim_rag_b200.sharded_store import ShardedB200VectorStore
    from verbatim_rag_b200.synthetic import csr_to_dicts
    m = 3001
    cid = [f"c{i:05d}" for i in range(m)]
    texts = [f"text {i}" for i in range(m)]
    metas = [{"year": 2000 + i % 20} for i in range(m)]
    sp_rows = csr_to_dicts(ip[:m + 1], ix, vl)
    qd = csr_to_dicts(qip, qix, qvl)
    flt = 'metadata["year"] >= 2010'
    answers = []
    for cls in (ShardedB200VectorStore, B200VectorStore):
        st = cls(dense_dim=768, enable_dense=True, enable_sparse=True, device=f"cuda:{local}")
        for a in range(0, m, 1000):
            b = min(m, a + 1000)
            st.add_vectors(cid[a:b], corpus[a:b].tolist(), sp_rows[a:b], texts[a:b], texts[a:b], metas[a:b])
        st.delete([cid[7], cid[2999]])
        answers.append((
            [[(r.id, r.score) for r in rs] for rs in st.query_batch(dense_queries=queries[:9], top_k=k, search_type="dense")],
            [[(r.id, r.score) for r in rs] for rs in st.query_batch(sparse_queries=qd[:9], top_k=k, search_type="sparse")],
            [[r.id for r in rs] for rs in st.query_batch(dense_queries=queries[:9], top_k=k, search_type="dense", filter=flt)],
            [(r.id, r.score, r.text) for r in st.query(dense_query=queries[0].tolist(), top_k=5, search_type="dense")]))
    assert answers[0] == answers[1], "sharded store differs from the unsharded store"
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
