"""N>1 host logic on CPU: world_size-2 gloo all_gather + merge equals the single-shard oracle top-k."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from verbatim_rag_b200.distributed import gather_and_merge, merge_host, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _worker(rank, world, port, n, k, q_out):
    from oracle.flat_topk import dense_cosine_scores
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    corpus = rng.standard_normal((n, 32)).astype(np.float32)
    corpus[5] = corpus[n - 3]                     # an exact cross-shard tie
    queries = rng.standard_normal((6, 32)).astype(np.float32)
    lo, hi = shard_bounds(n, rank, world)
    sc = dense_cosine_scores(corpus[lo:hi], queries)
    ids = np.full((6, k), -1, np.int64)
    s64 = np.full((6, k), -np.inf)
    for q in range(6):
        o = np.lexsort((np.arange(hi - lo), -sc[q]))[:k]
        ids[q, :len(o)] = o + lo
        s64[q, :len(o)] = sc[q][o]
    gi, gs32, gs64 = gather_and_merge(torch.from_numpy(ids), torch.from_numpy(s64), k)
    if rank == 0:
        q_out.put((gi.numpy(), gs64.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(101, 10), (7, 10)])
def test_two_rank_gather_merge_equals_global_topk(n, k):
    from oracle.flat_topk import dense_cosine_scores
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, s64 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    corpus = rng.standard_normal((n, 32)).astype(np.float32)
    corpus[5] = corpus[n - 3]
    queries = rng.standard_normal((6, 32)).astype(np.float32)
    sc = dense_cosine_scores(corpus, queries)
    for qi in range(6):
        o = np.lexsort((np.arange(n), -sc[qi]))[:k]
        assert ids[qi, :len(o)].tolist() == o.tolist()
        assert np.array_equal(s64[qi, :len(o)], sc[qi][o])
        assert np.all(ids[qi, len(o):] == -1)


def test_merge_host_rule():
    s = np.array([[0.5, 0.9, 0.9, -np.inf, 0.1]])
    i = np.array([[7, 3, 2, -1, 9]])
    oi, o32, o64 = merge_host(s, i, 3)
    assert oi.tolist() == [[2, 3, 7]] and o64.tolist() == [[0.9, 0.9, 0.5]]
