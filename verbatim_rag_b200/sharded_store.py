"""
ShardedB200VectorStore -- the vector store of BASELINE configs[3] / [4] on N GPUs.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests), every rank runs the same program
(SPMD): the VECTORS of every insert batch are split contiguously over the ranks and live in that rank's HBM, the
PAYLOAD (ids, texts, metadata: host strings) is replicated, so any rank can build ``SearchResult`` objects for any
hit.  A search runs on every rank against its shard, the per-rank top-k blocks are exchanged with ONE all-gather of
packed (score64, id) records and merged on every rank with the order (score desc, global row asc) -- bit-identical to
the unsharded ``B200VectorStore`` (tests/test_distributed_cpu.py, tests/sharded_check.py).

Same interface as ``B200VectorStore`` / the reference's ``LocalMilvusStore`` (verbatim_rag/vector_stores/
milvus_local.py, milvus_base.py:90-127, 189-313); every method must be called on all ranks with the same arguments.
``add_texts`` is the data-parallel index build: each rank ENCODES only its slice of the batch.
"""
from __future__ import annotations

import contextlib
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .distributed import (gather_and_merge, make_peer_exchange, shard_bounds, sharded_search_dense,
                          sharded_search_sparse)
from .vector_store import B200VectorStore, dicts_to_csr


class ShardedB200VectorStore(B200VectorStore):
    def __init__(self, *args, group=None, **kwargs):
        if kwargs.get("db_path") or (args and args[0]):
            raise ValueError("ShardedB200VectorStore is memory-only (db_path must be None)")
        kwargs.pop("id_base", None)
        super().__init__(*args, **kwargs)
        self._group = group
        self._rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._l2g = np.zeros(0, np.int64)      # local row -> global row (= position in the replicated payload lists)
        self._g2l: List[int] = []              # global row -> local row, -1 when another rank holds it
        self._l2g_dev: Optional[torch.Tensor] = None
        self._exchange = False   # False = not set up yet; None = unavailable (NCCL all-gather); else a PeerExchange
        self._on_gpu = any(hasattr(ix, "search_dense_device") for ix in (self._dense, self._sparse) if ix is not None)
        self._device = torch.device("cuda", self._ctx.device) if self._on_gpu else torch.device("cpu")

    # ------------------------------------------------------------------------------------------ insert
    def _device_add(self, dense, csr, first_row: int, n: int, local=None):
        lo, hi = local if local is not None else shard_bounds(n, self._rank, self._world)
        off = 0 if local is not None else lo   # row offset of this rank's slice inside the arrays given
        if self._dense is not None and hi > lo:
            self._dense.add_dense(dense[off:off + hi - lo])
        if self._sparse is not None and hi > lo:
            self._sparse.add_sparse(csr[0][off:off + hi - lo + 1], csr[1], csr[2])
        base = len(self._l2g)
        self._l2g = np.concatenate([self._l2g, first_row + np.arange(lo, hi, dtype=np.int64)])
        g2l = np.full(n, -1, np.int64)
        g2l[lo:hi] = base + np.arange(hi - lo)
        self._g2l.extend(g2l.tolist())
        self._l2g_dev = None

    def add_texts(self, ids, texts, enhanced_texts, metadatas, sparse_provider=None, dense_provider=None):
        """Data-parallel index build: this rank encodes and stores rows [lo, hi) of the batch, the payload of all rows
        is recorded everywhere.  ``sparse_provider.embed_batch_csr`` / ``dense_provider.embed_array`` are used when
        present (the B200 providers), the reference's ``embed_batch`` otherwise."""
        n = len(ids)
        lo, hi = shard_bounds(n, self._rank, self._world)
        mine = list(enhanced_texts[lo:hi])
        dense = csr = None
        if self._dense is not None:
            dense = (dense_provider.embed_array(mine) if hasattr(dense_provider, "embed_array")
                     else np.asarray(dense_provider.embed_batch(mine), np.float32).reshape(len(mine), self.dense_dim))
        if self._sparse is not None:
            if hasattr(sparse_provider, "embed_batch_csr"):
                csr = sparse_provider.embed_batch_csr(mine)
            else:
                csr = dicts_to_csr(sparse_provider.embed_batch(mine))
        self._insert(ids, dense, csr, texts, enhanced_texts, metadatas, local=(lo, hi))

    def _rollback(self, first_row: int, n: int, exc: Exception):
        self._broken = f"insert failed on rank {self._rank} ({exc}); the shards are no longer aligned"

    def _kill_rows(self, rows: Sequence[int], persist: bool = True):
        rows = [r for r in rows if self._alive[r]]
        if not rows:
            return
        local = [self._g2l[r] for r in rows if self._g2l[r] >= 0]
        if local:
            for ix in (self._dense, self._sparse):
                if ix is not None:
                    ix.mark_deleted(local)
        for r in rows:
            self._alive[r] = False
        self._n_deleted += len(rows)

    # ------------------------------------------------------------------------------------------ search
    @contextlib.contextmanager
    def _filtered(self, filter: Optional[str]):
        mask = self._exclude_mask(filter)
        live = [ix for ix in (self._dense, self._sparse) if ix is not None] if mask is not None else []
        if live and len(self._l2g):
            local_mask = np.ascontiguousarray(mask[self._l2g])
            for ix in live:
                ix.set_filter(local_mask)
        try:
            yield
        finally:
            if len(self._l2g):
                for ix in live:
                    ix.set_filter(None)

    def _peer_exchange(self):
        """NVLink peer-memory exchange for the per-shard top-k (distributed.PeerExchange), set up collectively on the
        first search; None (-> NCCL all-gather) when the group has no peer access."""
        if self._exchange is False:
            self._exchange = make_peer_exchange(self._ctx, self._device, self._group)
        return self._exchange

    def _l2g_tensor(self) -> torch.Tensor:
        if self._l2g_dev is None or len(self._l2g_dev) != max(1, len(self._l2g)):
            a = self._l2g if len(self._l2g) else np.zeros(1, np.int64)
            self._l2g_dev = torch.from_numpy(a).to(self._device)
        return self._l2g_dev

    def _hits_dense_batch(self, queries, limit: int) -> List[List[Dict[str, Any]]]:
        q = np.ascontiguousarray(np.asarray(queries, np.float32).reshape(-1, self.dense_dim))
        if self._on_gpu:
            ids, s32, _ = sharded_search_dense(self._dense, torch.from_numpy(q).to(self._device), limit, self._group,
                                               local_to_global=self._l2g_tensor(), exchange=self._peer_exchange())
            ids, s32 = ids.cpu().numpy(), s32.cpu().numpy()
        else:
            lid, _, s64 = self._dense.search_dense(q, limit, want64=True)
            gid = np.where(lid >= 0, self._l2g[np.maximum(lid, 0)] if len(self._l2g) else lid, lid)
            ids, s32, _ = (t.numpy() for t in gather_and_merge(torch.from_numpy(gid), torch.from_numpy(s64), limit,
                                                                self._group))
        return [self._hits(ids[i], s32[i], drop_zero=False) for i in range(ids.shape[0])]

    def _hits_sparse_batch(self, queries: Sequence[Dict[int, float]], limit: int) -> List[List[Dict[str, Any]]]:
        indptr = np.zeros(len(queries) + 1, np.int64)
        idx: List[int] = []
        val: List[float] = []
        for i, qd in enumerate(queries):
            ks = sorted(int(k) for k in qd.keys() if 0 <= int(k) < self.sparse_dim)
            idx.extend(ks)
            val.extend(float(qd[k]) for k in ks)
            indptr[i + 1] = len(idx)
        l2g = self._l2g if len(self._l2g) else np.zeros(1, np.int64)
        ids, s32, _ = sharded_search_sparse(self._sparse, indptr, np.asarray(idx, np.int32), np.asarray(val, np.float32),
                                            limit, self._device, self._group, local_to_global=l2g,
                                            exchange=self._peer_exchange() if self._on_gpu else None)
        ids, s32 = ids.cpu().numpy(), s32.cpu().numpy()
        return [self._hits(ids[i], s32[i], drop_zero=True) for i in range(ids.shape[0])]

    def _search_dense(self, q: Sequence[float], limit: int) -> List[Dict[str, Any]]:
        return self._hits_dense_batch(np.asarray(q, np.float32)[None, :], limit)[0]

    def _search_sparse(self, q: Dict[int, float], limit: int) -> List[Dict[str, Any]]:
        return self._hits_sparse_batch([q], limit)[0]
