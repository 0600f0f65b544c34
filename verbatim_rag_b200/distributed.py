"""
Multi-GPU exact top-k: corpus row-sharded across ranks, queries replicated, ONE collective.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Each rank searches its shard
(``vrag_index_search_*`` with ``id_base`` = global id of its row 0) producing per-query top-k
``(score64, global id)``; ONE ``all_gather`` of those blocks, packed as 16-byte (score64, id) records (160 KB per rank at
1000 queries, k = 10), is the only exchange on the path; every rank then merges the ``world * k`` candidates with
``vrag_topk_merge`` using the same order (score desc, id asc).  Because per-shard scores are fp64 re-scored, the
merged result is bit-identical to the single-GPU search over the concatenated corpus.

``PeerExchange`` replaces the NCCL call by plain stores into NVLink peer memory: every rank's result block is written
straight into a symmetric buffer on every other rank (``vrag_topk_publish``), one device-side barrier follows, and the
merge reads the local buffer (``vrag_topk_merge_packed``) -- no pack / permute kernels, no NCCL launch latency.  It is
used when torch's symmetric-memory rendezvous succeeds for the group (NVLink / NVSwitch peers); otherwise the NCCL
all-gather below carries the same records.  Both give bit-identical results.

On CPU-only machines (tests) the same logic runs over ``gloo`` with the shard search injected by the caller.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of ``rank``: the first ``n_total % world`` ranks get one extra row."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_host(scores64: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Host reference of the merge rule (used on gloo / for checking the device merge): (score desc, id asc);
    id < 0 marks an empty slot."""
    Q = scores64.shape[0]
    out_i = np.full((Q, k), -1, dtype=np.int64)
    out_s = np.full((Q, k), -np.inf, dtype=np.float64)
    for q in range(Q):
        ok = ids[q] >= 0
        s, i = scores64[q][ok], ids[q][ok]
        o = np.lexsort((i, -s))[:k]
        out_i[q, :len(o)] = i[o]
        out_s[q, :len(o)] = s[o]
    return out_i, out_s.astype(np.float32), out_s


def _lib_stream(ctx, device):
    """The library's stream as a torch stream: torch ops and the NCCL collective issued under it are ordered with the
    search / merge kernels on the device, so nothing between scan, gather and merge waits on the host."""
    return torch.cuda.ExternalStream(ctx.stream, device=device)


class PeerExchange:
    """Symmetric exchange buffer for the per-shard top-k blocks: ``[2 slots][world][capacity]`` 16-byte records on every
    rank, each rank's copy mapped into every process (``torch.distributed._symmetric_memory``).  A search writes its
    block into slot ``call & 1`` of every rank (``vrag_topk_publish``: stores over NVLink), waits on one device-side
    barrier and merges its own copy.  Two slots suffice: a rank can only pass barrier ``i`` once every rank has
    enqueued it, i.e. after every rank's merge ``i - 1`` -- so slot ``(i + 1) & 1`` is no longer being read when
    anybody's publish ``i + 1`` lands.  Collective constructor: every rank of ``group`` must create it with the same
    capacity.  Raises if the rendezvous is not possible (no peer access): callers fall back to the NCCL all-gather."""

    def __init__(self, ctx, device, group=None, capacity_records: int = 1024 * 32):
        import torch.distributed._symmetric_memory as symm

        self.ctx, self.device = ctx, torch.device(device)
        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.capacity = int(capacity_records)
        self.buf = symm.empty((2, self.world, self.capacity, 2), dtype=torch.int64, device=self.device)
        self.hdl = symm.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        self.slot_bytes = self.world * self.capacity * 16
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.calls = 0

    def exchange(self, local_ids: torch.Tensor, local_s64: torch.Tensor, k: int):
        """(ids [Q,k] int64 global, fp64 scores [Q,k]) of this rank -> global top-k on every rank.  Enqueued on the
        library's stream; no host synchronisation."""
        Q = local_ids.shape[0]
        if Q * k > self.capacity:
            raise ValueError(f"PeerExchange: {Q} x {k} records exceed the capacity {self.capacity}")
        dev = self.device
        st = _lib_stream(self.ctx, dev)
        st.wait_stream(torch.cuda.current_stream(dev))
        slot = self.calls & 1
        self.calls += 1
        with torch.cuda.stream(st):
            ids_c, s64_c = local_ids.contiguous(), local_s64.contiguous()
            # the per-call record layout is [world][Q][k] from the start of the slot
            self.ctx.topk_publish(s64_c, ids_c, Q, k, [p + slot * self.slot_bytes for p in self.peer_ptrs], self.rank)
            self.hdl.barrier(channel=0)
            out_i = torch.empty((Q, k), dtype=torch.int64, device=dev)
            out_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
            out_d = torch.empty((Q, k), dtype=torch.float64, device=dev)
            self.ctx.topk_merge_packed(self.buf.data_ptr() + slot * self.slot_bytes, self.world, Q, k, out_i, out_s, out_d)
            for t in (ids_c, s64_c):
                t.record_stream(st)
        torch.cuda.current_stream(dev).wait_stream(st)
        return out_i, out_s, out_d


def make_peer_exchange(ctx, device, group=None, capacity_records: int = 1024 * 32):
    """``PeerExchange`` when every rank of the group can set it up, else ``None`` (the NCCL all-gather is used).  The
    decision is collective: one all-reduce of a success flag, so that no rank waits on a barrier the others skipped."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or dist.get_backend(group) != "nccl":
        return None
    ok, ex = 1, None
    try:
        ex = PeerExchange(ctx, device, group, capacity_records)
    except Exception as e:  # noqa: BLE001 -- any rendezvous failure means "no peer memory here"
        import logging
        logging.getLogger(__name__).warning("peer-memory exchange unavailable (%s); using the NCCL all-gather", e)
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return ex if int(flag.item()) == 1 else None


def gather_and_merge(local_ids: torch.Tensor, local_s64: torch.Tensor, k: int, group=None, ctx=None, exchange=None):
    """all_gather the per-rank top-k blocks and merge them to the global top-k.

    ONE collective: each rank contributes a packed ``[Q, k, 2]`` int64 block (fp64 score bits, global id) = 16 bytes per
    candidate (160 KB per rank at 1000 queries, k = 10).  CUDA tensors + ``ctx`` (a ``_native.Context``): everything is
    enqueued on the library's stream -- pack, NCCL all_gather, unpack, ``vrag_topk_merge`` -- with no host
    synchronisation; the returned tensors are ready once that stream is (the caller's current stream is made to wait
    for it).  CPU tensors: gloo all_gather and the host merge.  Returns (ids [Q,k] int64, scores fp32, scores fp64).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    Q = local_ids.shape[0]
    if exchange is not None and local_ids.is_cuda and world > 1 and Q * k <= exchange.capacity:
        return exchange.exchange(local_ids, local_s64, k)
    if not local_ids.is_cuda:
        packed = torch.stack((local_s64.contiguous().view(torch.int64), local_ids.contiguous()), dim=-1)
        if world > 1:
            g = torch.empty((world * Q, k, 2), dtype=torch.int64)
            dist.all_gather_into_tensor(g, packed, group=group)          # rank r -> rows [r * Q, (r + 1) * Q)
            packed = g.view(world, Q, k, 2).permute(1, 0, 2, 3).reshape(Q, world * k, 2)
        oi, os32, os64 = merge_host(packed[..., 0].contiguous().view(torch.float64).numpy(),
                                    packed[..., 1].contiguous().numpy(), k)
        return torch.from_numpy(oi), torch.from_numpy(os32), torch.from_numpy(os64)
    assert ctx is not None, "device merge needs the native context"
    dev = local_ids.device
    st = _lib_stream(ctx, dev)
    st.wait_stream(torch.cuda.current_stream(dev))      # inputs produced on the caller's stream (device-side wait)
    with torch.cuda.stream(st):
        if world > 1:
            packed = torch.stack((local_s64.contiguous().view(torch.int64), local_ids.contiguous()), dim=-1)
            g = torch.empty((world * Q, k, 2), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(g, packed, group=group)          # rank r -> rows [r * Q, (r + 1) * Q)
            g = g.view(world, Q, k, 2).permute(1, 0, 2, 3).reshape(Q, world * k, 2)
            all_s = g[..., 0].contiguous().view(torch.float64)
            all_i = g[..., 1].contiguous()
        else:
            all_s, all_i = local_s64.contiguous(), local_ids.contiguous()
        out_i = torch.empty((Q, k), dtype=torch.int64, device=dev)
        out_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
        out_d = torch.empty((Q, k), dtype=torch.float64, device=dev)
        ctx.topk_merge(all_s, all_i, Q, all_i.shape[1], k, out_i, out_s, out_d)
        for t in (all_s, all_i):
            t.record_stream(st)
    torch.cuda.current_stream(dev).wait_stream(st)
    return out_i, out_s, out_d


def sharded_search_dense(index, queries: torch.Tensor, k: int, group=None, local_to_global: Optional[torch.Tensor] = None,
                         exchange: Optional[PeerExchange] = None):
    """``index``: this rank's ``_native.Index`` (id_base set, or ``local_to_global`` [n_local] int64 mapping its rows
    to global ids).  ``queries`` [Q, dim] fp32 CUDA tensor, replicated.  No host synchronisation."""
    Q = queries.shape[0]
    dev = queries.device
    st = _lib_stream(index.ctx, dev)
    st.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(st):
        ids = torch.empty((Q, k), dtype=torch.int64, device=dev)
        s32 = torch.empty((Q, k), dtype=torch.float32, device=dev)
        s64 = torch.empty((Q, k), dtype=torch.float64, device=dev)
        index.search_dense_device(queries, Q, k, ids, s32, s64)
        if local_to_global is not None:
            ids = torch.where(ids >= 0, local_to_global[ids.clamp_min(0)], ids)
        queries.record_stream(st)
    torch.cuda.current_stream(dev).wait_stream(st)
    return gather_and_merge(ids, s64, k, group, index.ctx, exchange)


def sharded_search_sparse(index, q_indptr, q_indices, q_values, k: int, device, group=None,
                          local_to_global: Optional[np.ndarray] = None, exchange: Optional[PeerExchange] = None):
    """Sparse queries as host CSR (replicated); returns global (ids, fp32 scores, fp64 scores) tensors on ``device``."""
    ids, s32, s64 = index.search_sparse(q_indptr, q_indices, q_values, k, want64=True)
    if local_to_global is not None:
        ids = np.where(ids >= 0, local_to_global[np.maximum(ids, 0)], ids)
    return gather_and_merge(torch.from_numpy(ids).to(device), torch.from_numpy(s64).to(device), k, group, index.ctx,
                            exchange)
