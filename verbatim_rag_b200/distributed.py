"""
Multi-GPU exact top-k: corpus row-sharded across ranks, queries replicated, ONE collective.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Each rank searches its shard
(``vrag_index_search_*`` with ``id_base`` = global id of its row 0) producing per-query top-k
``(score64, global id)``; an ``all_gather`` of those ``[Q, k]`` blocks (120 KB per rank at 1000 queries, k = 10)
is the only exchange on the path; every rank then merges the ``world * k`` candidates with
``vrag_topk_merge`` using the same order (score desc, id asc).  Because per-shard scores are fp64 re-scored, the
merged result is bit-identical to the single-GPU search over the concatenated corpus.

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


def gather_and_merge(local_ids: torch.Tensor, local_s64: torch.Tensor, k: int, group=None, ctx=None):
    """all_gather the per-rank [Q, k] (id int64, score fp64) blocks and merge to the global top-k.

    CUDA tensors + ``ctx`` (a ``_native.Context``): NCCL all_gather and the device merge kernel.
    CPU tensors: gloo all_gather and the host merge.  Returns (ids [Q,k] int64, scores fp32, scores fp64).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    Q = local_ids.shape[0]
    if world == 1:
        all_i, all_s = local_ids, local_s64
    else:
        gi = torch.empty((world * Q, k), dtype=local_ids.dtype, device=local_ids.device)
        gs = torch.empty((world * Q, k), dtype=local_s64.dtype, device=local_s64.device)
        dist.all_gather_into_tensor(gi, local_ids.contiguous(), group=group)    # rank r -> rows [r*Q, (r+1)*Q)
        dist.all_gather_into_tensor(gs, local_s64.contiguous(), group=group)
        all_i = gi.view(world, Q, k).permute(1, 0, 2).reshape(Q, world * k).contiguous()
        all_s = gs.view(world, Q, k).permute(1, 0, 2).reshape(Q, world * k).contiguous()
    if all_i.is_cuda:
        assert ctx is not None, "device merge needs the native context"
        out_i = torch.empty((Q, k), dtype=torch.int64, device=all_i.device)
        out_s = torch.empty((Q, k), dtype=torch.float32, device=all_i.device)
        out_d = torch.empty((Q, k), dtype=torch.float64, device=all_i.device)
        torch.cuda.current_stream(all_i.device).synchronize()   # NCCL ran on torch's stream, merge runs on ours
        ctx.topk_merge(all_s, all_i, Q, all_i.shape[1], k, out_i, out_s, out_d)
        ctx.sync()
        return out_i, out_s, out_d
    oi, os32, os64 = merge_host(all_s.numpy(), all_i.numpy(), k)
    return torch.from_numpy(oi), torch.from_numpy(os32), torch.from_numpy(os64)


def sharded_search_dense(index, queries: torch.Tensor, k: int, group=None):
    """``index``: this rank's ``_native.Index`` (id_base set).  ``queries`` [Q, dim] fp32 CUDA tensor, replicated."""
    Q = queries.shape[0]
    ids = torch.empty((Q, k), dtype=torch.int64, device=queries.device)
    s32 = torch.empty((Q, k), dtype=torch.float32, device=queries.device)
    s64 = torch.empty((Q, k), dtype=torch.float64, device=queries.device)
    torch.cuda.current_stream(queries.device).synchronize()
    index.search_dense_device(queries, Q, k, ids, s32, s64)
    index.ctx.sync()
    return gather_and_merge(ids, s64, k, group, index.ctx)


def sharded_search_sparse(index, q_indptr, q_indices, q_values, k: int, device, group=None):
    """Sparse queries as host CSR (replicated); returns global (ids, fp32 scores, fp64 scores) CUDA tensors."""
    ids, s32, s64 = index.search_sparse(q_indptr, q_indices, q_values, k, want64=True)
    return gather_and_merge(torch.from_numpy(ids).to(device), torch.from_numpy(s64).to(device), k, group, index.ctx)
