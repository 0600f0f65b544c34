"""Single-purpose workloads for ncu captures (tools/gpu_profile_r2.sh): each target warms up, then runs the kernels of
interest once more so that `--launch-skip` / `--kernel-name` can pick them.

    python tools/profile_targets.py precise_pass | topk_batched | sparse_1m | splade_pass
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402


def precise_pass():
    import torch
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    ctx = _native.default_context(0)
    spec = ModernBertSpec(layers=22)
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, make_modernbert_weights(1001, spec), 22, spec.vocab_size,
                          max_tokens=131072, precision="precise")
    nseq, L = 256, 512
    ids = np.random.default_rng(3).integers(5, 50279, size=(nseq, L), dtype=np.int32)
    ids[:, 0], ids[:, 30], ids[:, -1] = spec.cls_id, spec.sep_id, spec.sep_id
    cu = (np.arange(nseq + 1) * L).astype(np.int32)
    ids_d = torch.from_numpy(ids.reshape(-1)).cuda()
    probs = torch.empty(nseq * L, dtype=torch.float32, device="cuda")
    for _ in range(2):
        enc.span_forward_device(ids_d, cu, probs)
    ctx.sync()


def topk_batched():
    import torch
    ctx = _native.default_context(0)
    n, dim, k = 1_000_000, 768, 10
    g = torch.Generator(device="cuda").manual_seed(0)
    ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    ix.add_dense(torch.randn(n, dim, device="cuda", generator=g))
    for nq in (1000, 16):
        q = torch.randn(nq, dim, device="cuda", generator=g)
        ids = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        sc = torch.empty(nq, k, dtype=torch.float32, device="cuda")
        for _ in range(2):
            ix.search_dense_device(q, nq, k, ids, sc)
        ctx.sync()


def sparse_1m():
    from verbatim_rag_b200.synthetic import make_sparse_rows, make_sparse_rows_device
    ctx = _native.default_context(0)
    ip, ix_, vl = make_sparse_rows_device(1_000_000, seed=1002, device="cuda")
    qip, qix, qvl = make_sparse_rows(64, seed=2002, query=True)
    sx = _native.Index(ctx, _native.INDEX_SPARSE_IP, 30522)
    for a in range(0, 1_000_000, 250_000):
        sx.add_sparse(ip[a:a + 250_001], ix_, vl)
    for _ in range(2):
        sx.search_sparse(qip, qix, qvl, 10)


def splade_pass():
    import torch
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    ctx = _native.default_context(0)
    spec = BertSpec()
    enc = _native.Encoder(ctx, _native.ENC_BERT_MLM, make_bert_mlm_weights(1002, spec), 12, spec.vocab_size, max_tokens=65536)
    n, L = 256, 256
    ids = np.random.default_rng(2).integers(1000, spec.vocab_size, size=(n, L), dtype=np.int32)
    ids[:, 0], ids[:, -1] = spec.cls_id, spec.sep_id
    cu = (np.arange(n + 1) * L).astype(np.int32)
    ids_d = torch.from_numpy(ids.reshape(-1)).cuda()
    dense = torch.empty(n, spec.vocab_size, dtype=torch.float32, device="cuda")
    for _ in range(2):
        enc.splade_forward_device(ids_d, cu, dense)
    ctx.sync()


if __name__ == "__main__":
    {"precise_pass": precise_pass, "topk_batched": topk_batched, "sparse_1m": sparse_1m, "splade_pass": splade_pass}[sys.argv[1]]()
