"""One 16-query dense search over 1 M x 768 (for an ncu launch list of the scan / select / rescore / rank kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402

ctx = _native.default_context(0)
n, dim, k = 1_000_000, 768, 10
g = torch.Generator(device="cuda").manual_seed(0)
ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
ix.add_dense(torch.randn(n, dim, device="cuda", generator=g))
for nq in (16, 1):
    q = torch.randn(nq, dim, device="cuda", generator=g)
    ids = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    sc = torch.empty(nq, k, dtype=torch.float32, device="cuda")
    for _ in range(2):
        ix.search_dense_device(q, nq, k, ids, sc)
    ctx.sync()
