"""SPLADE encode timing (BERT-base MLM, 256-token chunks; BASELINE configs[1] shape) with the per-class profile.
Development aid; bench.py reports the number."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402
from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights  # noqa: E402


def main():
    ctx = _native.default_context(0)
    bspec = BertSpec()
    enc = _native.Encoder(ctx, _native.ENC_BERT_MLM, make_bert_mlm_weights(1002, bspec), bspec.layers,
                          bspec.vocab_size, max_tokens=int(os.environ.get("SPLADE_MAX_TOKENS", "65536")))
    nchunk, L = int(os.environ.get("SPLADE_CHUNKS", "1024")), 256
    rng = np.random.default_rng(1002)
    ids = rng.integers(1000, bspec.vocab_size, size=(nchunk, L), dtype=np.int32)
    ids[:, 0], ids[:, -1] = bspec.cls_id, bspec.sep_id
    cu = (np.arange(nchunk + 1) * L).astype(np.int32)
    ids_d = torch.from_numpy(ids.reshape(-1)).cuda()
    dense_d = torch.empty(nchunk, bspec.vocab_size, dtype=torch.float32, device="cuda")
    enc.splade_forward_device(ids_d, cu, dense_d)
    ctx.sync()
    ctx.profile(True)
    enc.splade_forward_device(ids_d, cu, dense_d)
    ctx.sync()
    pr = ctx.profile_read()
    ctx.profile(False)
    tot = sum(v["ms"] for v in pr.values())
    print(json.dumps({"chunks": nchunk, "ms": tot, "chunks_per_s": nchunk / tot * 1e3,
                      "tflops_algorithmic": 58.21e9 * nchunk / tot / 1e9, "classes": pr}, indent=1))


if __name__ == "__main__":
    main()
