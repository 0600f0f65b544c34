"""Attention timing table on the GPU box (development aid): the tcgen05 attention kernel alone at the bench's pass size
(256 sequences x 512 tokens) and at ragged / short shapes, global and local (+-64) layers; clocks per 128x64 key block
per SM next to the MUFU / TMEM floors of DESIGN.md section 4.  Writes gpurun_out/attn_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verbatim_rag_b200 import _native  # noqa: E402


def main():
    ctx = _native.default_context(0)
    out = []
    for nseq, L in ((256, 512), (128, 512), (455, 288), (1024, 128)):
        for window in (-1, 64):
            ms = ctx.bench_attention(nseq, L, window, iters=10)
            tiles = nseq * ((L + 127) // 128) * 12
            kblocks = (L + 63) // 64 if window < 0 else min((L + 63) // 64, 4)
            rec = {"nseq": nseq, "seq_len": L, "window": window, "ms": round(ms, 4),
                   "us_per_block_per_sm": round(ms * 1e3 * 148 / (tiles * kblocks), 4)}
            out.append(rec)
            print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "attn_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
