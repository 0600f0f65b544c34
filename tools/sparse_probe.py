"""Times the sparse search (sparse half) at the two shapes DESIGN.md quotes: 10 k docs x 1000
queries (bench config 3) and 1 M docs x 64 queries.  Usage on a GPU box: python tools/sparse_probe.py"""
import sys
import time

sys.path.insert(0, ".")
from verbatim_rag_b200 import _native  # noqa: E402
from verbatim_rag_b200.synthetic import make_sparse_rows, make_sparse_rows_device  # noqa: E402


def timed(ix, ctx, args, label):
    ix.search_sparse(*args)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ix.search_sparse(*args)
        best = min(best, time.perf_counter() - t0)
    ctx.profile(True)
    ix.search_sparse(*args)
    pr = ctx.profile_read()
    ctx.profile(False)
    print("%s: %.2f ms" % (label, best * 1e3),
          {k: (round(v["ms"], 3), v["launches"]) for k, v in pr.items() if v["launches"]}, flush=True)


def main():
    ctx = _native.default_context(0)
    ip, ixs, vl = make_sparse_rows(10000, seed=1002)
    qip, qix, qvl = make_sparse_rows(1000, seed=2002, query=True)
    sx = _native.Index(ctx, _native.INDEX_SPARSE_IP, 30522)
    sx.add_sparse(ip, ixs, vl)
    timed(sx, ctx, (qip, qix, qvl, 10), "sparse 10k docs, 1000 queries")
    bip, bix, bvl = make_sparse_rows_device(1000000, seed=1002, device="cuda")
    bx = _native.Index(ctx, _native.INDEX_SPARSE_IP, 30522)
    for a in range(0, 1000000, 250000):
        bx.add_sparse(bip[a:a + 250001], bix, bvl)
    timed(bx, ctx, (qip[:65], qix, qvl, 10), "sparse 1M docs (nnz %d), 64 queries" % int(bip[-1]))
    timed(bx, ctx, (qip[:2], qix, qvl, 10), "sparse 1M docs, 1 query")


if __name__ == "__main__":
    main()
