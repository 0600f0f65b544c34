#!/usr/bin/env python
"""
bench.py -- the contract benchmark of the B200 hot path (BASELINE.json metric, configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference-shaped CPU path (oracle port)

One "step" = one pass of the span-extraction hot path over one batch of synthetic input:
256 questions x 16 retrieved chunks = 4096 sequences of exactly 512 tokens
(``[CLS] + 29 question tokens + [SEP] + 480 chunk tokens + [SEP]``), ModernBERT-base token classifier, 22 layers,
seeded random-init weights (no checkpoint exists offline).  Weak scaling: every rank processes its own 4096
sequences per step; no collective on this path.

Printed JSON line (rank 0): the base contract keys plus
  e2e            the same step through the C ABI with HOST buffers: ids H2D, forward, probs D2H, span post-processing
  modes          both arithmetic modes: "fast" (fp16 operands, the headline `value`) and "precise" (split-precision
                 operands, span logits within 1e-3 of the fp32 reference), each with value + e2e
  roofline       tensor-core GEMM class: algorithmic FLOPs / summed CUDA-event kernel time, vs MEASURED_PEAKS.json
  cpu_baseline   the reference-shaped CPU plugin (oracle port) on this box's host cores (bounded sample; rank 0, N=1)
  secondary      the other BASELINE configs, same run, each with value, roofline, host-buffer e2e and cpu_baseline:
                 cfg2 SPLADE encode (10 k chunks @256) + sparse top-10 (10 k docs, 1 k queries; 1 M-doc variant),
                 cfg4 dense top-10 over 1 M x 768 (HBM-bound 16-query passes AND the tensor-bound batched GEMM search),
                 cfg5 SPLADE retrieve top-20 + span extraction, 10 k chunks / 1 k queries (sharded over the ranks)
Both arms print the same `config` dict.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "(question,512-tok chunk) span-extractions/sec"
UNIT = "extractions/s"
SEQ_LEN, Q_LEN = 512, 29
CTX_LEN = SEQ_LEN - Q_LEN - 3
CHUNKS_PER_Q = 16
SEQS_PER_STEP = 256 * CHUNKS_PER_Q
LAYERS = 22
FLOPS_LINEAR_PER_TOKEN = 221.78e6          # all Linear layers incl. head
FLOPS_PER_EXTRACTION = 122.65e9            # linears + windowed attention at L = 512
FLOPS_PER_SPLADE_CHUNK = 58.21e9           # BERT-base MLM at L = 256
WORKLOAD = "ModernBERT-v2 span extraction: 256-query batch x 16 retrieved chunks @512 tok (configs[2])"


def bench_config():
    """The workload, identical in both arms (the arms differ in `impl`, `dtype` and the sample they time)."""
    return {"workload": WORKLOAD, "seqs_per_step_per_gpu": SEQS_PER_STEP, "seq_len": SEQ_LEN, "question_tokens": Q_LEN,
            "chunk_tokens": CTX_LEN, "layers": LAYERS, "threshold": 0.2, "min_span_chars": 30, "merge_gap_chars": 20,
            "precision": "reference: fp32 on CPU; b200 arm: `value` = fast mode (fp16 tensor-core operands, |dlogit| <= "
                         "4e-3), modes.precise = split-precision operands (|dlogit| <= 1e-3, tests/test_gpu_precise.py)",
            "l2": "inputs larger than L2: >= 11.5 KB of activations per token, 24 GB per step vs 126 MB"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"source": "measured (MEASURED_PEAKS.json)", "hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"])}
    return {"source": "fallback (B200_PROFILING.md)", "hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0}


def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_batch(nseq: int, seed: int):
    """Token ids of ``nseq`` sequences ``[CLS] q(29) [SEP] chunk(480) [SEP]`` + synthetic char offsets of the chunk
    tokens (7 chars per token incl. the separating space) for the span post-processing."""
    from verbatim_rag_b200.synthetic import ModernBertSpec
    spec = ModernBertSpec()
    rng = np.random.default_rng(seed)
    ids = rng.integers(5, 50279, size=(nseq, SEQ_LEN), dtype=np.int32)
    ids[:, 0] = spec.cls_id
    ids[:, Q_LEN + 1] = spec.sep_id
    ids[:, -1] = spec.sep_id
    cu = (np.arange(nseq + 1, dtype=np.int64) * SEQ_LEN).astype(np.int32)
    tok_cs = np.tile(np.arange(CTX_LEN, dtype=np.int32) * 7, nseq)
    tok_ce = tok_cs + 6
    ctx_indptr = np.arange(nseq + 1, dtype=np.int64) * CTX_LEN
    return ids.reshape(-1), cu, tok_cs, tok_ce, ctx_indptr


class _R:   # the only attribute the extractor contract reads from a search result (extractors.py:208)
    def __init__(self, t):
        self.text = t


def make_text_batch(tk, n_questions: int, seed: int, chunks_per_q: int = CHUNKS_PER_Q, chunk_tokens: int = CTX_LEN,
                    q_tokens: int = Q_LEN):
    rng = np.random.default_rng(seed)
    qs = [tk.make_question(rng, q_tokens) for _ in range(n_questions)]
    rs = [[_R(tk.make_text(rng, chunk_tokens)) for _ in range(chunks_per_q)] for _ in range(n_questions)]
    return qs, rs


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference-shaped plugin (oracle/plugins.py OracleSpanExtractor): tokenise, ONE batch-1 forward per
# chunk, sequentially, then the process() post-processing -- the control flow of extractors.py:203-228
# --------------------------------------------------------------------------------------------------------------
def time_cpu_spans(n_pairs: int, steps: int, warmup: int):
    import torch
    try:   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core this process may run on
        torch.set_num_threads(_host_cores())
    except (AttributeError, RuntimeError):
        pass
    from oracle.plugins import OracleSpanExtractor
    from verbatim_rag_b200.synthetic import ModernBertSpec, SyntheticTokenizer, make_modernbert_weights
    spec = ModernBertSpec(layers=LAYERS)
    tk = SyntheticTokenizer("modernbert")
    weights = {k: torch.from_numpy(v) for k, v in make_modernbert_weights(1001, spec).items()}
    ext = OracleSpanExtractor(weights, tk, spec)
    nq = max(1, (n_pairs + CHUNKS_PER_Q - 1) // CHUNKS_PER_Q)
    qs, rs = make_text_batch(tk, nq, 1003)
    flat = [(q, r) for q, row in zip(qs, rs) for r in row][:n_pairs]
    for _ in range(warmup):
        ext.extract_spans(flat[0][0], [flat[0][1]])
    t0 = time.perf_counter()
    nspans = 0
    for _ in range(steps):
        for q, r in flat:
            nspans += sum(len(v) for v in ext.extract_spans(q, [r]).values())
    dt = (time.perf_counter() - t0) / steps
    return n_pairs / dt, dt, torch.get_num_threads(), nspans // max(steps, 1)


def cpu_baseline_block(rate, cores, sample):
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} of {SEQS_PER_STEP} (question, chunk) pairs per step, strings in: tokenise + batch-1 fp32 "
                      "forward per chunk + span post-processing (oracle/plugins.py OracleSpanExtractor, the control flow "
                      "of extractors.py:203-228; torch CPU, all host cores)"}


def run_reference_arm(args, out):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sample = args.cpu_sample
    rate, dt, cores, _ = time_cpu_spans(sample, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(),
        "cpu_baseline": cpu_baseline_block(rate, cores, sample),
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()


# --------------------------------------------------------------------------------------------------------------
# GPU arm helpers
# --------------------------------------------------------------------------------------------------------------
class Timer:
    """CUDA events on the library's stream (torch.cuda.Event on an ExternalStream of ctx.stream)."""

    def __init__(self, ctx, device):
        import torch
        self.ctx, self.torch = ctx, torch
        self.st = torch.cuda.ExternalStream(ctx.stream, device=device)

    def __call__(self, fn, iters=3, warm=1):
        torch = self.torch
        for _ in range(warm):
            fn()
        self.ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.st)
        for _ in range(iters):
            fn()
        e1.record(self.st)
        self.ctx.sync()
        return e0.elapsed_time(e1) / iters

    def profiled(self, fn, iters=1):
        self.ctx.profile(True)
        for _ in range(iters):
            fn()
        pr = self.ctx.profile_read()
        self.ctx.profile(False)
        return pr


def _all_max(x: float, device, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sec_dense(ctx, peaks, rank, world, device, cpu: bool):
    """BASELINE configs[3]: dense cosine top-10, 1 M x 768 fp32 corpus row-sharded over the ranks, both regimes."""
    import torch
    import torch.distributed as dist
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.distributed import shard_bounds, sharded_search_dense
    out = {}
    timed = Timer(ctx, device)
    n_total, dim, k = 1_000_000, 768, 10
    lo, hi = shard_bounds(n_total, rank, world)
    rows = hi - lo
    g = torch.Generator(device=device).manual_seed(1004 + rank)
    ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    ix.set_id_base(lo)
    ix.add_dense(torch.randn(rows, dim, device=device, generator=g))
    gq = torch.Generator(device=device).manual_seed(2004)
    pass_bytes = rows * dim * 4
    for nq in (1, 16, 1000):   # the reference's per-query call; one HBM-bound 16-query pass; the configs[3] batch
        q = torch.randn(nq, dim, device=device, generator=gq)
        ids_o = torch.empty(nq, k, dtype=torch.int64, device=device)
        s_o = torch.empty(nq, k, dtype=torch.float32, device=device)
        fn = lambda: ix.search_dense_device(q, nq, k, ids_o, s_o)   # noqa: E731
        ms = timed(fn, iters=3, warm=2)
        pr = timed.profiled(fn, iters=3)
        scan_ms = pr["scan"]["ms"] / 3
        launches = pr["scan"]["launches"] / 3
        rec = {"ms": ms, "queries_per_s": nq / ms * 1e3, "rows_per_gpu": rows, "scan_launches": launches,
               "scan_ms": scan_ms, "select_finish_ms": pr["select"]["ms"] / 3}
        if nq <= 16:    # HBM-bound regime: one corpus pass of N*dim*4 bytes per <= 16 queries
            gbs = launches * pass_bytes / scan_ms / 1e6 if scan_ms > 0 else 0.0
            rec["search_GBps_per_gpu"] = launches * pass_bytes / ms / 1e6
            rec["roofline"] = {"bound": "hbm", "kernel": "dense_scan_tma_kernel (fp32 FMA)" if nq < 5 else
                               "dense_scan_tc_kernel (split-tf32 tcgen05, 16 queries per pass)", "achieved": gbs,
                               "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                               "bytes_per_launch": pass_bytes, "avg_launch_ms": scan_ms / max(launches, 1)}
        else:           # tensor-bound regime: all queries against the corpus as one split-precision GEMM per 1024
            flops = 2.0 * nq * rows * dim
            tf = 3 * flops / scan_ms / 1e9 if scan_ms > 0 else 0.0        # three fp16 MMAs per product
            plane_bytes = ((nq + 255) // 256) * rows * dim * 4              # hi + lo planes, read once per 256 queries
            rec["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<EPI_SCORES_THRESH, split>: 3 fp16 MMAs per "
                               "product, selection fused into the epilogue (no [nq, N] score matrix)",
                               "achieved": tf, "peak": peaks["bf16_burst"], "unit": "TFLOP/s issued (3 x algorithmic)",
                               "frac": tf / peaks["bf16_burst"], "algorithmic_tflops": flops / scan_ms / 1e9,
                               "hbm_GBps": plane_bytes / scan_ms / 1e6, "hbm_frac": plane_bytes / scan_ms / 1e6 / peaks["hbm_gbs"],
                               "gemm_launches": launches, "avg_launch_ms": scan_ms / max(launches, 1)}
        out[f"top{k}_q{nq}"] = rec
    # host-buffer e2e through the C ABI: queries H2D + search + ids / scores D2H inside the call
    qh = np.random.default_rng(2004).standard_normal((1000, dim), dtype=np.float32)
    ix.search_dense(qh, k)   # warm-up at the timed batch size: the first call of a size allocates its device buffers
    dt = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        ix.search_dense(qh, k)
        dt = min(dt, time.perf_counter() - t0)
    out["e2e"] = {"value": 1000 / dt, "unit": "queries/s", "h2d_bytes": int(qh.nbytes), "d2h_bytes": 1000 * k * 12,
                  "path": "vrag_index_search_dense(host queries) -> host ids + scores, one call for 1000 queries"}
    if world > 1:
        # configs[3] as stated: queries replicated, per-rank top-k, ONE exchange of the [1000, 10] blocks + device merge, no
        # host sync: as stores into NVLink peer memory (PeerExchange) when the group has peer access, and as one packed
        # NCCL all_gather (timed beside it)
        from verbatim_rag_b200.distributed import make_peer_exchange
        q = torch.randn(1000, dim, device=device, generator=gq)
        ex = make_peer_exchange(ctx, device)

        def timed_global(exchange):
            for _ in range(2):
                sharded_search_dense(ix, q, k, exchange=exchange)
            dist.barrier()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                gi, gs, _ = sharded_search_dense(ix, q, k, exchange=exchange)
            e1.record()
            torch.cuda.synchronize(device)
            return _all_max(e0.elapsed_time(e1) / 3, device, world), gi, gs

        ms_nccl, gi_n, gs = timed_global(None)
        ms = ms_nccl
        if ex is not None:
            ms, gi_p, gs = timed_global(ex)
            assert bool((gi_p == gi_n).all().item()), "peer exchange and NCCL all-gather disagree"
        out["top10_q1000_global"] = {
            "ms": ms, "queries_per_s": 1000 / ms * 1e3, "ranks": world, "rows_total": n_total,
            "exchange": ("vrag_topk_publish: each rank stores its [1000, 10] x (f64 score, i64 id) block into every peer's "
                         "symmetric buffer over NVLink + one device barrier + vrag_topk_merge_packed") if ex is not None else
                        "one NCCL all_gather of packed [1000, 10] x (f64 score, i64 id) per rank + vrag_topk_merge",
            "ms_with_nccl_all_gather": ms_nccl, "device_ordered": "everything on the library stream, no host sync",
            "ids_sorted_by_score": bool((gs[:, :-1] >= gs[:, 1:]).all().item())}
    ix.close()
    if cpu:
        # the reference's store is milvus-lite FLAT on the host cores, one call per query (milvus_base.py:239-248):
        # oracle/flat_topk.py exact scan, 100 k of the 1 M rows x 32 queries
        from oracle.flat_topk import dense_cosine_topk
        rng = np.random.default_rng(1004)
        sub = rng.standard_normal((100_000, dim), dtype=np.float32)
        qs = rng.standard_normal((32, dim), dtype=np.float32)
        dense_cosine_topk(sub, qs[:1], k)
        t0 = time.perf_counter()
        for i in range(32):
            dense_cosine_topk(sub, qs[i:i + 1], k)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 32 / dt / 10.0, "unit": "queries/s over 1 M rows (scaled x 1/10 from 100 k rows)",
                               "scan_GBps": 32 * sub.nbytes / dt / 1e9, "cores": _host_cores(), "kind": "port",
                               "sample": "100 k of 1 M rows x 32 queries, one call per query (numpy/BLAS exact scan)"}
    return out


def sec_splade(ctx, peaks, rank, world, device, cpu: bool, n_chunks: int):
    """BASELINE configs[1], encode side: SPLADE encode of 256-token chunks (BERT-base MLM, 12 layers).  Every rank
    encodes its own ``n_chunks`` (weak scaling, no collective)."""
    import torch
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    out = {}
    timed = Timer(ctx, device)
    bspec = BertSpec()
    w = make_bert_mlm_weights(1002, bspec)
    Lc = 256
    rng = np.random.default_rng(1002 + rank)
    ids = rng.integers(1000, bspec.vocab_size, size=(n_chunks, Lc), dtype=np.int32)
    ids[:, 0], ids[:, -1] = bspec.cls_id, bspec.sep_id
    cu = (np.arange(n_chunks + 1) * Lc).astype(np.int32)
    for prec in ("fast", "precise"):
        nc = n_chunks if prec == "fast" else min(n_chunks, 2048)
        enc = _native.Encoder(ctx, _native.ENC_BERT_MLM, w, bspec.layers, bspec.vocab_size, max_tokens=65536,
                              precision=prec)
        ids_d = torch.from_numpy(ids[:nc].reshape(-1)).to(device)
        dense_d = torch.empty(nc, bspec.vocab_size, dtype=torch.float32, device=device)
        fn = lambda: enc.splade_forward_device(ids_d, cu[:nc + 1], dense_d)   # noqa: E731
        ms = _all_max(timed(fn, iters=2, warm=1), device, world)
        pr = timed.profiled(fn, iters=1)
        tf = FLOPS_PER_SPLADE_CHUNK * nc / ms / 1e9
        rec = {"chunks_per_s": world * nc / ms * 1e3, "ms": ms, "chunks_per_gpu": nc, "n_gpus": world,
               "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (BERT epilogues incl. EPI_SPLADE)",
                            "achieved": tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s algorithmic, whole encode",
                            "frac": tf / peaks["bf16_sustained"],
                            "share_of_encode": {k: v["ms"] / ms for k, v in pr.items() if v["launches"]}}}
        if prec == "fast":
            t0 = time.perf_counter()
            csr = enc.splade_forward(ids.reshape(-1), cu)           # host ids -> host CSR (the provider's path)
            dt = time.perf_counter() - t0
            rec["e2e"] = {"value": n_chunks / dt, "unit": "chunks/s", "h2d_bytes": int(ids.nbytes),
                          "d2h_bytes": int(csr["indices"].nbytes + csr["values"].nbytes + csr["indptr"].nbytes),
                          "path": "vrag_splade_forward(host ids) -> host CSR"}
            rec["mean_nnz"] = float(np.diff(csr["indptr"]).mean())
        out[prec] = rec
        enc.close()
    if cpu:
        import torch as _t
        _t.set_num_threads(_host_cores())
        from oracle.bert_splade import splade_encode, to_dicts_embed_batch
        ns = 64
        seqs = [ids[i].astype(np.int64) for i in range(ns)]
        t0 = time.perf_counter()
        to_dicts_embed_batch(splade_encode(w, seqs, bspec, batch_size=32))
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": ns / dt, "unit": "chunks/s", "cores": _host_cores(), "kind": "port",
                               "sample": f"{ns} of {n_chunks} chunks @256 tokens, batches of 32, fp32 BertForMaskedLM-equivalent + "
                                         "SPLADE pool + the reference's np.nonzero dict loop (embedding_providers.py:148-166)"}
    return out


def sec_sparse(ctx, peaks, device, cpu: bool, big_docs: int):
    """BASELINE configs[1], search side: sparse inner-product top-10, 1 k queries over 10 k docs (L2-resident corpus) and
    the HBM-resident variant SURVEY.md 8d asks for (``big_docs`` documents, 64 queries)."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import csr_to_dicts, make_sparse_rows, make_sparse_rows_device
    out = {}
    timed = Timer(ctx, device)
    V = 30522
    ip, ixs, vl = make_sparse_rows(10000, seed=1002)
    qip, qix, qvl = make_sparse_rows(1000, seed=2002, query=True)
    sx = _native.Index(ctx, _native.INDEX_SPARSE_IP, V)
    sx.add_sparse(ip, ixs, vl)
    sx.search_sparse(qip, qix, qvl, 10)   # warm-up at the timed batch size (first call of a size allocates its buffers)
    dt = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        sx.search_sparse(qip, qix, qvl, 10)
        dt = min(dt, time.perf_counter() - t0)
    pr = timed.profiled(lambda: sx.search_sparse(qip, qix, qvl, 10))
    passes = (1000 + 31) // 32   # 32 queries per corpus pass (all passes of a selection group share one launch)
    bytes_pass = 8 * int(ip[-1]) + 8 * (len(ip))
    # the scan gathers one 128-byte row of the dense query table (32 queries) per stored term (L1 / L2 hits): what bounds
    # it is the SMs' load-return path, 128 B/clk/SM (ncu: l1tex data-pipe wavefronts, profiles/r2b_sparse_scan_*), not the
    # 8 bytes per term streamed from HBM.
    import torch
    l2_peak = torch.cuda.get_device_properties(0).multi_processor_count * 128 * 1.965e9 / 1e9
    gather_pass = (8 + 128) * int(ip[-1])
    out["docs10k_q1000"] = {
        "e2e": {"value": 1000 / dt, "unit": "queries/s", "us_per_query": dt / 1000 * 1e6,
                "path": "vrag_index_search_sparse(host CSR queries) -> host ids + scores"},
        "nnz_corpus": int(ip[-1]), "scan_ms": pr["scan"]["ms"], "select_ms": pr["select"]["ms"], "corpus_passes": passes,
        "roofline": {"bound": "sm load-return path", "kernel": "sparse_scan_kernel (32 queries per corpus pass)",
                     "achieved": passes * gather_pass / pr["scan"]["ms"] / 1e6, "peak": l2_peak,
                     "unit": "GB/s into registers (8 B streamed + 128 B gathered per stored term)",
                     "frac": passes * gather_pass / pr["scan"]["ms"] / 1e6 / l2_peak,
                     "peak_source": "128 B/clk/SM x SMs x 1965 MHz",
                     "hbm_GBps": passes * bytes_pass / pr["scan"]["ms"] / 1e6,
                     "note": "the 13.6 MB CSR corpus is L2-resident (SURVEY.md 8d) and 32 passes take 0.6 ms: "
                             "latency-bound at this size; see docs1M for the HBM-resident corpus"}}
    sx.close()
    if big_docs > 0:
        bip, bix, bvl = make_sparse_rows_device(big_docs, seed=1002, device=device)
        bx = _native.Index(ctx, _native.INDEX_SPARSE_IP, V)
        step = 250_000
        for a in range(0, big_docs, step):
            bx.add_sparse(bip[a:min(big_docs, a + step) + 1], bix, bvl)
        nq = 64
        bx.search_sparse(qip[:nq + 1], qix, qvl, 10)
        dt = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            bx.search_sparse(qip[:nq + 1], qix, qvl, 10)
            dt = min(dt, time.perf_counter() - t0)
        pr = timed.profiled(lambda: bx.search_sparse(qip[:nq + 1], qix, qvl, 10))
        passes = (nq + 31) // 32
        bytes_pass = 8 * int(bip[-1]) + 8 * (big_docs + 1)
        gbs = passes * bytes_pass / pr["scan"]["ms"] / 1e6
        l2_gbs = passes * (8 + 128) * int(bip[-1]) / pr["scan"]["ms"] / 1e6
        out["docs1M_q64"] = {
            "docs": big_docs, "nnz_corpus": int(bip[-1]), "queries": nq, "queries_per_pass": nq / max(passes, 1),
            "e2e": {"value": nq / dt, "unit": "queries/s", "path": "vrag_index_search_sparse(host CSR queries)"},
            "scan_ms": pr["scan"]["ms"], "select_ms": pr["select"]["ms"],
            "roofline": {"bound": "sm load-return path", "kernel": "sparse_scan_kernel (32 queries per corpus pass)",
                         "achieved": l2_gbs, "peak": l2_peak,
                         "unit": "GB/s into registers (8 B streamed + 128 B gathered per stored term)",
                         "frac": l2_gbs / l2_peak, "peak_source": "128 B/clk/SM x SMs x 1965 MHz; ncu shows the l1tex data "
                         "pipe at 74 % (the shuffles that distribute indices / values share it)",
                         "hbm_GBps": gbs, "hbm_frac": gbs / peaks["hbm_gbs"], "hbm_bytes_per_launch": bytes_pass,
                         "avg_launch_ms": pr["scan"]["ms"] / max(passes, 1)}}
        bx.close()
    if cpu:
        from oracle.flat_topk import sparse_ip_topk
        qd = csr_to_dicts(qip[:33], qix, qvl)
        sparse_ip_topk(ip, ixs, vl, V, qd[:1], 10)
        t0 = time.perf_counter()
        for q in qd:
            sparse_ip_topk(ip, ixs, vl, V, [q], 10)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": len(qd) / dt, "unit": "queries/s over 10 k docs", "cores": _host_cores(), "kind": "port",
                               "sample": "32 of 1 k queries, one call per query (scipy CSR exact inner product, "
                                         "milvus_base.py:250-259 control flow)"}
    return out


def sec_heads(ctx, rank, world, device, cpu: bool):
    """the reference's two further model heads on the same kernels.  Host ids in -> host scores out
    through the C ABI (the calls are end to end by construction); every rank scores its own batch (no collective)."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import (BertSpec, ModernBertSpec, make_cross_encoder_weights, make_qa_model_weights)
    out = {}
    rng = np.random.default_rng(1004 + rank)
    # cross-encoder reranker (rerankers.py:109-134): BERT-base pair classifier, 2 048 (query, chunk) pairs of 256 tokens
    bspec = BertSpec()
    w = make_cross_encoder_weights(1004, bspec)
    npairs, L, lq = 2048, 256, 32
    ids = rng.integers(1000, bspec.vocab_size, size=(npairs, L), dtype=np.int32)
    ids[:, 0], ids[:, lq - 1], ids[:, -1] = bspec.cls_id, bspec.sep_id, bspec.sep_id
    types = np.zeros((npairs, L), np.int32)
    types[:, lq:] = 1
    cu = (np.arange(npairs + 1) * L).astype(np.int32)
    for prec in ("fast", "precise"):
        enc = _native.Encoder(ctx, _native.ENC_BERT_CLS, w, bspec.layers, bspec.vocab_size, max_tokens=65536, precision=prec)
        enc.rerank_forward(ids.reshape(-1), types.reshape(-1), cu)
        dt = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            sc = enc.rerank_forward(ids.reshape(-1), types.reshape(-1), cu)
            dt = min(dt, time.perf_counter() - t0)
        dt = _all_max(dt, device, world)
        out.setdefault("reranker_bert_base_256tok", {})[prec] = {
            "pairs_per_s": world * npairs / dt, "ms": dt * 1e3, "pairs_per_gpu": npairs, "n_gpus": world,
            "path": "vrag_rerank_forward(host ids + token types) -> host logits", "score_std": float(sc.std())}
        enc.close()
    # legacy QAModel sentence classifier (extractor_models/model.py:59-117): ModernBERT-base + sentence mean-pool + Linear;
    # 512 (question + document) sequences of 512 tokens, 12 sentences of 40 tokens each
    mspec = ModernBertSpec()
    wq = make_qa_model_weights(1001, mspec)
    nseq, Ls, ns, sl = 512, 512, 12, 40
    qids = rng.integers(1000, mspec.vocab_size, size=(nseq, Ls), dtype=np.int32)
    qids[:, 0], qids[:, -1] = mspec.cls_id, mspec.sep_id
    qcu = (np.arange(nseq + 1) * Ls).astype(np.int32)
    sip = (np.arange(nseq + 1) * ns).astype(np.int32)
    s0 = np.tile(30 + np.arange(ns) * sl, nseq).astype(np.int32)
    s1 = (s0 + sl - 1).astype(np.int32)
    for prec in ("fast", "precise"):
        enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_SENT, wq, mspec.layers, mspec.vocab_size, max_tokens=131072,
                              precision=prec)
        enc.sentence_forward(qids.reshape(-1), qcu, sip, s0, s1)
        dt = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            lg = enc.sentence_forward(qids.reshape(-1), qcu, sip, s0, s1)
            dt = min(dt, time.perf_counter() - t0)
        dt = _all_max(dt, device, world)
        out.setdefault("qa_model_sentences_512tok", {})[prec] = {
            "documents_per_s": world * nseq / dt, "sentences_per_s": world * nseq * ns / dt, "ms": dt * 1e3,
            "documents_per_gpu": nseq, "n_gpus": world,
            "path": "vrag_sentence_forward(host ids + sentence boundaries) -> host logits [n_sentences, 2]",
            "logit_std": float(lg.std())}
        enc.close()
    if cpu:
        import torch as _t
        _t.set_num_threads(_host_cores())
        from oracle.heads import cross_encoder_scores, qa_sentence_logits
        n1 = 16
        t0 = time.perf_counter()
        cross_encoder_scores(w, [ids[i] for i in range(n1)], [types[i] for i in range(n1)], bspec)
        dt = time.perf_counter() - t0
        out["reranker_bert_base_256tok"]["cpu_baseline"] = {
            "value": n1 / dt, "unit": "pairs/s", "cores": _host_cores(), "kind": "port",
            "sample": f"{n1} of {npairs} pairs, one fp32 forward per pair (CrossEncoder.predict order, rerankers.py:120-126)"}
        n2 = 4
        t0 = time.perf_counter()
        for i in range(n2):
            qa_sentence_logits(wq, qids[i], [(int(a), int(b)) for a, b in zip(s0[:ns], s1[:ns])], mspec)
        dt = time.perf_counter() - t0
        out["qa_model_sentences_512tok"]["cpu_baseline"] = {
            "value": n2 / dt, "unit": "documents/s", "cores": _host_cores(), "kind": "port",
            "sample": f"{n2} of {nseq} documents, one fp32 forward per document (extractors.py:246-268)"}
    return out


def sec_rag(ctx, rank, world, device, cpu: bool, n_chunks: int, n_queries: int):
    """BASELINE configs[4]: SPLADE retrieve top-20 + span extraction for ``n_queries`` questions over ``n_chunks`` chunks
    of 256 tokens, strings in / verbatim span strings out through the plugin classes.  N ranks: the index is built data-
    parallel (each rank encodes its slice), vectors are sharded (ShardedB200VectorStore: one all-gather per search), the
    questions are split over the ranks for extraction (pipeline.rag_query_batch's scheme)."""
    import torch.distributed as dist
    from verbatim_rag_b200 import B200SpanExtractor, B200SpladeProvider
    from verbatim_rag_b200.distributed import shard_bounds
    from verbatim_rag_b200.pipeline import index_query_batch
    from verbatim_rag_b200.sharded_store import ShardedB200VectorStore
    dev = f"cuda:{device.index}"
    k = 20
    prov = B200SpladeProvider("synthetic:1002", device=dev)
    ext = B200SpanExtractor("synthetic:1001", device=dev, max_tokens=131072)
    btok = prov._te.tokenizer
    rng = np.random.default_rng(1005)
    chunks = [btok.make_text(rng, 256) for _ in range(n_chunks)]
    questions = [btok.make_question(rng, int(rng.integers(12, 21))) for _ in range(n_queries)]
    ids = [f"c{i:07d}" for i in range(n_chunks)]
    store = ShardedB200VectorStore(enable_dense=False, enable_sparse=True, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
    barrier()
    t0 = time.perf_counter()
    store.add_texts(ids, chunks, chunks, [{} for _ in chunks], sparse_provider=prov)
    barrier()
    t_index = time.perf_counter() - t0
    index = types.SimpleNamespace(vector_store=store, sparse_provider=prov, dense_provider=None)
    lo, hi = shard_bounds(n_queries, rank, world)

    def run(qs, a, b):
        found = index_query_batch(index, qs, k=k)                        # replicated query encode + sharded search
        mine = ext.extract_spans_batch(qs[a:b], found[a:b])               # data-parallel over the ranks
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            mine = [x for p in parts for x in p]
        return found, mine

    w = min(n_queries, 4 * world)
    run(questions[:w], *shard_bounds(w, rank, world))
    best = None
    for _ in range(2):   # the first full-size batch grows pinned / device staging buffers; report the steady state
        barrier()
        t1 = time.perf_counter()
        found = index_query_batch(index, questions, k=k)
        t2 = time.perf_counter()
        mine = ext.extract_spans_batch(questions[lo:hi], found[lo:hi])
        t3 = time.perf_counter()
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            spans = [x for p in parts for x in p]
        else:
            spans = mine
        barrier()
        t4 = time.perf_counter()
        rec = (t4 - t1, t2 - t1, t3 - t2, t4 - t3)
        best = rec if best is None or rec[0] < best[0] else best
    n_ext = sum(len(r) for r in found)
    out = {"chunks": n_chunks, "chunk_tokens": 256, "queries": n_queries, "k": k, "n_gpus": world, "extractions": n_ext,
           "queries_per_s": n_queries / best[0], "extractions_per_s": n_ext / best[0],
           "stage_ms": {"retrieve (query encode + sharded sparse top-20)": best[1] * 1e3,
                        "extract (tokenise pairs + 22-layer forward + spans), slowest rank's share": best[2] * 1e3,
                        "gather responses": best[3] * 1e3},
           "index_build_chunks_per_s": n_chunks / t_index,
           "spans": int(sum(len(v) for d in spans for v in d.values())),
           "path": "strings in -> verbatim span strings out: B200SpladeProvider -> ShardedB200VectorStore -> "
                   "B200SpanExtractor via pipeline.index_query_batch / extract_spans_batch (the steps of "
                   "VerbatimRAG.query, core.py:210-277, minus template filling)"}
    ext._workers.close()
    if cpu and world == 1:
        import torch as _t
        _t.set_num_threads(_host_cores())
        from oracle.plugins import OracleFlatStore, OracleSpanExtractor, OracleSpladeProvider
        from verbatim_rag_b200.synthetic import (BertSpec, ModernBertSpec, SyntheticTokenizer, make_bert_mlm_weights,
                                                 make_modernbert_weights)
        bspec, mspec = BertSpec(), ModernBertSpec(layers=LAYERS)
        oprov = OracleSpladeProvider(make_bert_mlm_weights(1002, bspec), btok, bspec)
        oext = OracleSpanExtractor({k2: _t.from_numpy(v) for k2, v in make_modernbert_weights(1001, mspec).items()},
                                   SyntheticTokenizer("modernbert"), mspec)
        nc, nq_c, kc = 64, 2, 8
        ostore = OracleFlatStore(enable_dense=False, enable_sparse=True)
        t0 = time.perf_counter()
        ostore.add_vectors(ids[:nc], None, oprov.embed_batch(chunks[:nc]), chunks[:nc], chunks[:nc], [{} for _ in range(nc)])
        t_idx = time.perf_counter() - t0
        t0 = time.perf_counter()
        n_e = 0
        for q in questions[:nq_c]:      # the reference's loop: one query at a time (core.py:210-277)
            res = ostore.query(sparse_query=oprov.embed_text(q), top_k=kc, search_type="sparse")
            n_e += len(oext.extract_spans(q, res))
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_e / dt, "unit": "extractions/s (retrieve + extract, per query)",
                               "queries_per_s_at_k20": (nq_c / dt) * kc / k, "index_build_chunks_per_s": nc / t_idx,
                               "cores": _host_cores(), "kind": "port",
                               "sample": f"{nc} chunks indexed, {nq_c} queries x top-{kc}: OracleSpladeProvider -> OracleFlatStore "
                                         "-> OracleSpanExtractor, one query at a time"}
    return out


def run_gpu_arm(args, out):
    import torch
    import torch.distributed as dist
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = _peaks()
    ctx = _native.default_context(local)
    spec = ModernBertSpec(layers=LAYERS)
    weights = make_modernbert_weights(1001, spec)
    nseq = args.seqs_per_step
    ids_h, cu, tok_cs, tok_ce, ctx_indptr = make_batch(nseq, 1003 + rank)
    T = int(cu[-1])
    ids_d = torch.from_numpy(ids_h).to(device)
    probs_d = torch.empty(T, dtype=torch.float32, device=device)
    ids_pin = torch.from_numpy(ids_h).pin_memory().numpy()
    st = torch.cuda.ExternalStream(ctx.stream, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        ctx.sync()

    def measure(enc, steps, warmup, profile):
        """-> (ms per step of the device-resident step, max over ranks; profile; launches; e2e seconds per step)"""
        for _ in range(warmup):
            enc.span_forward_device(ids_d, cu, probs_d)
        barrier()
        l0 = ctx.launches
        if profile:
            ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            enc.span_forward_device(ids_d, cu, probs_d)
        e1.record(st)
        barrier()
        ms_total = e0.elapsed_time(e1)
        prof = ctx.profile_read() if profile else None
        if profile:
            ctx.profile(False)
        launches = ctx.launches - l0
        ms_step = _all_max(ms_total, device, world) / steps

        # e2e: host buffers through the C ABI (H2D + forward + D2H + span post-processing)
        def step_host():
            p = enc.span_forward(ids_pin, cu)
            c0 = Q_LEN + 2
            p_ctx = np.ascontiguousarray(p.reshape(nseq, SEQ_LEN)[:, c0:c0 + CTX_LEN]).reshape(-1)
            return _native.spans_from_probs(p_ctx, tok_cs, tok_ce, ctx_indptr, 0.2, 30, 20)
        step_host()
        barrier()
        t0 = time.perf_counter()
        nsp = 0
        reps = max(1, steps // 2)
        for _ in range(reps):
            nsp = len(step_host()["ctx"])
        ctx.sync()
        e2e_s = _all_max((time.perf_counter() - t0) / reps, device, world)
        return ms_step, ms_total, prof, launches, e2e_s, nsp

    # ---- fast mode = the headline ---------------------------------------------------------------------------------
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, weights, spec.layers, spec.vocab_size,
                          max_tokens=args.max_tokens, precision="fast")
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, ms_total, prof, launches, e2e_s, nsp = measure(enc, args.steps, args.warmup, True)
    clocks = sampler.stop()
    value = world * nseq / (ms_step / 1e3)
    e2e_value = world * nseq / e2e_s
    modes = {"fast": {"value": value, "unit": UNIT, "ms_per_step": ms_step, "e2e": e2e_value, "dtype": "f16",
                      "logit_tolerance": "4.0e-3 (measured 3.3e-3 on the cfg-1 goldens)"}}

    # ---- precise mode: split-precision operands, the north-star tolerance ---------------------------------------------
    if args.precise:
        enc_p = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, weights, spec.layers, spec.vocab_size,
                                max_tokens=args.max_tokens, precision="precise")
        sp2 = ClockSampler(local)
        sp2.start()
        p_ms, _, p_prof, _, p_e2e_s, _ = measure(enc_p, max(1, args.steps // 2), min(args.warmup, 2), True)
        p_clocks = sp2.stop()
        enc_p.close()
        modes["precise"] = {"value": world * nseq / (p_ms / 1e3), "unit": UNIT, "ms_per_step": p_ms,
                            "e2e": world * nseq / p_e2e_s, "dtype": "f16 hi + f16 lo planes, 3 tcgen05.mma per product",
                            "logit_tolerance": "1e-3 (measured 4.7e-5 on the cfg-1 goldens)", "clocks": p_clocks,
                            "tflops_issued": 3 * FLOPS_LINEAR_PER_TOKEN * T * max(1, args.steps // 2)
                                             / (p_prof["gemm"]["ms"] / 1e3) / 1e12 if p_prof["gemm"]["ms"] > 0 else None,
                            "share_of_step": {k: v["ms"] / (p_ms * max(1, args.steps // 2)) for k, v in p_prof.items() if v["launches"]}}

    # ONE table of secondary blocks for every rank.  `collective` blocks synchronise / reduce across the ranks and must run
    # on all of them, in this order; the others run on rank 0 only.  (A block that rank 0 alone entered with a collective
    # inside would leave the job hanging in the final barrier -- tests/test_bench_contract.py checks the table.)
    cpu = world == 1 and args.cpu_baseline
    secondary_blocks = (
        ("cfg4_dense_1Mx768", True, lambda: sec_dense(ctx, peaks, rank, world, device, cpu)),
        ("cfg2_splade_encode_256tok", True, lambda: sec_splade(ctx, peaks, rank, world, device, cpu, args.splade_chunks)),
        ("cfg2_sparse_top10", False, lambda: sec_sparse(ctx, peaks, device, cpu, args.sparse_docs if world == 1 else 0)),
        ("cfg5_rag_e2e", True, lambda: sec_rag(ctx, rank, world, device, cpu, args.rag_chunks, args.rag_queries)),
        ("f4_heads", True, lambda: sec_heads(ctx, rank, world, device, cpu)),
    )
    if rank != 0:
        enc.close()
        if args.secondary:
            for _, collective, fn in secondary_blocks:
                if collective:
                    fn()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    gemm_ms, gemm_n = prof["gemm"]["ms"], prof["gemm"]["launches"]
    gemm_flops = FLOPS_LINEAR_PER_TOKEN * T * args.steps
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    cfg = bench_config()
    cfg.update({"max_tokens_per_pass": args.max_tokens, "tflops_algorithmic": value * FLOPS_PER_EXTRACTION / 1e12 / world})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic", "config": cfg, "clocks": clocks, "modes": modes,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(T * 4 + (nseq + 1) * 4),
                "d2h_bytes_per_step": int(T * 4), "spans_per_step": int(nsp),
                "path": "vrag_span_forward(host ids) + vrag_spans_from_probs (C ABI, host buffers)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "peak_source": peaks["source"] + ", sustained cuBLAS bf16", "traffic": traffic,
                     "traffic_source": traffic_src,
                     "launches": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                     "flops_per_launch_avg": gemm_flops / max(gemm_n, 1),
                     "share_of_step": {k: v["ms"] / (ms_total) for k, v in prof.items() if v["launches"]}},
    }
    enc.close()
    if world == 1 and args.secondary:
        # plugin level: strings in -> verbatim span strings out (B200SpanExtractor.extract_spans_batch) on half a step
        # of fresh texts; host tokenisation (worker processes) and span post-processing overlap the GPU forward
        from verbatim_rag_b200 import B200SpanExtractor
        from verbatim_rag_b200.synthetic import SyntheticTokenizer
        tk = SyntheticTokenizer("modernbert")
        ext = B200SpanExtractor(weights=weights, tokenizer=tk, num_layers=spec.layers, vocab_size=spec.vocab_size,
                                max_tokens=args.max_tokens, device=f"cuda:{local}")
        nq_p = max(2, args.plugin_pairs // CHUNKS_PER_Q)
        qs, rs = make_text_batch(tk, nq_p, 7)
        ext.extract_spans_batch(qs[:40], rs[:40])     # starts the tokeniser workers, grows the staging buffers
        t0 = time.perf_counter()
        res = ext.extract_spans_batch(qs, rs)
        dtp = time.perf_counter() - t0
        line["e2e_plugin_strings"] = {"value": nq_p * CHUNKS_PER_Q / dtp, "unit": UNIT, "pairs": nq_p * CHUNKS_PER_Q,
                                      "spans": int(sum(len(v) for d in res for v in d.values())),
                                      "tokenizer_workers": ext._workers.n, "host_cores": _host_cores(),
                                      "note": "B200SpanExtractor.extract_spans_batch(strings): tokenisation of slices "
                                              "i+1, i+2 and span post-processing of slice i-1 overlap the forward of slice i"}
        ext._workers.close()
        del ext
    if world == 1 and args.cpu_baseline:
        rate, dt, cores, _ = time_cpu_spans(args.cpu_sample, 1, 1)
        line["cpu_baseline"] = cpu_baseline_block(rate, cores, args.cpu_sample)
    if args.secondary:
        sec = {}
        for name, _, fn in secondary_blocks:
            try:
                sec[name] = fn()
            except Exception as exc:  # noqa: BLE001 -- a secondary block must not take the headline line down
                if world > 1:
                    raise               # ... except where the other ranks would wait in a collective forever
                sec[name] = {"error": repr(exc)}
        line["secondary"] = sec
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs-per-step", type=int, default=SEQS_PER_STEP)
    ap.add_argument("--max-tokens", type=int, default=131072)  # pass-size sweep: profiles/README.md
    ap.add_argument("--cpu-sample", "--ref-sample", dest="cpu_sample", type=int, default=64)   # pairs timed by the CPU arm; the same in both arms
    ap.add_argument("--plugin-pairs", type=int, default=2048)  # pairs of the strings-in plugin measurement
    ap.add_argument("--splade-chunks", type=int, default=10000)
    ap.add_argument("--sparse-docs", type=int, default=1_000_000)
    ap.add_argument("--rag-chunks", type=int, default=10000)
    ap.add_argument("--rag-queries", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    ap.add_argument("--no-precise", dest="precise", action="store_false")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything libraries print while the run is in progress (NCCL's version
    # banner, torch warnings) goes to stderr instead
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(out_fd, "w")
    if args.impl == "reference":
        run_reference_arm(args, out)
    else:
        run_gpu_arm(args, out)


if __name__ == "__main__":
    main()
