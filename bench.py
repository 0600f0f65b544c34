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
  e2e          the same step through the C ABI with HOST buffers: ids H2D, forward, probs D2H, span post-processing
  roofline     tensor-core GEMM: algorithmic FLOPs / summed CUDA-event kernel time, vs MEASURED_PEAKS.json
  cpu_baseline the oracle port timed on this box's host cores (bounded sample; rank 0, N=1 only)
  secondary    top-k scan (BASELINE 'top-k GB/s') and SPLADE encode numbers, same run
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "(question,512-tok chunk) span-extractions/sec"
UNIT = "extractions/s"
SEQ_LEN, Q_LEN = 512, 29
SEQS_PER_STEP = 256 * 16
LAYERS = 22
FLOPS_LINEAR_PER_TOKEN = 221.78e6          # all Linear layers incl. head
FLOPS_PER_EXTRACTION = 122.65e9            # linears + windowed attention at L = 512
WORKLOAD = "ModernBERT-v2 span extraction: 256-query batch x 16 retrieved chunks @512 tok (configs[2])"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"source": "measured (MEASURED_PEAKS.json)", "hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"])}
    return {"source": "fallback (B200_PROFILING.md)", "hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0}


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
    n_ctx = SEQ_LEN - Q_LEN - 3
    tok_cs = np.tile(np.arange(n_ctx, dtype=np.int32) * 7, nseq)
    tok_ce = tok_cs + 6
    ctx_indptr = np.arange(nseq + 1, dtype=np.int64) * n_ctx
    return ids.reshape(-1), cu, tok_cs, tok_ce, ctx_indptr


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
# CPU arm: the reference-shaped path (oracle port), batch-1 forward per chunk like extractors.py:207-221
# --------------------------------------------------------------------------------------------------------------
def cpu_extract(weights, ids, cu, tok_cs, tok_ce, ctx_indptr, nseq):
    import torch
    from oracle.highlighter import spans_from_token_probs
    from oracle.modernbert import modernbert_forward, relevant_prob
    n_ctx = SEQ_LEN - Q_LEN - 3
    nspans = 0
    for i in range(nseq):
        seq = ids[cu[i]:cu[i + 1]].astype(np.int64)[None]
        p = relevant_prob(modernbert_forward(weights, seq).numpy()[0])[Q_LEN + 2:Q_LEN + 2 + n_ctx]
        offs = list(zip(tok_cs[:n_ctx].tolist(), tok_ce[:n_ctx].tolist()))
        nspans += len(spans_from_token_probs("x" * (7 * n_ctx), p, offs, 0.2, 30, 20))
    return nspans


def time_cpu(nseq_sample: int, steps: int, warmup: int):
    import torch
    try:   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core this process may run on
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        pass
    from verbatim_rag_b200.synthetic import make_modernbert_weights
    weights = {k: torch.from_numpy(v) for k, v in make_modernbert_weights(1001).items()}
    ids, cu, tcs, tce, cip = make_batch(nseq_sample, 1003)
    for _ in range(warmup):
        cpu_extract(weights, ids, cu, tcs, tce, cip, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_extract(weights, ids, cu, tcs, tce, cip, nseq_sample)
    dt = (time.perf_counter() - t0) / steps
    return nseq_sample / dt, dt, torch.get_num_threads()


def run_reference_arm(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample   # ~5-10 s of host work per step
    rate, dt, cores = time_cpu(sample, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "seq_len": SEQ_LEN, "layers": LAYERS},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {SEQS_PER_STEP} sequences per step, batch-1 forward per chunk "
                                   "(reference control flow, oracle ModernBERT fp32 on torch CPU) + span post-processing"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def secondary_metrics(ctx, peaks, rank, world, device):
    """Top-k scan + SPLADE encode numbers of the same run (bounded: ~10 s)."""
    import torch
    from verbatim_rag_b200 import _native
    out = {}
    st = torch.cuda.ExternalStream(ctx.stream, device=device)

    def timed(fn, iters=3):
        fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters):
            fn()
        e1.record(st)
        ctx.sync()
        return e0.elapsed_time(e1) / iters

    # dense cosine top-10, 1M x 768 fp32 corpus row-sharded over the ranks (configs[3])
    n_total, dim, k = 1_000_000, 768, 10
    from verbatim_rag_b200.distributed import shard_bounds
    lo, hi = shard_bounds(n_total, rank, world)
    g = torch.Generator(device=device).manual_seed(1004 + rank)
    ix = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    ix.set_id_base(lo)
    ix.add_dense(torch.randn(hi - lo, dim, device=device, generator=g))
    gq = torch.Generator(device=device).manual_seed(2004)
    for nq in (1, 16, 1000):  # per-query API call; one tensor-core pass; the configs[3] batch (63 passes of 16)
        q = torch.randn(nq, dim, device=device, generator=gq)
        ids_o = torch.empty(nq, k, dtype=torch.int64, device=device)
        s_o = torch.empty(nq, k, dtype=torch.float32, device=device)
        iters = 3 if nq <= 16 else 1
        ms = timed(lambda: ix.search_dense_device(q, nq, k, ids_o, s_o), iters=iters)
        ctx.profile(True)
        for _ in range(iters):
            ix.search_dense_device(q, nq, k, ids_o, s_o)
        pr = ctx.profile_read()
        ctx.profile(False)
        passes = pr["scan"]["launches"] / iters                                # corpus passes per search call
        pass_bytes = (hi - lo) * dim * 4
        scan_ms = pr["scan"]["ms"] / max(pr["scan"]["launches"], 1)          # per corpus pass (CUDA events)
        scan_gbs = pass_bytes / scan_ms / 1e6 if scan_ms > 0 else 0.0
        out[f"dense_top{k}_q{nq}"] = {
            "ms": ms, "queries_per_s": nq / ms * 1e3, "rows_per_gpu": hi - lo, "corpus_passes": passes,
            "queries_per_pass": nq / passes if passes else 0,
            "search_GBps_per_gpu": passes * pass_bytes / ms / 1e6,             # whole search incl. select / rescore
            "roofline": {"bound": "hbm", "kernel": "dense_scan_tma_kernel (fp32 FMA)" if nq < 5 else
                         "dense_scan_tc_kernel (split-tf32 tcgen05)", "achieved": scan_gbs,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": scan_gbs / peaks["hbm_gbs"],
                         "bytes_per_launch": pass_bytes, "avg_launch_ms": scan_ms},
            "select_finish_ms": pr["select"]["ms"] / iters}
    if world > 1:
        # configs[3] as stated: corpus row-sharded over the ranks, queries replicated, per-rank top-k, ONE NCCL all_gather
        # of [Q, k] (fp64 score, int64 global id) + the device merge -> the global top-10 on every rank
        import torch.distributed as dist
        from verbatim_rag_b200.distributed import sharded_search_dense
        q = torch.randn(1000, dim, device=device, generator=gq)
        sharded_search_dense(ix, q, k)
        dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        gi, gs, _ = sharded_search_dense(ix, q, k)
        torch.cuda.synchronize(device)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["dense_top10_q1000_global"] = {
            "ms": float(dt.item()) * 1e3, "queries_per_s": 1000 / float(dt.item()), "ranks": world,
            "rows_total": n_total, "corpus_GBps_all_gpus": 63 * n_total * dim * 4 / float(dt.item()) / 1e9,
            "collective": "one all_gather of [1000, 10] x (f64, i64) per rank + vrag_topk_merge",
            "ids_sorted_by_score": bool((gs[:, :-1] >= gs[:, 1:]).all().item())}
    ix.close()

    if rank == 0:
        # configs[1]: SPLADE encode of 256-token chunks (BERT-base MLM, 12 layers) + sparse-dot top-10 over 10k docs
        from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights, make_sparse_rows
        bspec = BertSpec()
        enc = _native.Encoder(ctx, _native.ENC_BERT_MLM, make_bert_mlm_weights(1002, bspec), bspec.layers,
                              bspec.vocab_size, max_tokens=65536)
        nchunk, Lc = 1024, 256
        rng = np.random.default_rng(1002)
        ids = rng.integers(1000, bspec.vocab_size, size=(nchunk, Lc), dtype=np.int32)
        ids[:, 0], ids[:, -1] = bspec.cls_id, bspec.sep_id
        cu = (np.arange(nchunk + 1) * Lc).astype(np.int32)
        ids_d = torch.from_numpy(ids.reshape(-1)).to(device)
        dense_d = torch.empty(nchunk, bspec.vocab_size, dtype=torch.float32, device=device)
        ms = timed(lambda: enc.splade_forward_device(ids_d, cu, dense_d), iters=2)
        t0 = time.perf_counter()
        csr = enc.splade_forward(ids.reshape(-1), cu)           # host ids -> host CSR (the provider's path)
        e2e_s = time.perf_counter() - t0
        out["splade_encode_256tok"] = {"chunks_per_s": nchunk / ms * 1e3, "ms": ms, "chunks": nchunk,
                                       "tflops_algorithmic": 58.21e9 * nchunk / ms / 1e9,
                                       "e2e_chunks_per_s_host_ids_to_host_csr": nchunk / e2e_s,
                                       "mean_nnz": float(np.diff(csr["indptr"]).mean())}
        enc.close()
        ip, ixs, vl = make_sparse_rows(10000, seed=1002)
        qip, qix, qvl = make_sparse_rows(64, seed=2002, query=True)
        sx = _native.Index(ctx, _native.INDEX_SPARSE_IP, bspec.vocab_size)
        sx.add_sparse(ip, ixs, vl)
        sx.search_sparse(qip, qix, qvl, 10)
        t0 = time.perf_counter()
        for _ in range(5):
            sx.search_sparse(qip, qix, qvl, 10)
        dt = (time.perf_counter() - t0) / 5
        out["sparse_top10_10k_docs"] = {"us_per_query_e2e_host": dt / 64 * 1e6, "queries": 64, "nnz_corpus": int(ip[-1]),
                                        "note": "12.8 MB CSR corpus is L2 resident: latency-bound, not an HBM roofline"}
        sx.close()
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
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, make_modernbert_weights(1001, spec), spec.layers,
                          spec.vocab_size, max_tokens=args.max_tokens)
    nseq = args.seqs_per_step
    ids_h, cu, tok_cs, tok_ce, ctx_indptr = make_batch(nseq, 1003 + rank)
    T = int(cu[-1])
    ids_d = torch.from_numpy(ids_h).to(device)
    probs_d = torch.empty(T, dtype=torch.float32, device=device)
    st = torch.cuda.ExternalStream(ctx.stream, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        ctx.sync()

    def step_device():
        enc.span_forward_device(ids_d, cu, probs_d)

    # ---- value: inputs resident in HBM, CUDA events on the launching stream ------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step_device()
    e1.record(st)
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - l0
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * nseq / (ms_step / 1e3)

    # ---- e2e: host buffers through the C ABI (H2D + forward + D2H + span post-processing) -----------------------
    ids_pin = torch.from_numpy(ids_h).pin_memory().numpy()
    def step_host():
        p = enc.span_forward(ids_pin, cu)
        c0 = Q_LEN + 2
        p_ctx = np.ascontiguousarray(p.reshape(nseq, SEQ_LEN)[:, c0:c0 + (SEQ_LEN - Q_LEN - 3)]).reshape(-1)
        return _native.spans_from_probs(p_ctx, tok_cs, tok_ce, ctx_indptr, 0.2, 30, 20)
    step_host()
    barrier()
    t0 = time.perf_counter()
    nsp = 0
    for _ in range(max(1, args.steps // 2)):
        nsp = len(step_host()["ctx"])
    ctx.sync()
    e2e_s = (time.perf_counter() - t0) / max(1, args.steps // 2)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * nseq / float(t.item())

    if rank != 0:
        if args.secondary:
            secondary_metrics(ctx, peaks, rank, world, device)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    gemm_ms, gemm_n = prof["gemm"]["ms"], prof["gemm"]["launches"]
    gemm_flops = FLOPS_LINEAR_PER_TOKEN * T * args.steps
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "seqs_per_step_per_gpu": nseq, "seq_len": SEQ_LEN, "layers": LAYERS,
                   "precision": "fp16 tensor-core operands, fp32 accumulate, two-plane residual stream (fp16 + e5m2, >= 14 bits)",
                   "max_tokens_per_pass": args.max_tokens,
                   "l2": "working set per step (>= 11.5 KB of activations per token, %.1f GB) exceeds the 126 MB L2"
                         % (11.5e3 * T / 1e9),
                   "tflops_algorithmic": value * FLOPS_PER_EXTRACTION / 1e12 / world},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(T * 4 + (nseq + 1) * 4),
                "d2h_bytes_per_step": int(T * 4), "spans_per_step": int(nsp),
                "path": "vrag_span_forward(host ids) + vrag_spans_from_probs (C ABI, host buffers)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "peak_source": peaks["source"] + ", sustained cuBLAS bf16", "traffic": traffic,
                     "launches": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                     "flops_per_launch_avg": gemm_flops / max(gemm_n, 1),
                     "share_of_step": {k: v["ms"] / (ms_total) for k, v in prof.items() if v["launches"]}},
    }
    if world == 1 and args.secondary:
        # plugin level: strings in -> verbatim span strings out (B200SpanExtractor.extract_spans_batch), small sample.
        # Host tokenisation (tokenizers library, word-level synthetic vocab) dominates this number, not the GPU.
        from verbatim_rag_b200 import B200SpanExtractor
        from verbatim_rag_b200.synthetic import SyntheticTokenizer
        tk = SyntheticTokenizer("modernbert")
        ext = B200SpanExtractor.__new__(B200SpanExtractor)
        ext.tokenizer, ext.threshold, ext.min_span_chars, ext.merge_gap_chars = tk, 0.2, 30, 20
        ext.max_length, ext.doc_stride, ext._enc, ext._ctx, ext.pipeline_pairs = 8192, 256, enc, ctx, 512
        ext._lock = threading.Lock()
        prng = np.random.default_rng(7)
        nq_p, nchunk_p = 32, 32

        class _R:
            def __init__(self, t):
                self.text = t
        qs = [tk.make_question(prng, Q_LEN) for _ in range(nq_p)]
        rs = [[_R(tk.make_text(prng, SEQ_LEN - Q_LEN - 3)) for _ in range(nchunk_p)] for _ in range(nq_p)]
        ext.extract_spans_batch(qs[:2], rs[:2])
        t0 = time.perf_counter()
        res = ext.extract_spans_batch(qs, rs)
        dtp = time.perf_counter() - t0
        line["e2e_plugin_strings"] = {"value": nq_p * nchunk_p / dtp, "unit": UNIT, "pairs": nq_p * nchunk_p,
                                      "spans": int(sum(len(v) for d in res for v in d.values())),
                                      "note": "extract_spans_batch(strings): host tokenisation of slice i+1 overlaps the GPU forward of slice i"}
    if world == 1 and args.cpu_baseline:
        rate, dt, cores = time_cpu(args.cpu_sample, 1, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_sample} of {SEQS_PER_STEP} sequences, batch-1 forward per chunk "
                                          "(reference control flow, oracle ModernBERT fp32 on torch CPU)"}
    if args.secondary:
        line["secondary"] = secondary_metrics(ctx, peaks, rank, world, device)
    if world == 1 and args.secondary:
        # configs[4] shape, bounded: SPLADE retrieve top-20 + span extraction through the plugin classes (strings in)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        try:
            import rag_bench
            enc.close()   # the plugins own their encoders
            line["secondary"]["rag_e2e"] = rag_bench.run(device=f"cuda:{local}")
        except Exception as exc:  # noqa: BLE001 -- a secondary block must not take the headline line down
            line["secondary"]["rag_e2e"] = {"error": repr(exc)}
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
    ap.add_argument("--cpu-sample", type=int, default=96)   # ~10-30 s of host work
    ap.add_argument("--ref-sample", type=int, default=24)    # sequences per step of the --impl reference arm
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
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
