"""GPU parity of the split-precision ("precise") mode: every fp16 weight / activation as hi + lo planes, three
tcgen05.mma per product (csrc/gemm.cuh).  This is the mode that meets the north-star tolerance -- span logits within
1e-3 of the reference's fp32 forward (packages/core/verbatim_core/extractors.py:151-157, 213-221 run the model in
fp32) -- so every bar in this file is 1e-3 or tighter."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-3   # BASELINE.json north_star: "span logits ... within 1e-3 fp"
PROB_TOL = 1e-3


def _diag(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_diag.jsonl"), "a") as f:
        f.write(json.dumps(kw, default=float) + "\n")


@pytest.fixture(scope="module")
def ctx():
    from verbatim_rag_b200 import _native
    return _native.default_context(0)


@pytest.mark.parametrize("epi,name", [(10, "f32"), (0, "f16"), (1, "rope_qkv"), (2, "resid_f32"), (3, "geglu")])
@pytest.mark.parametrize("M", [128, 300, 5000, 1001])
def test_split_gemm_matches_simt_reference(ctx, epi, name, M):
    """a_hi w_hi + a_lo w_hi + a_hi w_lo on the tensor cores (hi / lo units alternating in the operand ring) vs the same
    three products summed per k by the SIMT reference; fp16 outputs are compared as hi + lo sums, so the bar is the
    fp32 accumulation-order noise, not an fp16 rounding."""
    N, K = 2304, 768
    diff, ref_max = ctx.selftest_gemm_split(M, N, K, epi)
    _diag(test="gemm_split_selftest", epilogue=name, M=M, max_abs_diff=diff, ref_abs_max=ref_max)
    assert diff <= 2e-5 * max(ref_max, 1.0), (name, diff, ref_max)


@pytest.mark.parametrize("M,N,K", [(4096, 768, 1152), (20000, 768, 768), (129, 768, 3072), (640, 30720, 768)])
def test_split_gemm_encoder_shapes(ctx, M, N, K):
    diff, ref_max = ctx.selftest_gemm_split(M, N, K, 2 if N == 768 else 10)
    _diag(test="gemm_split_shapes", M=M, N=N, K=K, max_abs_diff=diff, ref_abs_max=ref_max)
    assert diff <= 2e-5 * max(ref_max, 1.0), (diff, ref_max)


def _attention_f64(qkv, cu, window):
    T = qkv.shape[0]
    x = qkv.astype(np.float64).reshape(T, 3, 12, 64)
    out = np.zeros((T, 12, 64))
    for s in range(len(cu) - 1):
        a, b = int(cu[s]), int(cu[s + 1])
        L = b - a
        idx = np.arange(L)
        mask = np.abs(idx[:, None] - idx[None, :]) <= window if window >= 0 else np.ones((L, L), bool)
        for h in range(12):
            q, k, v = x[a:b, 0, h], x[a:b, 1, h], x[a:b, 2, h]
            sc = np.where(mask, q @ k.T / 8.0, -np.inf)
            sc -= sc.max(axis=1, keepdims=True)
            p = np.exp(sc)
            out[a:b, h] = (p / p.sum(axis=1, keepdims=True)) @ v
    return out.reshape(T, 768)


@pytest.mark.parametrize("window", [-1, 64])
@pytest.mark.parametrize("case", ["moderate", "wide_scores", "rising_max"])
def test_split_attention_vs_float64(ctx, window, case):
    """Split-precision attention on fp32 rows (hi + lo planes carry ~22 bits of them) vs float64 on the same fp32
    values: ragged lengths, score ranges that force the online-softmax rescaling, monotonically rising maxima."""
    rng = np.random.default_rng(11)
    lens = [512, 200, 77, 1, 129, 640]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    T = int(cu[-1])
    qkv = rng.standard_normal((T, 3, 12, 64))
    if case == "wide_scores":
        qkv[:, :2] *= 4.0
    elif case == "rising_max":
        pos = np.concatenate([np.arange(n) for n in lens])[:, None]
        qkv[:, 0, :, 0] = 6.0
        qkv[:, 1, :, 0] = pos * 0.25
    qkv32 = qkv.reshape(T, 2304).astype(np.float32)
    hi = qkv32.astype(np.float16).astype(np.float32)
    planes = hi + (qkv32 - hi).astype(np.float16).astype(np.float32)     # what the kernel actually sees
    ref = _attention_f64(planes, cu, window)
    got = ctx.selftest_attention_split(qkv32, cu, window=window).astype(np.float64)
    err = np.abs(got - ref)
    _diag(test="attention_split_selftest", case=case, window=window, max_abs_err=err.max(),
          ref_abs_max=np.abs(ref).max())
    assert np.isfinite(got).all()
    # wide_scores: |s| reaches ~60 in log2 units, an fp32 rounding of s (2^-24 * 64) is a 4e-6 relative step of p
    assert err.max() <= (1e-4 if case == "wide_scores" else 2e-5) * max(np.abs(ref).max(), 1.0), (case, window, err.max())


def _modernbert_case(layers, lens, seed):
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    spec = ModernBertSpec(layers=layers)
    w = make_modernbert_weights(seed, spec)
    rng = np.random.default_rng(seed + 1)
    seqs = []
    for L in lens:
        s = rng.integers(5, 50279, size=L)
        s[0] = spec.cls_id
        s[-1] = spec.sep_id
        seqs.append(s.astype(np.int64))
    return spec, w, seqs


@pytest.mark.parametrize("use_ref_gemm", [True, False])
def test_modernbert_precise_vs_oracle(ctx, use_ref_gemm, monkeypatch):
    from verbatim_rag_b200 import _native
    from oracle.modernbert import modernbert_forward, modernbert_forward_varlen, relevant_prob
    lens = [150, 200, 333, 512, 64, 129, 7, 1, 700]
    spec, w, seqs = _modernbert_case(4, lens, 1001)
    monkeypatch.setenv("VRAG_GEMM_REFERENCE", "1" if use_ref_gemm else "0")
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=4096,
                          precision="precise")
    ids, cu = _native.Encoder._pack(seqs)
    probs, logits, hidden = enc.debug_span_hidden(ids, cu)
    enc.close()
    ref_logits = np.concatenate(modernbert_forward_varlen(w, seqs, spec, batch=1), axis=0)
    err = np.abs(logits - ref_logits)
    _, hid3 = modernbert_forward(w, seqs[3][None], None, spec, return_hidden=True)
    a, b = cu[3], cu[4]
    layer_err = [float(np.abs(hidden[l, a:b] - hid3[l][0].numpy()).max()) for l in range(spec.layers + 1)]
    perr = np.abs(probs - relevant_prob(ref_logits))
    _diag(test="modernbert_precise_vs_oracle", use_ref_gemm=use_ref_gemm, logit_max_err=float(err.max()),
          logit_rms_err=float(np.sqrt((err ** 2).mean())), prob_max_err=float(perr.max()), layer_max_err=layer_err)
    assert err.max() < 2e-4 and perr.max() < 1e-4, (err.max(), perr.max(), layer_err)


def test_span_extractor_config1_precise_vs_golden():
    """BASELINE config 1 (4 questions x 8 chunks @128 tokens, the full 22-layer model) through
    B200SpanExtractor(precision="precise"): logits within 1e-3 of the committed oracle outputs, spans identical on all
    32 pairs (the golden set's closest token is 3.9e-5 from the threshold; the measured probability error is below
    that, and the assertion on the spans does not depend on luck: any mismatch fails)."""
    import cases
    from verbatim_rag_b200 import B200SpanExtractor
    c = cases.span_cfg1()
    g = np.load(os.path.join(HERE, "golden", "span_cfg1.npz"))
    assert bytes(g["input_hash"]).decode() == c["hash"]
    ext = B200SpanExtractor(weights=c["weights"], tokenizer=c["tokenizer"], num_layers=c["spec"].layers,
                            vocab_size=c["spec"].vocab_size, max_tokens=8192, precision="precise")
    plan = ext._tokenize(c["pairs"])
    probs, logits = ext._enc.span_forward(plan["ids"], plan["cu"], want_logits=True)
    lerr = float(np.abs(logits - g["logits"]).max())
    e = np.exp(g["logits"] - g["logits"].max(axis=1, keepdims=True))
    p_ref = (e[:, 1] / e.sum(axis=1)).astype(np.float32)
    perr = float(np.abs(probs - p_ref).max())
    got = ext.extract_detailed(c["pairs"])
    exp = {i: [] for i in range(len(c["pairs"]))}
    for row in g["spans"].tolist():
        exp[row[0]].append(tuple(row[1:]))
    mismatched = [i for i, spans in enumerate(got)
                  if [(s["start"], s["end"], s["tok_start"], s["tok_end"]) for s in spans] != exp[i]]
    _diag(test="span_cfg1_precise", logit_max_err=lerr, prob_max_err=perr, pairs_with_span_mismatch=len(mismatched))
    assert lerr < LOGIT_TOL and perr < PROB_TOL, (lerr, perr)
    assert mismatched == []


def test_span_forward_bench_shape_precise(ctx):
    """BASELINE configs[2] shape (22 layers, 512-token sequences, two passes) in precise mode: idempotent, and sampled
    sequences within 1e-3 of the fp32 oracle."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    from oracle.modernbert import modernbert_forward_varlen, relevant_prob
    spec = ModernBertSpec(layers=22)
    w = make_modernbert_weights(1001, spec)
    nseq, L = 160, 512
    rng = np.random.default_rng(1003)
    ids2 = rng.integers(5, 50279, size=(nseq, L), dtype=np.int64)
    ids2[:, 0], ids2[:, 30], ids2[:, -1] = spec.cls_id, spec.sep_id, spec.sep_id
    seqs = [ids2[i] for i in range(nseq)]
    ids, cu = _native.Encoder._pack(seqs)
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=65536,
                          precision="precise")
    p1, l1 = enc.span_forward(ids, cu, want_logits=True)
    p2, l2 = enc.span_forward(ids, cu, want_logits=True)
    enc.close()
    assert np.array_equal(l1, l2) and np.array_equal(p1, p2)
    assert np.isfinite(l1).all() and (p1 >= 0).all() and (p1 <= 1).all()
    pick = [0, 127, 128, 159]
    ref = modernbert_forward_varlen(w, [seqs[i] for i in pick], spec, batch=1)
    err = max(float(np.abs(l1[i * L:(i + 1) * L] - r).max()) for i, r in zip(pick, ref))
    perr = max(float(np.abs(p1[i * L:(i + 1) * L] - relevant_prob(r)).max()) for i, r in zip(pick, ref))
    _diag(test="span_forward_bench_shape_precise", nseq=nseq, logit_max_err=err, prob_max_err=perr)
    assert err < LOGIT_TOL and perr < PROB_TOL, (err, perr)


def test_splade_precise_vs_oracle(ctx):
    """BERT-MLM + SPLADE head in precise mode vs the fp32 oracle at 1e-3 (fast mode: 3e-3)."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    from oracle.bert_splade import splade_encode
    spec = BertSpec(layers=2)
    w = make_bert_mlm_weights(1002, spec)
    rng = np.random.default_rng(5)
    seqs = []
    for L in [256, 256, 32, 100, 17, 300]:
        s = rng.integers(1000, spec.vocab_size, size=L)
        s[0], s[-1] = spec.cls_id, spec.sep_id
        seqs.append(s.astype(np.int64))
    enc = _native.Encoder(ctx, _native.ENC_BERT_MLM, w, spec.layers, spec.vocab_size, max_tokens=2048,
                          precision="precise")
    ids, cu = _native.Encoder._pack(seqs)
    out = enc.splade_forward(ids, cu, min_abs=0.0, want_dense=True)
    enc.close()
    ref = splade_encode(w, seqs, spec)
    err = np.abs(out["dense"] - ref)
    _diag(test="splade_precise_vs_oracle", max_err=float(err.max()), ref_max=float(ref.max()))
    assert err.max() < 2e-4
    sup_diff = (out["dense"] != 0) != (ref != 0)
    assert np.all(np.maximum(np.abs(ref), np.abs(out["dense"]))[sup_diff] < 2e-4)
