"""Further GPU-vs-oracle parity cases (round-2 additions): end-to-end retrieval ids, ``embed_text`` values, a full
8192-token window, outlier channels in the residual stream.  Both arithmetic modes are exercised; the precise mode is
held to the north-star tolerance (1e-3), the fast mode to its measured error + 20 %."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
pytestmark = pytest.mark.gpu


def _diag(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_diag.jsonl"), "a") as f:
        f.write(json.dumps(kw, default=float) + "\n")


@pytest.fixture(scope="module")
def ctx():
    from verbatim_rag_b200 import _native
    return _native.default_context(0)


@pytest.fixture(scope="module")
def retrieval_case():
    """BASELINE configs[1] / [4] shape, bounded: 160 chunks of 96-256 tokens and 24 questions through the full 12-layer
    BERT-MLM SPLADE encoder on the CPU oracle (reference control flow: OracleSpladeProvider -> OracleFlatStore)."""
    import cases
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    from oracle.plugins import OracleFlatStore, OracleSpladeProvider
    spec = BertSpec()
    w = make_bert_mlm_weights(1002, spec)
    tok = cases.tokenizer("bert")
    rng = np.random.default_rng(1005)
    chunks = [tok.make_text(rng, int(n)) for n in rng.integers(96, 257, size=160)]
    questions = [tok.make_question(rng, int(n)) for n in rng.integers(12, 21, size=24)]
    ids = [f"k{i:04d}" for i in range(len(chunks))]
    oprov = OracleSpladeProvider(w, tok, spec)
    ostore = OracleFlatStore(enable_dense=False, enable_sparse=True)
    ostore.add_vectors(ids, None, oprov.embed_batch(chunks), chunks, chunks, [{} for _ in chunks])
    oq = [oprov.embed_text(q) for q in questions]
    k = 20
    ref = [ostore.query(sparse_query=v, top_k=k, search_type="sparse") for v in oq]
    # score gaps of the oracle ranking (k-th vs (k+1)-th and between neighbours): where an approximate encoder may swap
    full = [ostore.query(sparse_query=v, top_k=len(chunks), search_type="sparse") for v in oq]
    return {"spec": spec, "w": w, "tok": tok, "chunks": chunks, "questions": questions, "ids": ids, "k": k,
            "ref": ref, "full": full, "oq": oq}


@pytest.mark.parametrize("precision", ["precise", "fast"])
def test_retrieval_ids_match_oracle_pipeline(retrieval_case, precision):
    """GPU SPLADE encode -> GPU sparse top-k (B200SpladeProvider -> B200VectorStore) against oracle SPLADE -> oracle exact
    scan on the same texts: retrieved chunk ids.  Precise mode: identical ids, in order, for every query, scores within
    1e-3.  Fast mode (term weights within ~3e-3): ids may only differ where the oracle's own scores of the swapped
    candidates are closer than the score tolerance."""
    from verbatim_rag_b200 import B200SpladeProvider, B200VectorStore
    c = retrieval_case
    prov = B200SpladeProvider(weights=c["w"], tokenizer=c["tok"], num_layers=c["spec"].layers,
                              vocab_size=c["spec"].vocab_size, max_tokens=32768, precision=precision)
    store = B200VectorStore(enable_dense=False, enable_sparse=True)
    store.add_csr(c["ids"], *prov.embed_batch_csr(c["chunks"]), c["chunks"], c["chunks"], [{} for _ in c["chunks"]])
    # (b) embed_text values against the oracle's, not only the key set
    worst_w = 0.0
    for q, ov in zip(c["questions"], c["oq"]):
        gv = prov.embed_text(q)
        for t in set(ov) | set(gv):
            worst_w = max(worst_w, abs(ov.get(t, 0.0) - gv.get(t, 0.0)))
        assert all(type(t) is int and type(v) is float and abs(v) > 1e-6 for t, v in gv.items())
    w_tol = 1e-3 if precision == "precise" else 4.1e-3   # fast: measured 3.39e-3 + 20 %
    assert worst_w < w_tol, worst_w
    got = store.query_batch_sparse([prov.embed_text(q) for q in c["questions"]], c["k"])
    exact, explained, worst_s = 0, 0, 0.0
    score_tol = 1e-3 if precision == "precise" else 1e-2   # fast: twice the measured score error (4.9e-3)
    for rs, ref, full in zip(got, c["ref"], c["full"]):
        worst_s = max(worst_s, max(abs(a.score - b.score) for a, b in zip(rs, ref)))
        if [r.id for r in rs] == [r.id for r in ref]:
            exact += 1
            continue
        oracle_score = {r.id: r.score for r in full}
        # every position where the ids differ must involve candidates the ORACLE scores closer than the tolerance
        ok = all(abs(oracle_score.get(a.id, 0.0) - b.score) < score_tol for a, b in zip(rs, ref) if a.id != b.id)
        explained += ok
        assert ok, ([r.id for r in rs], [r.id for r in ref])
    _diag(test="retrieval_ids_vs_oracle_pipeline", precision=precision, queries=len(got), exact=exact,
          explained_by_near_ties=explained, term_weight_max_err=worst_w, score_max_err=worst_s)
    if precision == "precise":
        assert exact == len(got) and worst_s < 1e-3


@pytest.mark.parametrize("precision,tol", [("precise", 1e-3), ("fast", 2.0e-3)])   # fast: measured 1.66e-3 + 20 %
def test_full_8192_token_window(ctx, precision, tol):
    """max_length = 8192 is the reference's default contract (extractors.py:88): one single-window sequence of exactly
    8192 tokens (64 query tiles x 128 key blocks on the global layer, RoPE table used to its last row) plus a short one
    in the same pass, 3 layers (global, local, local), against the fp32 oracle."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    from oracle.modernbert import modernbert_forward_varlen, relevant_prob
    spec = ModernBertSpec(layers=3)
    w = make_modernbert_weights(1001, spec)
    rng = np.random.default_rng(8192)
    seqs = []
    for L in (8192, 300):
        s = rng.integers(5, 50279, size=L).astype(np.int64)
        s[0], s[40], s[-1] = spec.cls_id, spec.sep_id, spec.sep_id
        seqs.append(s)
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=16384,
                          precision=precision)
    ids, cu = _native.Encoder._pack(seqs)
    probs, logits = enc.span_forward(ids, cu, want_logits=True)
    enc.close()
    ref = np.concatenate(modernbert_forward_varlen(w, seqs, spec, batch=1), axis=0)
    err = float(np.abs(logits - ref).max())
    perr = float(np.abs(probs - relevant_prob(ref)).max())
    _diag(test="window_8192", precision=precision, logit_max_err=err, prob_max_err=perr)
    assert np.isfinite(logits).all()
    assert err < tol and perr < tol, (err, perr)


def _outlier_weights(spec, scale=64.0, hot=(3, 97, 200, 411, 555, 767)):
    """Seeded weights re-parameterised so that `hot` channels of the RESIDUAL STREAM carry values `scale` times larger
    while the function computed stays well conditioned, the way trained checkpoints hold their outlier channels: the
    writers of those channels (embedding norm weight, Wo / mlp.Wo rows) are scaled up, their readers (the LayerNorm
    weights; layer 0 has no attn_norm, so its Wqkv columns) are scaled down by the same factor."""
    from verbatim_rag_b200.synthetic import make_modernbert_weights
    hot = list(hot)
    w = {k: v.copy() for k, v in make_modernbert_weights(1001, spec).items()}
    w["model.embeddings.norm.weight"][hot] *= scale
    w["model.layers.0.attn.Wqkv.weight"][:, hot] /= scale
    for i in range(spec.layers):
        p = f"model.layers.{i}."
        w[p + "attn.Wo.weight"][hot, :] *= scale
        w[p + "mlp.Wo.weight"][hot, :] *= scale
        if i > 0:
            w[p + "attn_norm.weight"][hot] /= scale
        w[p + "mlp_norm.weight"][hot] /= scale
    w["model.final_norm.weight"][hot] /= scale
    return w


@pytest.mark.parametrize("precision", ["precise", "fast"])
def test_outlier_channels_in_the_residual_stream(ctx, precision):
    """Trained ModernBERT checkpoints carry a few residual-stream channels far above the rest.  Emulated on the seeded
    weights (`_outlier_weights`): 6 channels of the stream sit ~64x above the others.  The fast path keeps the stream
    as fp16 hi + e5m2 lo planes and folds every LayerNorm into W'' = W gamma - mean(W gamma): this checks range,
    the one-pass row moments and the cancellation there; the precise path must stay within 1e-3."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import ModernBertSpec
    from oracle.modernbert import modernbert_forward, modernbert_forward_varlen, relevant_prob
    spec = ModernBertSpec(layers=6)
    w = _outlier_weights(spec)
    rng = np.random.default_rng(50)
    seqs = []
    for L in (512, 130, 64):
        s = rng.integers(5, 50279, size=L).astype(np.int64)
        s[0], s[-1] = spec.cls_id, spec.sep_id
        seqs.append(s)
    enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=4096,
                          precision=precision)
    ids, cu = _native.Encoder._pack(seqs)
    probs, logits, hidden = enc.debug_span_hidden(ids, cu)
    enc.close()
    ref = np.concatenate(modernbert_forward_varlen(w, seqs, spec, batch=1), axis=0)
    _, hid = modernbert_forward(w, seqs[0][None], None, spec, return_hidden=True)
    stream_max = float(max(h.abs().max() for h in hid))
    stream_med = float(hid[-1].abs().median())
    layer_rel = [float(np.abs(hidden[l, :512] - hid[l][0].numpy()).max() / hid[l][0].abs().max())
                 for l in range(spec.layers + 1)]
    err = float(np.abs(logits - ref).max())
    perr = float(np.abs(probs - relevant_prob(ref)).max())
    _diag(test="outlier_channels", precision=precision, logit_max_err=err, prob_max_err=perr, stream_abs_max=stream_max,
          stream_abs_median=stream_med, layer_rel_err=layer_rel, logit_std=float(ref.std()))
    assert np.isfinite(logits).all() and stream_max > 40.0 * stream_med
    assert err < (1e-3 if precision == "precise" else 1.4e-3), (err, layer_rel)   # fast: measured 1.13e-3 + 20 %


def test_plugins_load_a_checkpoint_directory(tmp_path):
    """Real-checkpoint load path on the GPU (extractors.py:151-157 / embedding_providers.py:125-136): a directory with
    model.safetensors + tokenizer.json gives bit-identical outputs to the plugin built from the in-memory weights."""
    import cases
    from safetensors.numpy import save_file
    from verbatim_rag_b200 import B200SpanExtractor, B200SpladeProvider
    from verbatim_rag_b200.synthetic import BertSpec, ModernBertSpec, make_bert_mlm_weights, make_modernbert_weights
    mspec, bspec = ModernBertSpec(layers=2), BertSpec(layers=2)
    mw, bw = make_modernbert_weights(9, mspec), make_bert_mlm_weights(4, bspec)
    mtok, btok = cases.tokenizer("modernbert"), cases.tokenizer("bert")
    for name, w, tok in (("m", mw, mtok), ("b", bw, btok)):
        os.makedirs(tmp_path / name)
        save_file(w, str(tmp_path / name / "model.safetensors"))
        tok.tok.save(str(tmp_path / name / "tokenizer.json"))
    rng = np.random.default_rng(2)
    q = mtok.make_question(rng, 11)
    docs = [mtok.make_text(rng, n) for n in (120, 40, 300)]
    a = B200SpanExtractor(str(tmp_path / "m"), max_tokens=2048)
    b = B200SpanExtractor(weights=mw, tokenizer=mtok, num_layers=2, vocab_size=mspec.vocab_size, max_tokens=2048)
    assert a.extract_detailed([(q, d) for d in docs]) == b.extract_detailed([(q, d) for d in docs])
    texts = [btok.make_text(rng, n) for n in (64, 9)]
    pa = B200SpladeProvider(str(tmp_path / "b"), max_tokens=2048)
    pb = B200SpladeProvider(weights=bw, tokenizer=btok, num_layers=2, vocab_size=bspec.vocab_size, max_tokens=2048)
    assert pa.embed_batch(texts) == pb.embed_batch(texts) and len(pa.embed_batch(texts)[0]) > 0


def test_device_span_postprocessing_equals_host(ctx):
    """vrag_span_extract (forward + threshold / runs / gap merge / min length on the device, only spans come
    back) must equal vrag_span_forward + vrag_spans_from_probs (the host function, itself checked against the oracle in
    tests/test_spans_host.py) bit for bit -- ragged contexts, several passes, a context without any token, thresholds that
    produce many / no spans."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    spec = ModernBertSpec(layers=2)
    w = make_modernbert_weights(1001, spec)
    rng = np.random.default_rng(33)
    lens = [512, 40, 300, 129, 64, 77, 250, 33, 400, 5]
    nq = [29, 10, 12, 20, 3, 15, 29, 8, 11, 2]
    seqs, first, clen, tcs, tce = [], [], [], [], []
    for L, q in zip(lens, nq):
        s_ = rng.integers(5, 50279, size=L).astype(np.int64)
        s_[0], s_[q + 1], s_[-1] = spec.cls_id, spec.sep_id, spec.sep_id
        seqs.append(s_)
        n_ctx = max(0, L - q - 3)
        first.append(q + 2)
        clen.append(n_ctx)
        starts = np.cumsum(rng.integers(2, 12, size=n_ctx)) if n_ctx else np.zeros(0, np.int64)
        tcs.append(starts.astype(np.int32))
        tce.append((starts + rng.integers(1, 9, size=n_ctx)).astype(np.int32))
    clen[4] = 0                                                     # a context without tokens
    tcs[4], tce[4] = tcs[4][:0], tce[4][:0]
    tcs, tce = np.concatenate(tcs), np.concatenate(tce)
    ids, cu = _native.Encoder._pack(seqs)
    for max_tokens, thr, min_chars, gap in ((4096, 0.2, 30, 20), (640, 0.5, 1, 0), (640, 0.9999, 30, 20), (4096, 0.0, 30, 500)):
        enc = _native.Encoder(ctx, _native.ENC_MODERNBERT_TOKCLS, w, spec.layers, spec.vocab_size, max_tokens=max_tokens)
        probs = enc.span_forward(ids, cu)
        p_ctx = np.concatenate([probs[cu[i] + first[i]: cu[i] + first[i] + clen[i]] for i in range(len(seqs))])
        indptr = np.concatenate([[0], np.cumsum(clen)]).astype(np.int64)
        host = _native.spans_from_probs(p_ctx, tcs, tce, indptr, thr, min_chars, gap)
        dev = enc.span_extract(ids, cu, first, clen, tcs, tce, thr, min_chars, gap)
        enc.close()
        for key in host:
            assert np.array_equal(host[key], dev[key]), (key, max_tokens, thr)
        if thr == 0.2:
            assert len(host["ctx"]) > 5
