"""GPU parity through the PLUGIN classes against the committed golden fixtures (tests/golden/*.npz, made by the oracle).
No oracle forward runs here: the fixtures are the oracle's stored outputs, so these tests are fast on the GPU box."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from conftest import ROOT  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(HERE, "golden")
# FAST mode (fp16 tensor-core operands through 22 layers).  Bars = measured maxima on these fixtures + 20 %
# (logits 3.34e-3, probabilities 1.05e-3).  The north-star 1e-3 is met by the precise mode and
# asserted at 1e-3 in tests/test_gpu_precise.py on the same fixtures.
PROB_TOL = 1.3e-3
LOGIT_TOL = 4.0e-3


def _diag(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_diag.jsonl"), "a") as f:
        f.write(json.dumps(kw, default=float) + "\n")


def test_span_extractor_config1_vs_golden():
    """BASELINE config 1: 4 questions x 8 chunks @128 tokens, full 22-layer model, through B200SpanExtractor."""
    import cases
    from verbatim_rag_b200 import B200SpanExtractor
    c = cases.span_cfg1()
    g = np.load(os.path.join(G, "span_cfg1.npz"))
    assert bytes(g["input_hash"]).decode() == c["hash"]
    ext = B200SpanExtractor(weights=c["weights"], tokenizer=c["tokenizer"], num_layers=c["spec"].layers,
                            vocab_size=c["spec"].vocab_size, max_tokens=8192)
    # raw forward: logits / probabilities vs the oracle's
    plan = ext._tokenize(c["pairs"])
    assert np.array_equal(plan["cu"], g["cu"])
    probs, logits = ext._enc.span_forward(plan["ids"], plan["cu"], want_logits=True)
    lerr = float(np.abs(logits - g["logits"]).max())
    e = np.exp(g["logits"] - g["logits"].max(axis=1, keepdims=True))
    p_ref = (e[:, 1] / e.sum(axis=1)).astype(np.float32)
    perr = float(np.abs(probs - p_ref).max())
    # spans through the public contract
    got = ext.extract_detailed(c["pairs"])
    exp = {i: [] for i in range(len(c["pairs"]))}
    for row in g["spans"].tolist():
        exp[row[0]].append(tuple(row[1:]))
    mismatched_pairs = 0
    for i, spans in enumerate(got):
        mine = [(s["start"], s["end"], s["tok_start"], s["tok_end"]) for s in spans]
        for s in spans:
            assert s["text"] == c["pairs"][i][1][s["start"]:s["end"]]          # verbatim substrings
        mismatched_pairs += mine != exp[i]
    _diag(test="span_cfg1_plugin", logit_max_err=lerr, prob_max_err=perr, pairs=len(got),
          pairs_with_span_mismatch=mismatched_pairs, golden_min_margin=float(g["min_margin_to_threshold"]))
    assert lerr < LOGIT_TOL and perr < PROB_TOL
    # the golden chunks keep every token >= 2.5e-3 from the threshold (tests/golden/cases.py), above the probability
    # tolerance: char and token offsets must be identical on all 32 pairs
    assert float(g["min_margin_to_threshold"]) > PROB_TOL
    assert mismatched_pairs == 0
    # the reference-shaped call: dict keyed by chunk text, duplicates collapse, empty contexts -> []
    class R:
        def __init__(self, t):
            self.text = t
    q0 = c["pairs"][0][0]
    d = ext.extract_spans(q0, [R(ctx) for _, ctx in c["pairs"][:8]] + [R(""), R("  ")])
    assert set(d.keys()) == {ctx for _, ctx in c["pairs"][:8]} | {"", "  "}
    assert d[""] == [] and d["  "] == []
    assert d[c["pairs"][3][1]] == [s["text"] for s in got[3]]


def test_span_extractor_long_documents_windows_vs_oracle():
    """Documents longer than max_length: overlapping windows (doc_stride), max P(relevant) over the windows that hold a
    token, ragged document lengths incl. a 1-token and an exactly-window-sized one -- B200SpanExtractor against the
    reference-shaped CPU plugin (oracle/plugins.py OracleSpanExtractor, restating the v2 process() contract,
    SURVEY.md App. B.2) on a 4-layer model (one global + three local-window layers)."""
    import cases
    from verbatim_rag_b200 import B200SpanExtractor
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    from oracle.plugins import OracleSpanExtractor
    spec = ModernBertSpec(layers=4)
    w = make_modernbert_weights(31, spec)
    tok = cases.tokenizer("modernbert")
    rng = np.random.default_rng(32)
    q = tok.make_question(rng, 9)
    max_length, stride = 160, 48          # 148 context tokens per window
    docs = [tok.make_text(rng, n) for n in (1, 40, 148, 149, 400, 1033)]
    kw = dict(max_length=max_length, doc_stride=stride)
    ext = B200SpanExtractor(weights=w, tokenizer=tok, num_layers=spec.layers, vocab_size=spec.vocab_size,
                            max_tokens=4096, **kw)
    ora = OracleSpanExtractor(w, tok, spec, **kw)
    plan = ext._tokenize([(q, d) for d in docs])
    assert len(plan["win_pair"]) > len(docs) + 5 and int(np.diff(plan["cu"]).max()) <= max_length
    got = ext.extract_detailed([(q, d) for d in docs])
    n_span, n_diff = 0, 0
    for d, spans in zip(docs, got):
        ref = ora.process(q, d)["spans"]
        mine = [(s["start"], s["end"]) for s in spans]
        exp = [(s["start"], s["end"]) for s in ref]
        n_span += len(exp)
        if mine != exp:   # only a token within the probability tolerance of the threshold may differ
            n_diff += 1
            probs = np.asarray(ora.process(q, d).get("token_probs", []), dtype=np.float32)
            assert probs.size == 0 or (np.abs(probs - np.float32(0.2)) < PROB_TOL).any(), (mine, exp)
        assert all(s["text"] == d[s["start"]:s["end"]] for s in spans)
    _diag(test="span_long_documents", docs=len(docs), windows=len(plan["win_pair"]), oracle_spans=n_span,
          docs_with_span_mismatch=n_diff)
    assert n_span > 0 and n_diff <= 1


def test_splade_provider_vs_golden():
    import cases
    from verbatim_rag_b200 import B200SpladeProvider
    s = cases.splade_small()
    g = np.load(os.path.join(G, "splade_small.npz"))
    assert bytes(g["input_hash"]).decode() == s["hash"]
    prov = B200SpladeProvider(weights=s["weights"], tokenizer=s["tokenizer"], num_layers=s["spec"].layers,
                              vocab_size=s["spec"].vocab_size, max_tokens=4096)
    batch = prov.embed_batch(s["texts"])
    worst = 0.0
    for i, d in enumerate(batch):
        a, b = g["indptr"][i], g["indptr"][i + 1]
        ref = dict(zip(g["indices"][a:b].tolist(), g["values"][a:b].tolist()))
        for t in set(ref) | set(d):
            worst = max(worst, abs(ref.get(t, 0.0) - d.get(t, 0.0)))
        assert all(type(k) is int and type(v) is float for k, v in d.items())
    _diag(test="splade_plugin", max_err=worst, nnz=[len(d) for d in batch])
    assert worst < 3.6e-3   # fast mode: measured 2.98e-3 + 20 %
    one = prov.embed_text(s["texts"][0])
    assert set(one) <= set(batch[0]) and all(abs(v) > 1e-6 for v in one.values())
    assert prov.get_dimension() == 30522


def test_vector_store_vs_golden():
    import cases
    from verbatim_rag_b200 import B200VectorStore
    from verbatim_rag_b200.synthetic import csr_to_dicts
    d, t = cases.topk_dense(), cases.topk_sparse()
    gd, gs = np.load(os.path.join(G, "topk_dense.npz")), np.load(os.path.join(G, "topk_sparse.npz"))
    n = d["corpus"].shape[0]
    ids = [f"c{i:07d}" for i in range(n)]
    store = B200VectorStore(dense_dim=768, enable_dense=True, enable_sparse=False)
    store.add_vectors(ids, d["corpus"], None, ids, ids, [{"document_id": f"d{i % 7}"} for i in range(n)])
    for qi in range(4):   # the reference's single-query API
        res = store.query(dense_query=d["queries"][qi].tolist(), top_k=d["k"], search_type="dense")
        assert [r.id for r in res] == [ids[j] for j in gd["ids"][qi]]
        assert np.abs(np.array([r.score for r in res], np.float32) - gd["scores"][qi]).max() <= 1e-6
        assert res[0].metadata["document_id"] == f"d{gd['ids'][qi][0] % 7}"
    batch = store.query_batch_dense(d["queries"], d["k"])
    assert [[r.id for r in rs] for rs in batch] == [[ids[j] for j in row] for row in gd["ids"]]

    ns = len(t["corpus"][0]) - 1
    sids = [f"s{i:05d}" for i in range(ns)]
    sstore = B200VectorStore(enable_dense=False, enable_sparse=True)
    sstore.add_csr(sids, *t["corpus"], sids, sids, [{} for _ in range(ns)])
    res = sstore.query_batch_sparse(t["query_dicts"], t["k"])
    assert [[r.id for r in rs] for rs in res] == [[sids[j] for j in row] for row in gs["ids"]]
    one = sstore.query(sparse_query=t["query_dicts"][0], top_k=t["k"], search_type="sparse")
    assert [r.id for r in one] == [sids[j] for j in gs["ids"][0]]
    assert np.abs(np.array([r.score for r in one], np.float32) - gs["scores"][0]).max() <= 1e-5


def test_sharded_search_two_gpus():
    """2 ranks over NCCL: sharded top-k is bit-identical to the single-GPU top-k (skips on a 1-GPU box)."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(HERE, "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout


def test_filter_pushdown_on_the_scan_kernels(tmp_path):
    """the row mask of ``vrag_index_set_filter`` is honoured by every scan path (FMA scan for one
    query, tensor-core scan for a batch, sparse scan) and equals post-filtered exact search; the same store reopened
    from its on-disk form (8f-2) answers identically."""
    from oracle import flat_topk
    from verbatim_rag_b200 import B200VectorStore
    from verbatim_rag_b200.synthetic import csr_to_dicts, make_dense_corpus, make_dense_queries, make_sparse_rows
    n, k = 30000, 10
    corpus = make_dense_corpus(n, 768, seed=77)
    queries = make_dense_queries(20, 768, seed=78)
    ip, ix, vl = make_sparse_rows(n, seed=79)
    qip, qix, qvl = make_sparse_rows(6, seed=80, query=True)
    ids = [f"c{i:06d}" for i in range(n)]
    metas = [{"year": 1990 + i % 40, "document_id": f"d{i % 11}"} for i in range(n)]
    path = str(tmp_path / "store")
    store = B200VectorStore(db_path=path, dense_dim=768, enable_dense=True, enable_sparse=True)
    for a in range(0, n, 10000):
        b = a + 10000
        store.add_csr(ids[a:b], ip[a:b + 1], ix, vl, ids[a:b], ids[a:b], metas[a:b], dense=corpus[a:b])
    store.delete([ids[3], ids[17]])
    expr = 'metadata["year"] >= 2020 and document_id != "d3"'
    keep = np.array([(m["year"] >= 2020 and m["document_id"] != "d3") for m in metas])
    keep[[3, 17]] = False
    sc = flat_topk.dense_cosine_scores(corpus, queries)
    sc[:, ~keep] = -np.inf
    exp = np.argsort(-sc, axis=1, kind="stable")[:, :k]
    qd = csr_to_dicts(qip, qix, qvl)
    ssc = flat_topk.sparse_ip_scores(ip, ix, vl, 30522, qd)
    ssc[:, ~keep] = -np.inf

    def check(st):
        one = st.query(dense_query=queries[0].tolist(), top_k=k, search_type="dense", filter=expr)      # FMA scan
        assert [r.id for r in one] == [ids[j] for j in exp[0]]
        batch = st.query_batch(dense_queries=queries, top_k=k, search_type="dense", filter=expr)        # tcgen05 scan
        assert [[r.id for r in rs] for rs in batch] == [[ids[j] for j in row] for row in exp]
        assert all(r.metadata["year"] >= 2020 for rs in batch for r in rs)
        sp = st.query_batch(sparse_queries=qd, top_k=k, search_type="sparse", filter=expr)
        for qi, rs in enumerate(sp):
            order = np.lexsort((np.arange(n), -ssc[qi]))[:k]
            order = [j for j in order if ssc[qi][j] > 0]
            assert [r.id for r in rs] == [ids[j] for j in order]
        free = st.query_batch(dense_queries=queries[:5], top_k=k, search_type="dense")                  # mask is gone
        assert any(not keep[int(r.id[1:])] for rs in free for r in rs)

    check(store)
    check(B200VectorStore(db_path=path, dense_dim=768, enable_dense=True, enable_sparse=True))


def test_dense_topk_full_size_properties():
    """BASELINE config 4 at full size (1 M x 768): size-independent properties instead of a CPU oracle --
    (1) a corpus row used as the query returns itself first with cosine 1; (2) scaling a row does not change its
    cosine; (3) the 1000-query batch (tensor-core scan, 63 passes) equals the per-query FMA scan on a sample, ids bit
    for bit; (4) two half-corpus shards merged equal the unsharded search (the multi-GPU merge rule on one GPU)."""
    import torch
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.distributed import merge_host
    ctx = _native.default_context(0)
    n, dim, k = 1_000_000, 768, 10
    g = torch.Generator(device="cuda").manual_seed(1004)
    corpus = torch.randn(n, dim, device="cuda", generator=g)
    corpus[123456] *= 37.5
    full = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
    full.add_dense(corpus)
    probe_rows = [0, 123456, 999_999, 500_000]
    q_self = corpus[probe_rows].cpu().numpy()
    ids, s32 = full.search_dense(q_self, k)
    assert ids[:, 0].tolist() == probe_rows
    assert np.abs(s32[:, 0] - 1.0).max() <= 1e-6
    queries = torch.randn(1000, dim, device="cuda", generator=g).cpu().numpy()
    bid, bs32, bs64 = full.search_dense(queries, k, want64=True)
    assert (np.diff(bs64, axis=1) <= 0).all()                       # sorted by score
    for qi in (0, 499, 999):
        oid, _, os64 = full.search_dense(queries[qi:qi + 1], k, want64=True)
        assert np.array_equal(oid[0], bid[qi]) and np.array_equal(os64[0], bs64[qi])
    half = n // 2
    parts = []
    for lo, hi in ((0, half), (half, n)):
        sh = _native.Index(ctx, _native.INDEX_DENSE_COSINE, dim)
        sh.set_id_base(lo)
        sh.add_dense(corpus[lo:hi].contiguous())
        parts.append(sh.search_dense(queries[:64], k, want64=True))
        sh.close()
    cat_ids = np.concatenate([p[0] for p in parts], axis=1)
    cat_s64 = np.concatenate([p[2] for p in parts], axis=1)
    mid, _, ms64 = merge_host(cat_s64, cat_ids, k)
    assert np.array_equal(mid, bid[:64]) and np.array_equal(ms64, bs64[:64])
    full.close()


def test_end_to_end_batch_pipeline_matches_per_query_plugins():
    """BASELINE config 5 shape on the GPU (SPLADE retrieve top-k + span extraction) through the batched entry points
    of verbatim_rag_b200.pipeline: the batch result equals the plugins called one query at a time."""
    import types
    import cases
    from verbatim_rag_b200 import B200SpanExtractor, B200SpladeProvider, B200VectorStore
    from verbatim_rag_b200.pipeline import index_query_batch
    from verbatim_rag_b200.synthetic import BertSpec, ModernBertSpec, make_bert_mlm_weights, make_modernbert_weights
    mspec, bspec = ModernBertSpec(layers=3), BertSpec(layers=2)
    mtok, btok = cases.tokenizer("modernbert"), cases.tokenizer("bert")
    ext = B200SpanExtractor(weights=make_modernbert_weights(5, mspec), tokenizer=mtok, num_layers=mspec.layers,
                            vocab_size=mspec.vocab_size)
    prov = B200SpladeProvider(weights=make_bert_mlm_weights(6, bspec, decoder_bias_sigmas=3.0), tokenizer=btok,
                              num_layers=bspec.layers, vocab_size=bspec.vocab_size)
    rng = np.random.default_rng(1005)
    chunks = [btok.make_text(rng, 200) for _ in range(300)]
    questions = [btok.make_question(rng, 14) for _ in range(24)]
    store = B200VectorStore(enable_dense=False, enable_sparse=True)
    cid = [f"k{i:04d}" for i in range(len(chunks))]
    store.add_csr(cid, *prov.embed_batch_csr(chunks), chunks, chunks, [{} for _ in chunks])
    index = types.SimpleNamespace(vector_store=store, sparse_provider=prov, dense_provider=None)
    got = index_query_batch(index, questions, k=20)
    for q, rs in zip(questions, got):
        one = store.query(sparse_query=prov.embed_text(q), top_k=20, search_type="sparse")
        assert [r.id for r in rs] == [r.id for r in one]
        assert np.allclose([r.score for r in rs], [r.score for r in one], atol=1e-5)
    spans_b = ext.extract_spans_batch(questions, got)
    for q, rs, sb in zip(questions[:6], got[:6], spans_b[:6]):
        assert ext.extract_spans(q, rs) == sb
        assert all(s in text for text, sp in sb.items() for s in sp)
    # sliced host pipelines (tokenise slice i + 1 while the GPU runs slice i) give bit-identical outputs
    whole = ext.extract_spans_batch(questions, got)
    ext.pipeline_pairs = 64
    assert ext.extract_spans_batch(questions, got) == whole
    ip, ix, vl = prov.embed_batch_csr(chunks)
    prov.pipeline_texts = 77
    ip2, ix2, vl2 = prov.embed_batch_csr(chunks)
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2) and np.array_equal(vl, vl2)
