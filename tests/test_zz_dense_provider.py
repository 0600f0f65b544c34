"""GPU parity of the dense-embedding path (SentenceTransformersProvider -> B200DenseProvider ->
vrag_dense_forward): BERT encoder stack -> mean / CLS pooling -> L2 normalisation vs the CPU oracle.  Collected last
(file name) so that a regression here never masks the hot-path suites under ``-x``."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# normalised embeddings: entries ~ 1/sqrt(768) = 0.036; fp16 operands give ~1e-3 relative on the hidden states and the
# mean over the tokens averages it down -- measured bar with a 10x margin
EMB_TOL = 5e-4


def _case():
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    spec = BertSpec(layers=2)
    w = make_bert_mlm_weights(1002, spec)
    rng = np.random.default_rng(6)
    seqs = []
    for L in [256, 32, 100, 17, 300, 2]:
        s = rng.integers(1000, spec.vocab_size, size=L)
        s[0], s[-1] = spec.cls_id, spec.sep_id
        seqs.append(s.astype(np.int64))
    return spec, w, seqs


@pytest.mark.parametrize("deferred_ln", [True, False])
@pytest.mark.parametrize("pooling", ["mean", "cls"])
def test_dense_forward_vs_oracle(pooling, deferred_ln, monkeypatch):
    from verbatim_rag_b200 import _native
    from oracle.bert_splade import dense_encode
    spec, w, seqs = _case()
    monkeypatch.setenv("VRAG_BERT_DEFERRED_LN", "1" if deferred_ln else "0")
    ctx = _native.default_context(0)
    enc = _native.Encoder(ctx, _native.ENC_BERT_DENSE, w, spec.layers, spec.vocab_size, max_tokens=2048)
    ids, cu = _native.Encoder._pack(seqs)
    got = enc.dense_forward(ids, cu, _native.POOL_MEAN if pooling == "mean" else _native.POOL_CLS, True)
    raw = enc.dense_forward(ids, cu, _native.POOL_MEAN if pooling == "mean" else _native.POOL_CLS, False)
    enc.close()
    ref = dense_encode(w, seqs, spec, pooling=pooling, normalize=True)
    assert got.shape == ref.shape == (len(seqs), 768) and np.isfinite(got).all()
    assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() < 1e-5
    cos = (got * ref).sum(axis=1)
    assert np.abs(got - ref).max() < EMB_TOL, (float(np.abs(got - ref).max()), cos.tolist())
    assert cos.min() > 0.99999
    # normalize=False returns the pooled vector itself: same direction, norm != 1
    rn = np.linalg.norm(raw, axis=1, keepdims=True)
    assert np.abs(raw / rn - got).max() < 1e-6 and (np.abs(rn - 1.0) > 1e-3).all()


def test_dense_provider_surface():
    """B200DenseProvider: the reference's DenseEmbeddingProvider surface (embed_text / embed_batch -> Python floats,
    get_dimension), batch == one-by-one bit for bit (every text is independent in every kernel)."""
    from verbatim_rag_b200 import B200DenseProvider
    from verbatim_rag_b200.synthetic import BertSpec, SyntheticTokenizer, make_bert_mlm_weights
    spec = BertSpec(layers=2)
    tok = SyntheticTokenizer("bert")
    prov = B200DenseProvider(weights=make_bert_mlm_weights(1002, spec), tokenizer=tok, num_layers=spec.layers,
                             vocab_size=spec.vocab_size)
    rng = np.random.default_rng(3)
    texts = [tok.make_text(rng, n) for n in (12, 40, 700)]      # the last one is truncated to max_seq_length
    batch = prov.embed_batch(texts)
    assert prov.get_dimension() == 768 and len(batch) == 3
    assert all(type(x) is float for v in batch for x in v) and all(len(v) == 768 for v in batch)
    for t, b in zip(texts, batch):
        assert prov.embed_text(t) == b


@pytest.mark.parametrize("precision,tol", [("fast", 5e-4), ("precise", 2e-5)])
@pytest.mark.parametrize("pooling", ["mean", "cls"])
def test_minilm_shape_dense_forward_vs_oracle(pooling, precision, tol):
    """The reference's DEFAULT dense model is all-MiniLM-L6-v2 (embedding_providers.py:55): BERT architecture, 6 layers,
    hidden 384, 12 heads x 32, FFN 1536.  Same kernels: heads zero-padded to 64 dims at load time, N padded to the GEMM
    tile with clipped stores, 384-wide row kernels."""
    from verbatim_rag_b200 import _native
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    from oracle.bert_splade import dense_encode
    spec = BertSpec(hidden=384, intermediate=1536, layers=6, heads=12, head_dim=32)
    w = make_bert_mlm_weights(1003, spec)
    rng = np.random.default_rng(7)
    seqs = []
    for L in [128, 256, 31, 2, 300, 77]:
        s = rng.integers(1000, spec.vocab_size, size=L)
        s[0], s[-1] = spec.cls_id, spec.sep_id
        seqs.append(s.astype(np.int64))
    ctx = _native.default_context(0)
    enc = _native.Encoder(ctx, _native.ENC_BERT_DENSE, w, spec.layers, spec.vocab_size, max_tokens=2048, precision=precision)
    assert enc.hidden == 384
    ids, cu = _native.Encoder._pack(seqs)
    got = enc.dense_forward(ids, cu, _native.POOL_MEAN if pooling == "mean" else _native.POOL_CLS, True)
    enc.close()
    ref = dense_encode(w, seqs, spec, pooling=pooling, normalize=True)
    assert got.shape == ref.shape == (len(seqs), 384) and np.isfinite(got).all()
    err = float(np.abs(got - ref).max())
    assert err < tol, err
    assert (got * ref).sum(axis=1).min() > 0.99999


def test_minilm_shape_provider_and_store():
    """B200DenseProvider(MiniLM shape).get_dimension() == 384 and its vectors drop into B200VectorStore(dense_dim=384),
    the reference's default pairing (milvus_local.py:22 dense_dim=384)."""
    from verbatim_rag_b200 import B200DenseProvider, B200VectorStore
    from verbatim_rag_b200.synthetic import BertSpec, SyntheticTokenizer, make_bert_mlm_weights
    spec = BertSpec(hidden=384, intermediate=1536, layers=2, heads=12, head_dim=32)
    tok = SyntheticTokenizer("bert")
    prov = B200DenseProvider(weights=make_bert_mlm_weights(1003, spec), tokenizer=tok, num_layers=spec.layers,
                             vocab_size=spec.vocab_size)
    assert prov.get_dimension() == 384
    rng = np.random.default_rng(8)
    texts = [tok.make_text(rng, int(n)) for n in rng.integers(10, 120, size=40)]
    vecs = prov.embed_batch(texts)
    assert all(len(v) == 384 for v in vecs)
    store = B200VectorStore(dense_dim=384, enable_dense=True, enable_sparse=False)
    ids = [f"t{i}" for i in range(len(texts))]
    store.add_vectors(ids, vecs, None, texts, texts, [{} for _ in texts])
    for i in (0, 17, 39):
        hit = store.query(dense_query=prov.embed_text(texts[i]), top_k=1, search_type="dense")[0]
        assert hit.id == ids[i] and abs(hit.score - 1.0) < 1e-5
