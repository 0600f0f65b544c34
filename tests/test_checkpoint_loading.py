"""Real-checkpoint load path (reference: ``AutoModel.from_pretrained`` in ModelSpanExtractor._init_highlighter,
packages/core/verbatim_core/extractors.py:151-157; ``SparseEncoder(model_name)`` in SpladeProvider._load_model,
verbatim_rag/embedding_providers.py:125-136): a directory with ``model.safetensors`` (HF parameter names) and
``tokenizer.json`` must resolve to the same weights / token ids as the in-memory objects it was written from, and the
plugins must come up on it.  CPU: the encoder handle is the oracle-backed fake; the GPU twin of this test is
tests/test_gpu_parity_more.py::test_plugins_load_a_checkpoint_directory."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))


def write_checkpoint(path, weights, tokenizer, half=False):
    from safetensors.numpy import save_file
    os.makedirs(path, exist_ok=True)
    w = {k: (v.astype(np.float16) if half else v) for k, v in weights.items()}
    save_file(w, os.path.join(path, "model.safetensors"))
    tokenizer.tok.save(os.path.join(path, "tokenizer.json"))


def test_modernbert_directory_round_trip(tmp_path):
    import cases
    from verbatim_rag_b200.models import resolve_modernbert
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_modernbert_weights
    spec = ModernBertSpec(layers=3)
    w = make_modernbert_weights(9, spec)
    tok = cases.tokenizer("modernbert")
    d = str(tmp_path / "ckpt")
    write_checkpoint(d, w, tok)
    w2, tok2, layers, vocab = resolve_modernbert(d)
    assert layers == 3 and vocab == spec.vocab_size
    assert set(w2) == set(w) and all(w2[k].dtype == np.float32 and np.array_equal(w2[k], w[k]) for k in w)
    rng = np.random.default_rng(0)
    text = tok.make_text(rng, 50)
    a, b = tok.tok.encode(text, add_special_tokens=False), tok2.tok.encode(text, add_special_tokens=False)
    assert a.ids == b.ids and a.offsets == b.offsets
    assert (tok2.cls_id, tok2.sep_id, tok2.pad_id) == (spec.cls_id, spec.sep_id, spec.pad_id)
    # half-precision checkpoints (the usual hub format) are widened to fp32 on load
    write_checkpoint(d, w, tok, half=True)
    w3 = resolve_modernbert(d)[0]
    assert all(v.dtype == np.float32 for v in w3.values())
    assert np.array_equal(w3["classifier.weight"], w["classifier.weight"].astype(np.float16).astype(np.float32))
    with pytest.raises(FileNotFoundError):
        resolve_modernbert(str(tmp_path / "nope"))


def test_bert_directory_round_trip_with_and_without_prefix(tmp_path):
    import cases
    from verbatim_rag_b200.models import resolve_bert
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    spec = BertSpec(layers=2)
    w = make_bert_mlm_weights(4, spec)
    tok = cases.tokenizer("bert")
    d = str(tmp_path / "mlm")
    write_checkpoint(d, w, tok)
    w2, _, layers, vocab = resolve_bert(d)
    assert layers == 2 and vocab == spec.vocab_size and all(np.array_equal(w2[k], w[k]) for k in w)
    # a bare BertModel checkpoint (sentence-transformers dense models) has no 'bert.' prefix and no MLM head
    bare = {k[len("bert."):]: v for k, v in w.items() if k.startswith("bert.")}
    d2 = str(tmp_path / "bare")
    write_checkpoint(d2, bare, tok)
    w3, _, layers3, _ = resolve_bert(d2)
    assert layers3 == 2 and all(np.array_equal(w3["bert." + k], v) for k, v in bare.items())


def test_plugins_come_up_on_a_checkpoint_directory(tmp_path, monkeypatch):
    """B200SpanExtractor(model_path=<dir>) / B200SpladeProvider(model_name=<dir>) == the same plugins built from the
    in-memory weights (host logic; the fake encoder answers with the oracle)."""
    import cases
    import fake_native
    fake_native.install(monkeypatch)
    from verbatim_rag_b200 import B200SpladeProvider
    from verbatim_rag_b200.synthetic import BertSpec, make_bert_mlm_weights
    spec = BertSpec(layers=1)
    w = make_bert_mlm_weights(4, spec)
    tok = cases.tokenizer("bert")
    d = str(tmp_path / "splade")
    write_checkpoint(d, w, tok)
    rng = np.random.default_rng(1)
    texts = [tok.make_text(rng, 20), tok.make_text(rng, 7)]
    from_dir = B200SpladeProvider(d)
    in_mem = B200SpladeProvider(weights=w, tokenizer=tok, num_layers=1, vocab_size=spec.vocab_size)
    assert from_dir.embed_batch(texts) == in_mem.embed_batch(texts)
    assert from_dir.get_dimension() == spec.vocab_size
