"""Pin the oracle restatements against the installed third-party model classes
(transformers 5.5.0 ModernBertForTokenClassification / BertForMaskedLM) on seeded weights."""
import numpy as np
import pytest
import torch

from verbatim_rag_b200.synthetic import (BertSpec, ModernBertSpec, make_bert_mlm_weights,
                                         make_modernbert_weights)
from oracle.modernbert import modernbert_forward
from oracle.bert_splade import bert_mlm_hidden, splade_pool


def test_modernbert_matches_transformers():
    from transformers import ModernBertConfig, ModernBertForTokenClassification

    spec = ModernBertSpec(layers=4, vocab_size=2048, cls_id=2041, sep_id=2042, pad_id=2043, unk_id=2040)
    w = make_modernbert_weights(7, spec)
    cfg = ModernBertConfig(vocab_size=spec.vocab_size, num_hidden_layers=spec.layers, num_labels=2,
                           pad_token_id=spec.pad_id, cls_token_id=spec.cls_id, sep_token_id=spec.sep_id,
                           bos_token_id=spec.cls_id, eos_token_id=spec.sep_id, attn_implementation="eager")
    m = ModernBertForTokenClassification(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    rng = np.random.default_rng(0)
    B, L = 3, 200  # > 128 so the sliding window actually masks
    ids = rng.integers(5, 2000, size=(B, L))
    am = np.ones((B, L), dtype=np.int64)
    am[1, 150:] = 0
    am[2, 77:] = 0
    ids[am == 0] = spec.pad_id
    with torch.no_grad():
        ref = m(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(am)).logits.numpy()
    got = modernbert_forward(w, ids, am, spec).numpy()
    valid = am.astype(bool)
    assert np.abs(got[valid] - ref[valid]).max() < 2e-4


def test_bert_mlm_matches_transformers():
    from transformers import BertConfig, BertForMaskedLM

    spec = BertSpec(layers=2, vocab_size=3000)
    w = make_bert_mlm_weights(11, spec)
    cfg = BertConfig(vocab_size=spec.vocab_size, num_hidden_layers=spec.layers, attn_implementation="eager")
    m = BertForMaskedLM(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    sd["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    sd["cls.predictions.decoder.bias"] = sd["cls.predictions.bias"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k for k in missing), missing
    rng = np.random.default_rng(1)
    B, L = 2, 48
    ids = rng.integers(1000, 3000, size=(B, L))
    am = np.ones((B, L), dtype=np.int64)
    am[1, 30:] = 0
    ids[am == 0] = 0
    with torch.no_grad():
        logits = m(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(am)).logits
    ref = (torch.log1p(torch.relu(logits)) * torch.from_numpy(am)[..., None]).max(dim=1).values.numpy()
    hid = bert_mlm_hidden(w, ids, am, spec)
    got = splade_pool(w, hid, am)
    assert np.abs(got - ref).max() < 2e-4


def test_bert_dense_encode_matches_transformers():
    """oracle.bert_splade.dense_encode (encoder -> mean / CLS pooling -> L2 normalise) against transformers' BertModel
    last_hidden_state pooled the way sentence-transformers' Pooling + Normalize modules do."""
    from transformers import BertConfig, BertModel
    from oracle.bert_splade import dense_encode

    spec = BertSpec(layers=2, vocab_size=3000)
    w = make_bert_mlm_weights(12, spec)
    cfg = BertConfig(vocab_size=spec.vocab_size, num_hidden_layers=spec.layers, attn_implementation="eager")
    m = BertModel(cfg, add_pooling_layer=False).eval()
    sd = {k[len("bert."):]: torch.from_numpy(v) for k, v in w.items() if k.startswith("bert.")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k for k in missing), missing
    rng = np.random.default_rng(2)
    seqs = [rng.integers(1000, 3000, size=n) for n in (48, 7, 1, 130)]
    for pooling in ("mean", "cls"):
        got = dense_encode(w, seqs, spec, pooling=pooling, normalize=True)
        for s, g in zip(seqs, got):
            with torch.no_grad():
                h = m(input_ids=torch.from_numpy(np.asarray(s, dtype=np.int64))[None]).last_hidden_state[0]
            v = h.mean(dim=0) if pooling == "mean" else h[0]
            v = torch.nn.functional.normalize(v, dim=0)
            assert np.abs(g - v.numpy()).max() < 2e-5, pooling


def test_cross_encoder_matches_transformers():
    """oracle.heads.cross_encoder_scores against transformers' BertForSequenceClassification (num_labels = 1, pair inputs
    with token types) -- the class behind the reference's CrossEncoder reranker (verbatim_rag/rerankers.py:109-134)."""
    from transformers import BertConfig, BertForSequenceClassification
    from verbatim_rag_b200.synthetic import make_cross_encoder_weights
    from oracle.heads import cross_encoder_scores

    spec = BertSpec(layers=2, vocab_size=3000)
    w = make_cross_encoder_weights(13, spec)
    cfg = BertConfig(vocab_size=spec.vocab_size, num_hidden_layers=spec.layers, num_labels=1, attn_implementation="eager")
    m = BertForSequenceClassification(cfg).eval()
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k for k in missing), missing
    rng = np.random.default_rng(3)
    seqs, types = [], []
    for nq, nd in ((6, 20), (3, 1), (10, 60)):
        seqs.append(np.concatenate([[101], rng.integers(1000, 3000, nq), [102], rng.integers(1000, 3000, nd), [102]]))
        types.append(np.asarray([0] * (nq + 2) + [1] * (nd + 1)))
    got = cross_encoder_scores(w, seqs, types, spec)
    for s, t, g in zip(seqs, types, got):
        with torch.no_grad():
            ref = m(input_ids=torch.from_numpy(s)[None], token_type_ids=torch.from_numpy(t)[None]).logits[0, 0].item()
        assert abs(g - ref) < 2e-5, (g, ref)


def test_qa_sentence_head_matches_transformers(reference_pkgs):
    """oracle.heads.qa_sentence_logits against the reference's own QAModel.forward (packages/core/verbatim_core/
    extractor_models/model.py:59-117) run on a transformers ModernBertModel with the same seeded weights -- QAModel's
    constructor downloads from the hub, so its forward is called on an object assembled by hand."""
    from transformers import ModernBertConfig, ModernBertModel
    from verbatim_rag_b200.synthetic import make_qa_model_weights
    from oracle.heads import qa_sentence_logits
    from verbatim_core.extractor_models.model import QAModel

    spec = ModernBertSpec(layers=3, vocab_size=2048, cls_id=2041, sep_id=2042, pad_id=2043, unk_id=2040)
    w = make_qa_model_weights(5, spec)
    cfg = ModernBertConfig(vocab_size=spec.vocab_size, num_hidden_layers=spec.layers, pad_token_id=spec.pad_id,
                           cls_token_id=spec.cls_id, sep_token_id=spec.sep_id, bos_token_id=spec.cls_id,
                           eos_token_id=spec.sep_id, attn_implementation="eager")
    enc = ModernBertModel(cfg).eval()
    sd = {k[len("model."):]: torch.from_numpy(v) for k, v in w.items() if k.startswith("model.")}
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    qa = QAModel.__new__(QAModel)
    torch.nn.Module.__init__(qa)
    qa.bert = enc
    qa.classifier = torch.nn.Linear(spec.hidden, 2)
    with torch.no_grad():
        qa.classifier.weight.copy_(torch.from_numpy(w["classifier.weight"]))
        qa.classifier.bias.copy_(torch.from_numpy(w["classifier.bias"]))
    rng = np.random.default_rng(4)
    ids = rng.integers(5, 2000, size=150)
    bounds = [(8, 30), (32, 32), (34, 149)]
    with torch.no_grad():
        ref = qa.eval()(torch.from_numpy(ids)[None], torch.ones(1, 150, dtype=torch.long), [bounds])[0].numpy()
    got = qa_sentence_logits(w, ids, bounds, spec)
    assert np.abs(got - ref).max() < 2e-4, (got, ref)
