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
