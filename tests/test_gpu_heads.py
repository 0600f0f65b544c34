"""GPU parity of the two further heads: the cross-encoder reranker (SentenceTransformersReranker,
verbatim_rag/rerankers.py:109-134) and the legacy QAModel sentence classifier (extractors.py:230-283,
extractor_models/model.py:59-117) against the CPU oracle (oracle/heads.py)."""
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


class R:
    def __init__(self, t, i):
        self.text, self.enhanced_text, self.id = t, "", i


@pytest.mark.parametrize("precision,tol", [("fast", 5e-3), ("precise", 1e-4)])
@pytest.mark.parametrize("hidden", [768, 384])
def test_reranker_vs_oracle(precision, tol, hidden):
    """B200Reranker: relevance logits vs the oracle BertForSequenceClassification (BERT-base shape and the MiniLM shape of
    the reference's default cross-encoder/ms-marco-MiniLM-L-6-v2), and the reordering contract: the first rerank_k
    results sorted by score, the tail untouched."""
    import cases
    from verbatim_rag_b200 import B200Reranker
    from verbatim_rag_b200.synthetic import BertSpec, make_cross_encoder_weights
    from oracle.heads import cross_encoder_scores
    spec = (BertSpec(layers=2) if hidden == 768 else BertSpec(hidden=384, intermediate=1536, layers=3, heads=12, head_dim=32))
    w = make_cross_encoder_weights(1004, spec)
    tok = cases.tokenizer("bert")
    rr = B200Reranker(weights=w, tokenizer=tok, num_layers=spec.layers, vocab_size=spec.vocab_size, rerank_k=6,
                      max_length=128, max_tokens=4096, precision=precision)
    rng = np.random.default_rng(9)
    q = tok.make_question(rng, 12)
    results = [R(tok.make_text(rng, int(n)), i) for i, n in enumerate([40, 9, 200, 64, 1, 33, 50, 20])]   # 200 is truncated
    ids, types, cu = rr._pack([(q, r.text) for r in results])
    assert int(np.diff(cu).max()) <= 128 and types[0] == 0 and types[cu[1] - 1] == 1
    got = rr.predict([(q, r.text) for r in results])
    ref = cross_encoder_scores(w, [ids[cu[i]:cu[i + 1]] for i in range(len(results))],
                               [types[cu[i]:cu[i + 1]] for i in range(len(results))], spec)
    err = float(np.abs(got - ref).max())
    _diag(test="reranker_vs_oracle", precision=precision, hidden=hidden, max_err=err, score_std=float(ref.std()))
    assert err < tol, (err, got, ref)
    out = rr.rerank(q, results)
    head = sorted(range(6), key=lambda i: -ref[i])
    assert [r.id for r in out] == head + [6, 7]
    assert rr.rerank(q, []) == []
    two = rr.rerank_batch([q, q], [results, results[:3]])
    assert [r.id for r in two[0]] == [r.id for r in out] and len(two[1]) == 3


@pytest.mark.parametrize("precision,tol", [("fast", 5e-3), ("precise", 1e-4)])
def test_qa_sentence_extractor_vs_oracle(precision, tol):
    """B200QAExtractor: sentence logits vs the oracle QAModel head on the oracle-built input, then the extractor contract
    (sentences above the threshold, keyed by chunk text; over-long documents lose their trailing sentences)."""
    import cases
    from verbatim_rag_b200 import B200QAExtractor
    from verbatim_rag_b200.qa_extractor import encode_question_and_sentences, split_into_sentences
    from verbatim_rag_b200.synthetic import ModernBertSpec, make_qa_model_weights
    from oracle import heads
    spec = ModernBertSpec(layers=3)
    w = make_qa_model_weights(1001, spec)
    tok = cases.tokenizer("modernbert")
    ext = B200QAExtractor(weights=w, tokenizer=tok, num_layers=spec.layers, vocab_size=spec.vocab_size, threshold=0.5,
                          max_length=256, max_tokens=4096, precision=precision)
    rng = np.random.default_rng(10)
    q = tok.make_question(rng, 10)
    docs = [tok.make_text(rng, n) for n in (90, 30, 400, 12)]          # 400 tokens: does not fit 256
    worst, n_sent = 0.0, 0
    for d in docs:
        sents = split_into_sentences(d)
        ids, bounds = encode_question_and_sentences(tok, q, sents, 256)
        oids, obounds = heads.encode_question_and_sentences(tok, q, sents, 256)
        assert list(oids) == list(ids) and obounds == bounds and len(ids) <= 254
        p = ext.sentence_probs([(q, sents)])[0]
        lg = heads.qa_sentence_logits(w, oids, obounds, spec)
        e = np.exp(lg - lg.max(axis=1, keepdims=True))
        pref = e[:, 1] / e.sum(axis=1)
        assert len(p) == len(bounds)
        worst = max(worst, float(np.abs(p - pref).max()))
        n_sent += len(bounds)
    _diag(test="qa_sentences_vs_oracle", precision=precision, max_prob_err=worst, sentences=n_sent)
    assert worst < tol and n_sent > 10
    out = ext.extract_spans(q, [R(d, i) for i, d in enumerate(docs)] + [R("", 9)])
    assert set(out) == set(docs) | {""} and out[""] == []
    for d in docs:
        assert all(s in split_into_sentences(d) for s in out[d])
    assert len(split_into_sentences(docs[2])) > len(ext.sentence_probs([(q, split_into_sentences(docs[2]))])[0])
