"""
Oracle: the two further heads of SURVEY.md 8f-4, plain torch fp32 on CPU.  TEST INFRASTRUCTURE.

* ``cross_encoder_scores`` -- what ``SentenceTransformersReranker.rerank`` obtains from ``CrossEncoder.predict``
  (verbatim_rag/rerankers.py:109-134): ``BertForSequenceClassification`` with ``num_labels = 1`` on pair inputs
  ``[CLS] q [SEP] doc [SEP]`` with token types 0 / 1: pooler ``tanh(W_p h_CLS + b_p)`` -> ``W_c . + b_c``
  (transformers modeling_bert.py BertPooler / BertForSequenceClassification; pinned against that class on seeded
  weights in tests/test_oracle_pin.py).  Raw logits: the reference only uses the scores' ORDER.
* ``qa_sentence_logits`` -- ``QAModel.forward`` (packages/core/verbatim_core/extractor_models/model.py:59-117):
  encoder last hidden state, mean over the token rows ``[start, end]`` (inclusive) of each sentence, ``Linear(H, 2)``.
* ``encode_question_and_sentences`` -- the input builder of that model,
  ``QADataset.encode_question_and_sentences_with_offsets`` (extractor_models/dataset.py:108-243):
  ``[CLS] question [SEP] S1 [SEP] S2 ... [SEP]``, sentences that do not fit the budget are dropped.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .bert_splade import bert_encoder_hidden
from .modernbert import modernbert_forward


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


@torch.no_grad()
def cross_encoder_scores(weights: Dict[str, np.ndarray], seqs: Sequence[np.ndarray], type_ids: Sequence[np.ndarray],
                         spec=None) -> np.ndarray:
    out = []
    for ids, tt in zip(seqs, type_ids):
        a = np.asarray(ids, np.int64)[None]
        h = bert_encoder_hidden(weights, a, np.ones_like(a), spec, token_type_ids=np.asarray(tt, np.int64)[None])[0, 0]
        pooled = torch.tanh(_t(weights["bert.pooler.dense.weight"]) @ h + _t(weights["bert.pooler.dense.bias"]))
        out.append(float(_t(weights["classifier.weight"]).reshape(-1) @ pooled + _t(weights["classifier.bias"]).reshape(-1)[0]))
    return np.asarray(out, np.float32)


@torch.no_grad()
def qa_sentence_logits(weights: Dict[str, np.ndarray], ids: np.ndarray, boundaries: Sequence[Tuple[int, int]],
                       spec=None) -> np.ndarray:
    """-> [n_sentences, 2] for one sequence (model.py:85-113: mean of sequence_output[start : end + 1], classifier)."""
    h = modernbert_forward(weights, np.asarray(ids, np.int64)[None], None, spec, return_final=True)[0]
    W, b = _t(weights["classifier.weight"]), _t(weights["classifier.bias"])
    rows = [h[s:e + 1].mean(dim=0) for s, e in boundaries]
    if not rows:
        return np.zeros((0, 2), np.float32)
    return (torch.stack(rows) @ W.t() + b).numpy().astype(np.float32)


def encode_question_and_sentences(tokenizer, question: str, sentences: List[str], max_length: int = 512):
    """dataset.py:108-243 restated for a ``tokenizers``-backed tokenizer object (``.tok``, ``cls_id``, ``sep_id``):
    -> (ids int64, [(start, end)] inclusive token boundaries of the sentences that fit)."""
    budget = max_length - 2                                     # dataset.py:127
    q = list(tokenizer.tok.encode(question, add_special_tokens=False).ids)
    ids = [tokenizer.cls_id] + q[:budget - 2]                   # :130-145: encode with specials, truncate, pop the [SEP]
    bounds: List[Tuple[int, int]] = []
    for sent in sentences:
        s_ids = list(tokenizer.tok.encode(sent, add_special_tokens=False).ids)[:budget]
        if len(ids) + len(s_ids) + 1 > budget:
            break
        ids.append(tokenizer.sep_id)
        start = len(ids)
        ids.extend(s_ids)
        bounds.append((start, len(ids) - 1))
    if len(ids) < budget:
        ids.append(tokenizer.sep_id)
    return np.asarray(ids, np.int64), bounds
