"""
B200Reranker -- drop-in for ``SentenceTransformersReranker`` (verbatim_rag/rerankers.py:109-134): a local cross-encoder
(``BertForSequenceClassification``, one relevance logit per (question, text) pair) that reorders the first ``rerank_k``
results.  Same constructor knobs and ``rerank(question, results) -> results`` contract; the forward runs on the GPU
through ``vrag_rerank_forward`` (BERT stack -> pooler -> classifier), all pairs of a call -- or of many calls,
``rerank_batch`` -- in one packed varlen pass instead of ``CrossEncoder.predict``'s padded mini-batches.
"""
from __future__ import annotations

import threading
from typing import Any, List, Sequence, Tuple

import numpy as np

from . import _native
from .interfaces import BaseReranker
from .models import parse_device, resolve_bert


class B200Reranker(BaseReranker):
    def __init__(self, model: str = "synthetic:1004", device: str = "cuda", rerank_k: int = 50, text_field: str = "text",
                 *, max_length: int = 512, max_tokens: int = 65536, weights=None, tokenizer=None, num_layers=None,
                 vocab_size=None, precision: str = "fast"):
        super().__init__(rerank_k=rerank_k, text_field=text_field)
        self.model_name = model
        self.max_length = max_length
        if weights is None:
            weights, tok, layers, vocab = resolve_bert(model, head="cross_encoder")
            tokenizer = tokenizer or tok
            num_layers = num_layers or layers
            vocab_size = vocab_size or vocab
        self.tokenizer = tokenizer
        self._ctx = _native.default_context(parse_device(device))
        self._enc = _native.Encoder(self._ctx, _native.ENC_BERT_CLS, weights, int(num_layers), int(vocab_size),
                                    max_tokens=max_tokens, precision=precision)
        self._lock = threading.Lock()

    def _pack(self, pairs: Sequence[Tuple[str, str]]):
        """``[CLS] q [SEP] text [SEP]`` with token types 0 (first segment incl. its [SEP]) / 1; the text is truncated
        so that the pair fits ``max_length`` (the question is kept whole unless it alone exceeds the budget)."""
        tk = self.tokenizer
        uq, ut = {}, {}
        for q, t in pairs:
            uq.setdefault(q, len(uq))
            ut.setdefault(t, len(ut))
        qe = tk.tok.encode_batch(list(uq), add_special_tokens=False)
        te = tk.tok.encode_batch(list(ut), add_special_tokens=False)
        ids: List[int] = []
        types: List[int] = []
        cu = [0]
        for q, t in pairs:
            qi = qe[uq[q]].ids[: self.max_length - 3]
            ti = te[ut[t]].ids[: max(0, self.max_length - 3 - len(qi))]
            ids += [tk.cls_id] + qi + [tk.sep_id] + ti + [tk.sep_id]
            types += [0] * (len(qi) + 2) + [1] * (len(ti) + 1)
            cu.append(len(ids))
        return np.asarray(ids, np.int32), np.asarray(types, np.int32), np.asarray(cu, np.int32)

    def predict(self, pairs: Sequence[Tuple[str, str]]) -> np.ndarray:
        """Relevance logits of (question, text) pairs (``CrossEncoder.predict`` with the identity activation)."""
        if not pairs:
            return np.zeros(0, np.float32)
        ids, types, cu = self._pack(pairs)
        with self._lock:
            return self._enc.rerank_forward(ids, types, cu)

    def rerank(self, question: str, results: List[Any]) -> List[Any]:
        return self.rerank_batch([question], [results])[0]

    def rerank_batch(self, questions: Sequence[str], results_lists: Sequence[List[Any]]) -> List[List[Any]]:
        """``[self.rerank(q, r) for q, r in zip(...)]`` with one forward for all pairs."""
        heads, tails, pairs = [], [], []
        for q, results in zip(questions, results_lists):
            head, tail = self._split_results(results)
            heads.append(head)
            tails.append(tail)
            pairs += [(q, t) for t in self._get_texts(head)]
        scores = self.predict(pairs).tolist()
        out, o = [], 0
        for results, head, tail in zip(results_lists, heads, tails):
            if not head:
                out.append(results)
                continue
            sc = scores[o:o + len(head)]
            o += len(head)
            # the reference sorts (score, result) tuples descending (rerankers.py:132); equal scores would compare the
            # results themselves there -- here ties keep the retrieval order (stable)
            order = sorted(range(len(head)), key=lambda i: -sc[i])
            out.append([head[i] for i in order] + tail)
        return out
