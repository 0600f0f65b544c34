"""
B200QAExtractor -- the legacy ``qa_model`` format of ``ModelSpanExtractor`` (packages/core/verbatim_core/extractors.py:
159-188, 230-283) on the GPU: documents are split into sentences, ``[CLS] question [SEP] S1 [SEP] S2 ... [SEP]`` is
built with the token boundaries of every sentence (``QADataset.encode_question_and_sentences_with_offsets``,
extractor_models/dataset.py:108-243), the encoder runs once and ``QAModel``'s head -- mean of the final hidden states
over each sentence, ``Linear(hidden, 2)`` (extractor_models/model.py:59-117) -- classifies each sentence; sentences with
``softmax(logits)[1] > threshold`` are the spans.  The reference runs one document per forward; here all documents of
all questions of a batch go through one packed varlen pass (``vrag_sentence_forward``).
"""
from __future__ import annotations

import logging
import re
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .interfaces import SpanExtractor
from .models import parse_device, resolve_modernbert

logger = logging.getLogger(__name__)


def split_into_sentences(text: str) -> List[str]:
    """extractors.py:190-195."""
    return [s.strip() for s in re.split(r"(?<=[.!?])\s+", text) if s.strip()]


def encode_question_and_sentences(tokenizer, question: str, sentences: Sequence[str], max_length: int = 512):
    """dataset.py:108-243: -> (token ids, [(start, end)] inclusive boundaries of the sentences that fit the budget)."""
    budget = max_length - 2
    q = tokenizer.tok.encode(question, add_special_tokens=False).ids
    ids = [tokenizer.cls_id] + list(q[: budget - 2])     # encoded with specials, truncated to the budget, [SEP] popped
    bounds: List[Tuple[int, int]] = []
    for enc in tokenizer.tok.encode_batch(list(sentences), add_special_tokens=False):
        s_ids = enc.ids[:budget]
        if len(ids) + len(s_ids) + 1 > budget:           # this and all later sentences are dropped (dataset.py:172-180)
            logger.warning("Legacy QA input exceeded the %d-token budget; dropping %d sentence(s)", max_length,
                           len(sentences) - len(bounds))
            break
        ids.append(tokenizer.sep_id)
        start = len(ids)
        ids.extend(s_ids)
        bounds.append((start, len(ids) - 1))
    if len(ids) < budget:
        ids.append(tokenizer.sep_id)
    return ids, bounds


class B200QAExtractor(SpanExtractor):
    def __init__(self, model_path: str = "synthetic:1001", device: Optional[str] = None, threshold: float = 0.5, *,
                 max_length: int = 512, max_tokens: int = 65536, weights=None, tokenizer=None, num_layers=None,
                 vocab_size=None, precision: str = "fast"):
        self.model_path, self.threshold, self.max_length = model_path, threshold, max_length
        if weights is None:
            weights, tok, layers, vocab = resolve_modernbert(model_path, head="qa_model")
            tokenizer = tokenizer or tok
            num_layers = num_layers or layers
            vocab_size = vocab_size or vocab
        self.tokenizer = tokenizer
        self._ctx = _native.default_context(parse_device(device))
        self._enc = _native.Encoder(self._ctx, _native.ENC_MODERNBERT_SENT, weights, int(num_layers), int(vocab_size),
                                    max_tokens=max_tokens, precision=precision)
        self._lock = threading.Lock()

    def extract_spans(self, question: str, search_results: List[Any]) -> Dict[str, List[str]]:
        return self.extract_spans_batch([question], [search_results])[0]

    def sentence_probs(self, pairs: Sequence[Tuple[str, List[str]]]) -> List[np.ndarray]:
        """P(relevant) of the sentences of each (question, sentences) pair (sentences beyond the budget get none)."""
        ids: List[int] = []
        cu, sip, s0, s1 = [0], [0], [], []
        for q, sents in pairs:
            i, b = encode_question_and_sentences(self.tokenizer, q, sents, self.max_length)
            ids += i
            cu.append(len(ids))
            s0 += [x[0] for x in b]
            s1 += [x[1] for x in b]
            sip.append(len(s0))
        if not s0:
            return [np.zeros(0, np.float32) for _ in pairs]
        with self._lock:
            lg = self._enc.sentence_forward(np.asarray(ids, np.int32), np.asarray(cu, np.int32), np.asarray(sip, np.int32),
                                            np.asarray(s0, np.int32), np.asarray(s1, np.int32))
        m = lg.max(axis=1, keepdims=True)
        e = np.exp(lg - m)
        p = (e[:, 1] / e.sum(axis=1)).astype(np.float32)
        return [p[sip[k]:sip[k + 1]] for k in range(len(pairs))]

    def extract_spans_batch(self, questions: Sequence[str], results_lists: Sequence[List[Any]]) -> List[Dict[str, List[str]]]:
        outs: List[Dict[str, List[str]]] = []
        work: List[Tuple[int, str, List[str]]] = []
        for qi, (q, results) in enumerate(zip(questions, results_lists)):
            rel: Dict[str, List[str]] = {}
            for r in results:
                raw = getattr(r, "text", "")
                rel[raw] = []
                sents = split_into_sentences(raw)
                if sents:
                    work.append((qi, raw, sents))
            outs.append(rel)
        if not work:
            return outs
        try:
            probs = self.sentence_probs([(questions[qi], sents) for qi, _, sents in work])
            for (qi, raw, sents), p in zip(work, probs):
                outs[qi][raw] = [s for s, pi in zip(sents, p.tolist()) if pi > self.threshold]
        except Exception as exc:  # noqa: BLE001 -- reference convention: log, spans stay [] (extractors.py:277-281)
            logger.error("B200 QA-model extraction failed: %s", exc)
        return outs
