"""
Oracle: the v2 ``Highlighter.process`` contract, restated.  TEST INFRASTRUCTURE.

The reference calls ``self.model.process(question=, context=, threshold=, min_span_chars=,
merge_gap_chars=, max_length=, doc_stride=)`` per chunk and keeps ``sp["text"]`` of every
returned span (packages/core/verbatim_core/extractors.py:203-228).  ``process`` itself is
HF remote code (not in /root/reference, not fetchable offline) -- PARITY UNPINNED; the steps
below are the contract of SURVEY.md App. B.2, with the in-repo analogue
``SemanticHighlightExtractor._find_span_regions`` (extractors.py:438-469) as pattern precedent:

1. windows ``[CLS] q [SEP] ctx[s:s+cap] [SEP]`` with ``cap = max_length - len(q) - 3`` and window
   step ``cap - doc_stride`` (== HF ``truncation='only_second', stride=doc_stride``);
2. forward each window -> logits [L, 2] -> ``p = softmax(logits)[:, 1]`` in fp32;
3. per context token: max p over the windows containing it;
4. ``keep = p > float32(threshold)`` on context tokens only;
5. maximal runs of kept tokens -> (char_start of first token, char_end of last token);
6. merge consecutive spans whose gap ``next.start - prev.end <= merge_gap_chars``;
7. drop spans with ``end - start < min_span_chars``;
8. ``text = context[start:end]`` (a verbatim substring: response_builder.py:122 str.find's it),
   ``score`` = mean p of the kept tokens inside the span.
"""

from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np


def plan_windows(n_q: int, n_ctx: int, max_length: int, doc_stride: int) -> List[Tuple[int, int]]:
    """Context-token ranges [(start, end)) of the sliding windows (step 1)."""
    cap = max_length - n_q - 3
    if cap <= 0:
        raise ValueError("question leaves no room for context tokens")
    if n_ctx <= cap:
        return [(0, n_ctx)]
    step = cap - doc_stride
    if step <= 0:
        raise ValueError("doc_stride must be smaller than the context capacity of a window")
    out = []
    s = 0
    while True:
        e = min(s + cap, n_ctx)
        out.append((s, e))
        if e >= n_ctx:
            break
        s += step
    return out


def spans_from_token_probs(context: str, probs: np.ndarray, offsets: Sequence[Tuple[int, int]],
                           threshold: float, min_span_chars: int, merge_gap_chars: int) -> List[Dict]:
    """Steps 4-8 for one context: ``probs[i]`` / ``offsets[i]`` per context token."""
    thr = np.float32(threshold)
    p = np.asarray(probs, dtype=np.float32)
    runs = []  # (char_start, char_end, sum_p, n_tok, tok_start, tok_end)
    i, n = 0, len(p)
    while i < n:
        if p[i] > thr:
            j = i
            acc = 0.0
            while j < n and p[j] > thr:
                acc += float(p[j])
                j += 1
            runs.append([int(offsets[i][0]), int(offsets[j - 1][1]), acc, j - i, i, j])
            i = j
        else:
            i += 1
    merged: List[list] = []
    for r in runs:
        if merged and r[0] - merged[-1][1] <= merge_gap_chars:
            m = merged[-1]
            m[1] = r[1]
            m[2] += r[2]
            m[3] += r[3]
            m[5] = r[5]
        else:
            merged.append(list(r))
    out = []
    for cs, ce, acc, cnt, ts, te in merged:
        if ce - cs < min_span_chars:
            continue
        out.append({"text": context[cs:ce], "start": cs, "end": ce, "score": acc / cnt,
                    "tok_start": ts, "tok_end": te})
    return out


def process(question: str, context: str, *, tokenizer, forward: Callable[[List[np.ndarray]], List[np.ndarray]],
            threshold: float = 0.2, min_span_chars: int = 30, merge_gap_chars: int = 20,
            max_length: int = 8192, doc_stride: int = 256) -> Dict:
    """Full ``process`` contract for one (question, context).

    ``tokenizer`` is a ``tokenizers.Tokenizer``-like object (``encode(text, add_special_tokens=False)``
    giving ``.ids`` and ``.offsets``) with ``cls_id`` / ``sep_id``; ``forward`` maps a list of id
    arrays to a list of logits arrays [L_i, 2].
    """
    from .modernbert import relevant_prob

    q = tokenizer.tok.encode(question, add_special_tokens=False)
    c = tokenizer.tok.encode(context, add_special_tokens=False)
    q_ids = list(q.ids)
    c_ids, c_off = list(c.ids), list(c.offsets)
    if not c_ids:
        return {"spans": []}
    wins = plan_windows(len(q_ids), len(c_ids), max_length, doc_stride)
    seqs = [np.asarray([tokenizer.cls_id] + q_ids + [tokenizer.sep_id] + c_ids[s:e] + [tokenizer.sep_id],
                       dtype=np.int64) for s, e in wins]
    logits = forward(seqs)
    p_ctx = np.full(len(c_ids), -1.0, dtype=np.float32)
    c0 = len(q_ids) + 2
    for (s, e), lg in zip(wins, logits):
        pw = relevant_prob(lg)[c0:c0 + (e - s)]
        p_ctx[s:e] = np.maximum(p_ctx[s:e], pw)
    spans = spans_from_token_probs(context, p_ctx, c_off, threshold, min_span_chars, merge_gap_chars)
    return {"spans": spans, "token_probs": p_ctx}
