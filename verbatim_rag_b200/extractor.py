"""
B200SpanExtractor -- drop-in for ``ModelSpanExtractor`` (highlighter format),
packages/core/verbatim_core/extractors.py:57-228.

Same constructor knobs, same ``extract_spans(question, search_results) -> Dict[chunk_text, List[span_text]]``
contract, same error convention (a data error yields ``[]`` for the chunk, never an exception; extractors.py:225-227).
Where the reference runs one batch-1 forward per chunk through HF remote code (extractors.py:207-221), this class
tokenises all (question, chunk) pairs, runs ONE packed varlen forward on the GPU through the C ABI
(``vrag_span_forward``) and post-processes all spans in one native call (``vrag_spans_from_probs``).
``extract_spans_batch`` extends the same to many questions at once.
"""
from __future__ import annotations

import logging
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._tokworker import TokenizerWorkers, encode_with
from .interfaces import SpanExtractor
from .models import parse_device, resolve_modernbert

logger = logging.getLogger(__name__)


def plan_windows(n_q: int, n_ctx: int, max_length: int, doc_stride: int) -> List[Tuple[int, int]]:
    """Sliding windows over the context tokens: capacity ``max_length - n_q - 3`` ([CLS] q [SEP] ctx [SEP]),
    consecutive windows overlap by ``doc_stride`` tokens (HF ``truncation='only_second', stride=doc_stride``)."""
    cap = max_length - n_q - 3
    if cap <= 0:
        raise ValueError("question leaves no room for context tokens")
    if n_ctx <= cap:
        return [(0, n_ctx)]
    step = cap - doc_stride
    if step <= 0:
        raise ValueError("doc_stride must be smaller than the context capacity of a window")
    out, s = [], 0
    while True:
        e = min(s + cap, n_ctx)
        out.append((s, e))
        if e >= n_ctx:
            return out
        s += step


class B200SpanExtractor(SpanExtractor):
    DEFAULT_MODEL = "synthetic:1001"

    def __init__(
        self,
        model_path: str = DEFAULT_MODEL,
        device: Optional[str] = None,
        threshold: float = 0.2,
        extraction_mode: str = "individual",
        max_display_spans: int = 5,
        min_span_chars: int = 30,
        merge_gap_chars: int = 20,
        max_length: int = 8192,
        doc_stride: int = 256,
        *,
        max_tokens: int = 65536,
        weights: Optional[Dict[str, np.ndarray]] = None,
        tokenizer=None,
        num_layers: Optional[int] = None,
        vocab_size: Optional[int] = None,
        precision: str = "fast",
        tokenizer_workers: int = 2,
    ):
        """``precision``: ``"fast"`` = fp16 tensor-core operands (logits within ~3.5e-3 of the reference's fp32
        forward, identical spans except at tokens that close to the threshold); ``"precise"`` = split-precision
        operands, logits within 1e-3 (the reference runs the model in fp32, extractors.py:151-157), ~3x the time."""
        self.model_path = model_path
        self.precision = precision
        self.threshold = threshold
        self.extraction_mode = extraction_mode
        self.max_display_spans = max_display_spans
        self.min_span_chars = min_span_chars
        self.merge_gap_chars = merge_gap_chars
        self.max_length = max_length
        self.doc_stride = doc_stride
        self.device = parse_device(device)
        if weights is None:
            weights, tok, layers, vocab = resolve_modernbert(model_path)
            tokenizer = tokenizer or tok
            num_layers = num_layers or layers
            vocab_size = vocab_size or vocab
        self.tokenizer = tokenizer
        self._ctx = _native.default_context(self.device)   # raises NativeError without a CUDA device / library
        self._enc = _native.Encoder(self._ctx, _native.ENC_MODERNBERT_TOKCLS, weights, int(num_layers),
                                    int(vocab_size), max_tokens=max_tokens, precision=precision)
        self._lock = threading.Lock()  # shared across to_thread workers (extractors.py:48-54)
        self.pipeline_pairs = 512      # pairs per slice of the tokenise / forward / post-process pipeline
        # batches of >= worker_min_texts distinct contexts are tokenised by worker processes (_tokworker.py)
        self._workers = TokenizerWorkers(tokenizer.tok, tokenizer_workers)
        self.worker_min_texts = 256
        self.device_spans = True       # span post-processing on the device when every context fits one window

    # -- reference interface -------------------------------------------------------------------------
    def extract_spans(self, question: str, search_results: List[Any]) -> Dict[str, List[str]]:
        return self.extract_spans_batch([question], [search_results])[0]

    # -- batched entry point ---------------------------------------------------------------------------
    def extract_spans_batch(self, questions: Sequence[str], results_lists: Sequence[List[Any]]
                            ) -> List[Dict[str, List[str]]]:
        """One GPU pass for many questions; element i is exactly ``extract_spans(questions[i], results_lists[i])``.

        Error convention of the reference (extractors.py:225-227): a failing chunk yields ``[]`` for THAT chunk and an
        ERROR log line, nothing else is affected.  Pairs that cannot be planned (question longer than the window,
        ``doc_stride`` >= window capacity) are isolated at tokenisation time; if a forward fails, the batch is retried
        one question at a time so that only the offending question's chunks come back empty."""
        pairs: List[Tuple[int, str]] = []          # (question index, context)
        outs: List[Dict[str, List[str]]] = []
        for qi, (q, results) in enumerate(zip(questions, results_lists)):
            rel: Dict[str, List[str]] = {}
            for r in results:
                context = getattr(r, "text", "")
                rel[context] = []
                if context.strip():
                    pairs.append((qi, context))
            outs.append(rel)
        if not pairs:
            return outs

        def run(sub: List[Tuple[int, str]]):
            detailed = self.extract_detailed([(questions[qi], c) for qi, c in sub])
            for (qi, context), spans in zip(sub, detailed):
                outs[qi][context] = [s["text"] for s in spans if s["text"].strip()]

        try:
            run(pairs)
        except Exception as exc:  # noqa: BLE001
            logger.error("B200 highlighter extraction failed for a batch of %d pairs (%s); retrying per question",
                         len(pairs), exc)
            by_q: Dict[int, List[Tuple[int, str]]] = {}
            for qi, c in pairs:
                by_q.setdefault(qi, []).append((qi, c))
            for qi, sub in by_q.items():
                try:
                    run(sub)
                except Exception as exc2:  # noqa: BLE001 -- reference convention: log and return [] for these chunks
                    logger.error("B200 highlighter extraction failed: %s", exc2)
        return outs

    def extract_detailed(self, pairs: Sequence[Tuple[str, str]]) -> List[List[Dict[str, Any]]]:
        """Spans with offsets and scores for each (question, context) pair:
        ``{"text", "start", "end", "score", "tok_start", "tok_end"}`` (char offsets into the context)."""
        pairs = list(pairs)
        step = self.pipeline_pairs
        if len(pairs) <= step:
            plan = self._tokenize(pairs)
            return self._postprocess(pairs, plan, self._forward(plan))
        # Large batches: host tokenisation of slices i + 1, i + 2 and span post-processing of slice i - 1 overlap the GPU
        # forward of slice i (one thread owns the encoder, one runs ahead with the tokeniser: the tokenizers library, the
        # worker processes and the ctypes call all release the GIL).  Every (question, context) pair is independent in
        # every kernel, so slicing does not change any result.
        from concurrent.futures import ThreadPoolExecutor
        out: List[List[Dict[str, Any]]] = []
        slices = [pairs[a:a + step] for a in range(0, len(pairs), step)]
        with ThreadPoolExecutor(max_workers=1) as gpu, ThreadPoolExecutor(max_workers=1) as tk:
            plans = [tk.submit(self._tokenize, sl) for sl in slices[:2]]
            pending = None
            for i, sl in enumerate(slices):
                plan = plans[i].result()
                if i + 2 < len(slices):
                    plans.append(tk.submit(self._tokenize, slices[i + 2]))
                fut = gpu.submit(self._forward, plan)
                if pending is not None:
                    out.extend(self._postprocess(pending[0], pending[1], pending[2].result()))
                pending = (sl, plan, fut)
            out.extend(self._postprocess(pending[0], pending[1], pending[2].result()))
        return out

    def _forward(self, plan: Dict[str, Any]):
        """-> the spans dict of the device post-processing (single-window plans: one sequence per context)
        or the per-token probabilities (documents split into overlapping windows: max over the windows on the host)."""
        if len(plan["cu"]) <= 1:
            return np.zeros(0, np.float32)
        with self._lock:
            if plan["single_window"] and self.device_spans and hasattr(self._enc, "span_extract"):
                wp = plan["win_pair"]
                ip = plan["ctx_indptr"]
                live = np.zeros(len(ip) - 1, bool)
                live[wp] = True
                # offsets of the live contexts only, in sequence order (contexts without a sequence have no tokens or failed)
                sel = np.repeat(live, np.diff(ip))
                sp = self._enc.span_extract(plan["ids"], plan["cu"], plan["win_c0"], plan["win_e"], plan["tok_cs"][sel],
                                            plan["tok_ce"][sel], self.threshold, self.min_span_chars, self.merge_gap_chars)
                sp["ctx"] = wp[sp["ctx"]]           # sequence index -> pair index
                return sp
            return self._enc.span_forward(plan["ids"], plan["cu"])

    # -- internals -----------------------------------------------------------------------------------------
    def _encode_unique(self, texts: List[str]):
        if len(texts) >= self.worker_min_texts:
            return self._workers.encode(texts)
        return encode_with(self.tokenizer.tok, texts)

    def _tokenize(self, pairs: Sequence[Tuple[str, str]]) -> Dict[str, Any]:
        """Token ids of every window ``[CLS] q [SEP] ctx_window [SEP]`` packed back to back, plus the bookkeeping that
        maps window tokens back to context tokens and characters.  Distinct questions / contexts are tokenised once
        (popular chunks are retrieved for many questions of a batch)."""
        tk = self.tokenizer
        P = len(pairs)
        uq: Dict[str, int] = {}
        uc: Dict[str, int] = {}
        q_of = np.empty(P, np.int64)
        c_of = np.empty(P, np.int64)
        for pi, (q, c) in enumerate(pairs):
            q_of[pi] = uq.setdefault(q, len(uq))
            c_of[pi] = uc.setdefault(c, len(uc))
        q_len, q_ids, _ = encode_with(tk.tok, list(uq.keys()))
        c_len, c_ids, c_off = self._encode_unique(list(uc.keys()))
        q_start = np.concatenate([[0], np.cumsum(q_len, dtype=np.int64)])
        c_start = np.concatenate([[0], np.cumsum(c_len, dtype=np.int64)])
        lq = q_len[q_of].astype(np.int64)
        lc = c_len[c_of].astype(np.int64)
        ctx_indptr = np.concatenate([[0], np.cumsum(lc, dtype=np.int64)])

        def rep_gather(starts, lens):
            """indices start_i .. start_i + len_i - 1 for every i, concatenated"""
            tot = int(lens.sum())
            base = np.repeat(starts - (np.cumsum(lens) - lens), lens)
            return base + np.arange(tot, dtype=np.int64)

        g = rep_gather(c_start[c_of], lc)
        tok_cs = c_off[g, 0] if g.size else np.zeros(0, np.int32)
        tok_ce = c_off[g, 1] if g.size else np.zeros(0, np.int32)
        plan: Dict[str, Any] = {"ctx_indptr": ctx_indptr, "tok_cs": np.ascontiguousarray(tok_cs, np.int32),
                                "tok_ce": np.ascontiguousarray(tok_ce, np.int32), "failed": []}
        cap = self.max_length - lq - 3
        bad = (cap <= 0) | ((lc > cap) & (cap - self.doc_stride <= 0))
        for pi in np.nonzero(bad)[0].tolist():   # reference convention: this chunk yields [] (extractors.py:225-227)
            logger.error("B200 highlighter extraction failed: pair %d cannot be windowed (question of %d tokens, "
                         "max_length %d, doc_stride %d)", pi, int(lq[pi]), self.max_length, self.doc_stride)
            plan["failed"].append(pi)
        live = ~bad & (lc > 0)
        if not bool(((lc > cap) & live).any()):
            # every live pair fits one window: fully vectorised assembly
            w_pair = np.nonzero(live)[0]
            wq, wc = lq[w_pair], lc[w_pair]
            seq_len = wq + wc + 3
            cu = np.zeros(len(w_pair) + 1, dtype=np.int64)
            np.cumsum(seq_len, out=cu[1:])
            ids = np.empty(int(cu[-1]), np.int32)
            s0 = cu[:-1]
            ids[s0] = tk.cls_id
            ids[s0 + 1 + wq] = tk.sep_id
            ids[cu[1:] - 1] = tk.sep_id
            ids[rep_gather(s0 + 1, wq)] = q_ids[rep_gather(q_start[q_of[w_pair]], wq)]
            ids[rep_gather(s0 + 2 + wq, wc)] = c_ids[rep_gather(c_start[c_of[w_pair]], wc)]
            plan.update(ids=ids, cu=cu.astype(np.int32), win_pair=w_pair, win_s=np.zeros(len(w_pair), np.int64),
                        win_e=wc, win_c0=wq + 2, single_window=True)
            return plan
        # long documents: overlapping windows (doc_stride), a few pairs at most -- plain loop
        chunks: List[np.ndarray] = []
        seq_len_l: List[int] = []
        win_pair: List[int] = []
        win_s: List[int] = []
        win_e: List[int] = []
        win_c0: List[int] = []
        cls_a, sep_a = np.asarray([tk.cls_id], np.int32), np.asarray([tk.sep_id], np.int32)
        for pi in np.nonzero(live)[0].tolist():
            qi = q_ids[q_start[q_of[pi]]:q_start[q_of[pi]] + lq[pi]]
            cids = c_ids[c_start[c_of[pi]]:c_start[c_of[pi]] + lc[pi]]
            for s, e in plan_windows(len(qi), len(cids), self.max_length, self.doc_stride):
                chunks.extend((cls_a, qi, sep_a, cids[s:e], sep_a))
                seq_len_l.append(len(qi) + (e - s) + 3)
                win_pair.append(pi)
                win_s.append(s)
                win_e.append(e)
                win_c0.append(len(qi) + 2)
        cu = np.zeros(len(seq_len_l) + 1, dtype=np.int32)
        np.cumsum(seq_len_l, out=cu[1:])
        plan.update(ids=np.concatenate(chunks) if chunks else np.zeros(0, np.int32), cu=cu,
                    win_pair=np.asarray(win_pair, np.int64), win_s=np.asarray(win_s, np.int64),
                    win_e=np.asarray(win_e, np.int64), win_c0=np.asarray(win_c0, np.int64), single_window=False)
        return plan

    def _context_probs(self, plan: Dict[str, Any], probs: np.ndarray) -> np.ndarray:
        """Per context token: max P(relevant) over the windows containing it (-1 where no window ran)."""
        ctx_indptr, cu = plan["ctx_indptr"], plan["cu"].astype(np.int64)
        p_ctx = np.full(int(ctx_indptr[-1]), -1.0, dtype=np.float32)
        wp, ws, we, c0 = plan["win_pair"], plan["win_s"], plan["win_e"], plan["win_c0"]
        if len(wp) == 0:
            return p_ctx
        if plan["single_window"]:
            n = we - ws
            tot = int(n.sum())
            off = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(n) - n, n)
            p_ctx[np.repeat(ctx_indptr[wp], n) + off] = probs[np.repeat(cu[:-1] + c0, n) + off]
            return p_ctx
        for w in range(len(wp)):
            a = int(cu[w] + c0[w])
            dst = p_ctx[ctx_indptr[wp[w]] + ws[w]: ctx_indptr[wp[w]] + we[w]]
            np.maximum(dst, probs[a:a + int(we[w] - ws[w])], out=dst)
        return p_ctx

    def _postprocess(self, pairs, plan, probs) -> List[List[Dict[str, Any]]]:
        if isinstance(probs, dict):
            sp = probs
        else:
            p_ctx = self._context_probs(plan, probs)
            sp = _native.spans_from_probs(p_ctx, plan["tok_cs"], plan["tok_ce"], plan["ctx_indptr"], self.threshold,
                                          self.min_span_chars, self.merge_gap_chars)
        out: List[List[Dict[str, Any]]] = [[] for _ in pairs]
        for c, s, e, sc, ts, te in zip(sp["ctx"].tolist(), sp["start"].tolist(), sp["end"].tolist(),
                                       sp["score"].tolist(), sp["tok_start"].tolist(), sp["tok_end"].tolist()):
            out[c].append({"text": pairs[c][1][s:e], "start": s, "end": e, "score": sc, "tok_start": ts, "tok_end": te})
        return out
