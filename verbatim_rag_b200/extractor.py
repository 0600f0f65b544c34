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

    # -- reference interface -------------------------------------------------------------------------
    def extract_spans(self, question: str, search_results: List[Any]) -> Dict[str, List[str]]:
        return self.extract_spans_batch([question], [search_results])[0]

    # -- batched entry point ---------------------------------------------------------------------------
    def extract_spans_batch(self, questions: Sequence[str], results_lists: Sequence[List[Any]]
                            ) -> List[Dict[str, List[str]]]:
        """One GPU pass for many questions; element i is exactly ``extract_spans(questions[i], results_lists[i])``."""
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
        try:
            detailed = self.extract_detailed([(questions[qi], c) for qi, c in pairs])
            for (qi, context), spans in zip(pairs, detailed):
                outs[qi][context] = [s["text"] for s in spans if s["text"].strip()]
        except Exception as exc:  # reference convention: log and return [] (extractors.py:225-227)
            logger.error("B200 highlighter extraction failed: %s", exc)
        return outs

    def extract_detailed(self, pairs: Sequence[Tuple[str, str]]) -> List[List[Dict[str, Any]]]:
        """Spans with offsets and scores for each (question, context) pair:
        ``{"text", "start", "end", "score", "tok_start", "tok_end"}`` (char offsets into the context)."""
        pairs = list(pairs)
        step = self.pipeline_pairs
        if len(pairs) <= step:
            plan = self._tokenize(pairs)
            return self._postprocess(pairs, plan, self._forward(plan))
        # Large batches: host tokenisation of slice i + 1 and span post-processing of slice i - 1 overlap the GPU forward
        # of slice i (one worker thread owns the encoder; the tokenizers library and the ctypes call release the GIL).
        # Every (question, context) pair is independent in every kernel, so slicing does not change any result.
        from concurrent.futures import ThreadPoolExecutor
        out: List[List[Dict[str, Any]]] = []
        with ThreadPoolExecutor(max_workers=1) as gpu:
            pending = None
            for a in range(0, len(pairs), step):
                sl = pairs[a:a + step]
                plan = self._tokenize(sl)
                fut = gpu.submit(self._forward, plan)
                if pending is not None:
                    out.extend(self._postprocess(pending[0], pending[1], pending[2].result()))
                pending = (sl, plan, fut)
            out.extend(self._postprocess(pending[0], pending[1], pending[2].result()))
        return out

    def _forward(self, plan: Dict[str, Any]) -> np.ndarray:
        with self._lock:
            return self._enc.span_forward(plan["ids"], plan["cu"])

    # -- internals -----------------------------------------------------------------------------------------
    def _tokenize(self, pairs: Sequence[Tuple[str, str]]) -> Dict[str, Any]:
        tk = self.tokenizer
        uq = {}
        for q, _ in pairs:
            if q not in uq:
                uq[q] = None
        q_enc = tk.tok.encode_batch(list(uq.keys()), add_special_tokens=False)
        for q, e in zip(uq.keys(), q_enc):
            uq[q] = np.asarray(e.ids, dtype=np.int32)
        # every distinct context is tokenised once (popular chunks are retrieved for many questions of a batch)
        uc: Dict[str, Any] = {}
        for _, c in pairs:
            if c not in uc:
                uc[c] = None
        for c, e in zip(uc.keys(), tk.tok.encode_batch(list(uc.keys()), add_special_tokens=False)):
            uc[c] = (np.asarray(e.ids, dtype=np.int32), np.asarray(e.offsets, dtype=np.int32).reshape(-1, 2))
        cls_a, sep_a = np.asarray([tk.cls_id], np.int32), np.asarray([tk.sep_id], np.int32)
        chunks: List[np.ndarray] = []
        seq_len: List[int] = []
        win_pair: List[int] = []       # window -> pair
        win_range: List[Tuple[int, int]] = []
        win_c0: List[int] = []         # offset of the first context token inside the window's sequence
        ctx_ntok = np.zeros(len(pairs) + 1, dtype=np.int64)
        tok_cs: List[np.ndarray] = []
        tok_ce: List[np.ndarray] = []
        for pi, (q, c) in enumerate(pairs):
            qi = uq[q]
            cids, off = uc[c]
            ctx_ntok[pi + 1] = ctx_ntok[pi] + len(cids)
            tok_cs.append(off[:, 0])
            tok_ce.append(off[:, 1])
            if len(cids) == 0:
                continue
            for s, e in plan_windows(len(qi), len(cids), self.max_length, self.doc_stride):
                chunks.extend((cls_a, qi, sep_a, cids[s:e], sep_a))
                seq_len.append(len(qi) + (e - s) + 3)
                win_pair.append(pi)
                win_range.append((s, e))
                win_c0.append(len(qi) + 2)
        cu = np.zeros(len(seq_len) + 1, dtype=np.int32)
        np.cumsum(seq_len, out=cu[1:])
        ids = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
        return {"ids": ids, "cu": cu, "win_pair": win_pair, "win_range": win_range, "win_c0": win_c0,
                "ctx_indptr": ctx_ntok, "tok_cs": np.concatenate(tok_cs) if tok_cs else np.zeros(0, np.int32),
                "tok_ce": np.concatenate(tok_ce) if tok_ce else np.zeros(0, np.int32)}

    def _context_probs(self, plan: Dict[str, Any], probs: np.ndarray) -> np.ndarray:
        """Per context token: max P(relevant) over the windows containing it."""
        ctx_indptr, cu = plan["ctx_indptr"], plan["cu"]
        p_ctx = np.full(int(ctx_indptr[-1]), -1.0, dtype=np.float32)
        for w, (pi, (s, e), c0) in enumerate(zip(plan["win_pair"], plan["win_range"], plan["win_c0"])):
            a = int(cu[w]) + c0
            dst = p_ctx[ctx_indptr[pi] + s: ctx_indptr[pi] + e]
            np.maximum(dst, probs[a:a + (e - s)], out=dst)
        return p_ctx

    def _postprocess(self, pairs, plan, probs) -> List[List[Dict[str, Any]]]:
        p_ctx = self._context_probs(plan, probs)
        sp = _native.spans_from_probs(p_ctx, plan["tok_cs"], plan["tok_ce"], plan["ctx_indptr"], self.threshold,
                                      self.min_span_chars, self.merge_gap_chars)
        out: List[List[Dict[str, Any]]] = [[] for _ in pairs]
        for c, s, e, sc, ts, te in zip(sp["ctx"].tolist(), sp["start"].tolist(), sp["end"].tolist(),
                                       sp["score"].tolist(), sp["tok_start"].tolist(), sp["tok_end"].tolist()):
            out[c].append({"text": pairs[c][1][s:e], "start": s, "end": e, "score": sc, "tok_start": ts, "tok_end": te})
        return out
