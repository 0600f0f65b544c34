"""
B200SpladeProvider / B200DenseProvider -- drop-ins for ``SpladeProvider`` and ``SentenceTransformersProvider``
(verbatim_rag/embedding_providers.py:52-80, 117-169).

``embed_text`` / ``embed_batch`` / ``get_dimension`` keep the reference's signatures and output types
(``Dict[int, float]`` with Python int keys / float values; ``List[float]``).  The encoder forward, the SPLADE
``max_L(log1p(relu(logits)))`` pooling and the sparse extraction run on the GPU through the C ABI
(``vrag_splade_forward`` / ``vrag_dense_forward``); the reference's ``[N, 30522]`` dense intermediate and its
Python filtering loops (embedding_providers.py:142-145, 157-163) are replaced by a device-side CSR extraction
with the same filters (``abs(w) > 1e-6`` for embed_text, ``!= 0`` for embed_batch; weights are >= 0).
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .interfaces import DenseEmbeddingProvider, SparseEmbeddingProvider
from .models import parse_device, resolve_bert


class _BertTextEncoder:
    def __init__(self, model_name: str, device, kind: int, max_seq_length: int, max_tokens: int,
                 weights=None, tokenizer=None, num_layers=None, vocab_size=None, precision: str = "fast"):
        self.model_name = model_name
        self.precision = precision
        self.device_index = parse_device(device)
        if weights is None:
            weights, tok, layers, vocab = resolve_bert(model_name)
            tokenizer = tokenizer or tok
            num_layers = num_layers or layers
            vocab_size = vocab_size or vocab
        self.tokenizer = tokenizer
        self.vocab_size = int(vocab_size)
        self.max_seq_length = max_seq_length
        self._ctx = _native.default_context(self.device_index)
        self._enc = _native.Encoder(self._ctx, kind, weights, int(num_layers), int(vocab_size), max_tokens=max_tokens,
                                    precision=precision)
        self._lock = threading.Lock()

    def tokenize(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """[CLS] tokens[:max_seq_length-2] [SEP] per text, packed (ids int32, cu_seqlens int32)."""
        tk = self.tokenizer
        enc = tk.tok.encode_batch(list(texts), add_special_tokens=False)
        body = self.max_seq_length - 2
        cls_a, sep_a = np.asarray([tk.cls_id], np.int32), np.asarray([tk.sep_id], np.int32)
        parts: List[np.ndarray] = []
        lens = []
        for e in enc:
            t = np.asarray(e.ids[:body], dtype=np.int32)
            parts.extend((cls_a, t, sep_a))
            lens.append(len(t) + 2)
        cu = np.zeros(len(lens) + 1, dtype=np.int32)
        np.cumsum(lens, out=cu[1:])
        return (np.concatenate(parts) if parts else np.zeros(0, np.int32)), cu


class B200SpladeProvider(SparseEmbeddingProvider):
    """SPLADE sparse embedding provider on the GPU (reference: embedding_providers.py:117-169)."""

    def __init__(self, model_name: str = "synthetic:1002", device: str = "cuda", *, max_seq_length: int = 512,
                 max_tokens: int = 65536, weights=None, tokenizer=None, num_layers=None, vocab_size=None,
                 precision: str = "fast"):
        """``precision``: "fast" (fp16 tensor-core operands, weights within ~3e-3 of the fp32 reference) or "precise"
        (split-precision operands, within 1e-3; the reference's SparseEncoder runs in fp32)."""
        self.model_name = model_name
        self.device = device
        self._te = _BertTextEncoder(model_name, device, _native.ENC_BERT_MLM, max_seq_length, max_tokens, weights,
                                    tokenizer, num_layers, vocab_size, precision)
        self.pipeline_texts = 1024   # texts per slice of the tokenise / encode pipeline of embed_batch_csr

    def embed_text(self, text: str) -> Dict[int, float]:
        ip, idx, val = self.embed_batch_csr([text], min_abs=1e-6)
        return {int(i): float(v) for i, v in zip(idx.tolist(), val.tolist())}

    def embed_batch(self, texts: List[str]) -> List[Dict[int, float]]:
        ip, idx, val = self.embed_batch_csr(texts, min_abs=0.0)
        idx_l, val_l = idx.tolist(), val.tolist()
        return [dict(zip(idx_l[ip[i]:ip[i + 1]], val_l[ip[i]:ip[i + 1]])) for i in range(len(texts))]

    def embed_batch_csr(self, texts: Sequence[str], min_abs: float = 0.0):
        """CSR (indptr int64, indices int32 ascending, values fp32) -- feeds ``B200VectorStore.add_csr`` without
        building Python dicts."""
        if len(texts) == 0:
            return np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32)
        texts = list(texts)
        step = self.pipeline_texts

        def forward(ids, cu):
            with self._te._lock:
                return self._te._enc.splade_forward(ids, cu, min_abs=min_abs)

        if len(texts) <= step:
            out = forward(*self._te.tokenize(texts))
            return out["indptr"], out["indices"], out["values"]
        # index builds: tokenisation of slice i + 1 overlaps the GPU encode of slice i (texts are independent)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as gpu:
            futs = [gpu.submit(forward, *self._te.tokenize(texts[a:a + step])) for a in range(0, len(texts), step)]
            parts = [f.result() for f in futs]
        indptr = [np.zeros(1, np.int64)]
        base = 0
        for o in parts:
            indptr.append(o["indptr"][1:].astype(np.int64) + base)
            base += int(o["indptr"][-1])
        return (np.concatenate(indptr), np.concatenate([o["indices"] for o in parts]),
                np.concatenate([o["values"] for o in parts]))

    def get_dimension(self) -> int:
        return self._te.vocab_size


class B200DenseProvider(DenseEmbeddingProvider):
    """Dense sentence embeddings (BERT-architecture encoder, mean or CLS pooling, L2 normalised) on the GPU
    (reference: SentenceTransformersProvider, embedding_providers.py:52-80)."""

    def __init__(self, model_name: str = "synthetic:1002", device: str = "cuda", *, pooling: str = "mean",
                 normalize: bool = True, max_seq_length: int = 512, max_tokens: int = 65536, weights=None,
                 tokenizer=None, num_layers=None, vocab_size=None, precision: str = "fast"):
        self.model_name = model_name
        self.device = device
        self.pooling = {"mean": _native.POOL_MEAN, "cls": _native.POOL_CLS}[pooling]
        self.normalize = normalize
        self._te = _BertTextEncoder(model_name, device, _native.ENC_BERT_DENSE, max_seq_length, max_tokens, weights,
                                    tokenizer, num_layers, vocab_size, precision)

    def embed_array(self, texts: Sequence[str]) -> np.ndarray:
        if len(texts) == 0:
            return np.zeros((0, self.get_dimension()), np.float32)
        ids, cu = self._te.tokenize(texts)
        with self._te._lock:
            return self._te._enc.dense_forward(ids, cu, self.pooling, self.normalize)

    def embed_text(self, text: str) -> List[float]:
        return self.embed_array([text])[0].tolist()

    def embed_batch(self, texts: List[str]) -> List[List[float]]:
        return self.embed_array(texts).tolist()

    def get_dimension(self) -> int:
        return int(getattr(self._te._enc, "hidden", 768))
