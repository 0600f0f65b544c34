"""Model resolution for the plugins: where weights and the tokenizer come from.

``model_path`` / ``model_name`` accepted by the plugins:

* ``"synthetic:<seed>[:<layers>]"`` -- seeded synthetic weights + the synthetic word-level tokenizer
  (``synthetic.py``); the only option offline.
* a local directory holding ``model.safetensors`` (HF parameter names) and
  ``tokenizer.json`` -- a real checkpoint such as ``KRLabsOrg/verbatim-rag-modern-bert-v2`` or
  ``naver/splade-v3`` downloaded beforehand.  (The reference resolves hub ids through
  ``AutoModel.from_pretrained`` / ``SparseEncoder``: extractors.py:151-157, embedding_providers.py:125-136;
  hub download is out of scope here.)
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np


class HFTokenizerAdapter:
    """Wraps a ``tokenizers.Tokenizer`` loaded from tokenizer.json behind the small surface the plugins use."""

    def __init__(self, path: str, cls_id: int, sep_id: int, pad_id: int):
        from tokenizers import Tokenizer

        self.tok = Tokenizer.from_file(path)
        self.tok.no_truncation()
        self.tok.no_padding()
        self.cls_id, self.sep_id, self.pad_id = cls_id, sep_id, pad_id


def _load_safetensors(path: str) -> Dict[str, np.ndarray]:
    from safetensors.numpy import load_file

    return {k: np.asarray(v, dtype=np.float32) for k, v in load_file(path).items()}


def resolve_modernbert(model_path: str, head: str = "highlighter"):
    """-> (weights, tokenizer, num_layers, vocab_size).  ``head``: "highlighter" (token classifier, the v2 format) or
    "qa_model" (the legacy sentence classifier: encoder + ``classifier`` only; its checkpoints name the encoder
    ``bert.*`` -- QAModel.bert, extractor_models/model.py:51 -- which is mapped to ``model.*`` here)."""
    from .synthetic import (ModernBertSpec, SyntheticTokenizer, make_modernbert_weights, make_qa_model_weights)

    if model_path.startswith("synthetic"):
        parts = model_path.split(":")
        seed = int(parts[1]) if len(parts) > 1 and parts[1] else 1001
        layers = int(parts[2]) if len(parts) > 2 else 22
        spec = ModernBertSpec(layers=layers)
        make = make_qa_model_weights if head == "qa_model" else make_modernbert_weights
        return make(seed, spec), SyntheticTokenizer("modernbert"), spec.layers, spec.vocab_size
    if os.path.isdir(model_path):
        w = _load_safetensors(os.path.join(model_path, "model.safetensors"))
        if head == "qa_model" and any(k.startswith("bert.") for k in w):
            w = {("model." + k[len("bert."):] if k.startswith("bert.") else k): v for k, v in w.items()}
        layers = 1 + max(int(k.split(".")[2]) for k in w if k.startswith("model.layers."))
        vocab = w["model.embeddings.tok_embeddings.weight"].shape[0]
        spec = ModernBertSpec()
        tok = HFTokenizerAdapter(os.path.join(model_path, "tokenizer.json"), spec.cls_id, spec.sep_id, spec.pad_id)
        return w, tok, layers, vocab
    raise FileNotFoundError(
        f"{model_path!r}: not 'synthetic:<seed>' and not a local checkpoint directory (hub download is not available)")


def resolve_bert(model_name: str, mlm: bool = True, head: str = "mlm"):
    """-> (weights, tokenizer, num_layers, vocab_size).  ``head``: "mlm" (SPLADE / dense providers) or "cross_encoder"
    (``BertForSequenceClassification`` with one label: the reranker)."""
    from .synthetic import BertSpec, SyntheticTokenizer, make_bert_mlm_weights, make_cross_encoder_weights

    if model_name.startswith("synthetic"):
        parts = model_name.split(":")
        seed = int(parts[1]) if len(parts) > 1 and parts[1] else (1004 if head == "cross_encoder" else 1002)
        layers = int(parts[2]) if len(parts) > 2 else 12
        spec = BertSpec(layers=layers)
        make = make_cross_encoder_weights if head == "cross_encoder" else make_bert_mlm_weights
        return make(seed, spec), SyntheticTokenizer("bert"), spec.layers, spec.vocab_size
    if os.path.isdir(model_name):
        w = _load_safetensors(os.path.join(model_name, "model.safetensors"))
        if not any(k.startswith("bert.") for k in w):  # bare BertModel checkpoints lack the 'bert.' prefix
            w = {("bert." + k if not k.startswith("cls.") else k): v for k, v in w.items()}
        layers = 1 + max(int(k.split(".")[3]) for k in w if k.startswith("bert.encoder.layer."))
        vocab = w["bert.embeddings.word_embeddings.weight"].shape[0]
        spec = BertSpec()
        tok = HFTokenizerAdapter(os.path.join(model_name, "tokenizer.json"), spec.cls_id, spec.sep_id, spec.pad_id)
        return w, tok, layers, vocab
    raise FileNotFoundError(
        f"{model_name!r}: not 'synthetic:<seed>' and not a local checkpoint directory (hub download is not available)")


def parse_device(device) -> int:
    """'cuda', 'cuda:3', 3, None -> CUDA device index.  'cpu' is refused: there is no CPU path."""
    if device is None or device == "cuda":
        return int(os.environ.get("LOCAL_RANK", "0")) if device is None else 0
    if isinstance(device, int):
        return device
    s = str(device)
    if s.startswith("cuda:"):
        return int(s.split(":")[1])
    raise ValueError(f"device {device!r}: the B200 plugins run on CUDA devices only (no CPU fallback)")
