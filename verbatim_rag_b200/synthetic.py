"""
Seeded synthetic data for the hot path: model weights, a word-level tokenizer,
texts, and dense / sparse corpora.

There is no tokenizer file, checkpoint or dataset offline,
so every test and benchmark input is generated here from
``numpy.random.Generator(PCG64(seed))`` -- identical on every machine.
The weight dictionaries use the HuggingFace parameter names of the real
checkpoints, so a real ``state_dict`` drops into the same
loader (``weights.pack_modernbert`` / ``weights.pack_bert``).

This module is shared test/bench data plumbing; it contains no arithmetic of the
hot path and never imports ``oracle``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# architecture constants
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ModernBertSpec:
    """ModernBERT-base token classifier (verbatim-rag-modern-bert-v2 shape).

    Constants follow transformers' ModernBertConfig defaults;
    the reference loads this architecture at
    packages/core/verbatim_core/extractors.py:151-157.
    """

    vocab_size: int = 50368
    hidden: int = 768
    intermediate: int = 1152
    layers: int = 22
    heads: int = 12
    head_dim: int = 64
    norm_eps: float = 1e-5
    global_every: int = 3          # layer i is global iff i % 3 == 0
    half_window: int = 64          # |i - j| <= 64 on local layers
    theta_global: float = 160000.0
    theta_local: float = 10000.0
    max_pos: int = 8192
    num_labels: int = 2
    cls_id: int = 50281
    sep_id: int = 50282
    pad_id: int = 50283
    unk_id: int = 50280

    def is_global(self, layer: int) -> bool:
        return layer % self.global_every == 0


@dataclass(frozen=True)
class BertSpec:
    """BERT-base (naver/splade-v3 shape; reference: verbatim_rag/embedding_providers.py:117-136)."""

    vocab_size: int = 30522
    hidden: int = 768
    intermediate: int = 3072
    layers: int = 12
    heads: int = 12
    head_dim: int = 64
    norm_eps: float = 1e-12
    max_pos: int = 512
    type_vocab: int = 2
    cls_id: int = 101
    sep_id: int = 102
    pad_id: int = 0
    unk_id: int = 100


def _normal(rng: np.random.Generator, shape, std: float) -> np.ndarray:
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def _norm_weight(rng: np.random.Generator, n: int) -> np.ndarray:
    return (1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)).astype(np.float32)


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def make_modernbert_weights(seed: int = 1001, spec: ModernBertSpec = ModernBertSpec(),
                            classifier_scale: float = 2.0,
                            classifier_bias: Tuple[float, float] = (0.5, -0.5)) -> Dict[str, np.ndarray]:
    """Seeded ModernBERT token-classifier weights, HF parameter names.

    ``normal(0, 0.02)`` linears, LayerNorm weights 1 + 0.1 N(0,1).  The classifier
    is scaled so P(relevant) crosses the reference's default threshold 0.2
    (extractors.py:83) on a realistic fraction of tokens (random-init would sit at 0.5).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    H, I = spec.hidden, spec.intermediate
    w: Dict[str, np.ndarray] = {}
    w["model.embeddings.tok_embeddings.weight"] = _normal(rng, (spec.vocab_size, H), 0.5)
    w["model.embeddings.norm.weight"] = _norm_weight(rng, H)
    for i in range(spec.layers):
        p = f"model.layers.{i}."
        if i > 0:
            w[p + "attn_norm.weight"] = _norm_weight(rng, H)
        # q/k scaled up so attention is not uniform (softmax has real work to do)
        wqkv = _normal(rng, (3 * H, H), 0.02)
        wqkv[: 2 * H] *= np.float32(2.5)
        w[p + "attn.Wqkv.weight"] = wqkv
        w[p + "attn.Wo.weight"] = _normal(rng, (H, H), 0.02)
        w[p + "mlp_norm.weight"] = _norm_weight(rng, H)
        w[p + "mlp.Wi.weight"] = _normal(rng, (2 * I, H), 0.02)
        w[p + "mlp.Wo.weight"] = _normal(rng, (H, I), 0.02)
    w["model.final_norm.weight"] = _norm_weight(rng, H)
    w["head.dense.weight"] = _normal(rng, (H, H), 0.02)
    w["head.norm.weight"] = _norm_weight(rng, H)
    w["classifier.weight"] = _normal(rng, (spec.num_labels, H), 0.02 * classifier_scale)
    w["classifier.bias"] = np.asarray(classifier_bias, dtype=np.float32)
    return w


def make_bert_mlm_weights(seed: int = 1002, spec: BertSpec = BertSpec(),
                          decoder_bias_sigmas: float = 4.0) -> Dict[str, np.ndarray]:
    """Seeded BERT-MLM weights (HF ``BertForMaskedLM`` names; decoder tied to word embeddings).

    The MLM decoder bias is set to ``-decoder_bias_sigmas * sigma_logit`` so the
    SPLADE vector max_L(log1p(relu(logit))) has a realistic number of non-zeros
    (hundreds, not ~30522; SURVEY.md section 7 'hard parts').
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    H, I = spec.hidden, spec.intermediate
    w: Dict[str, np.ndarray] = {}
    e = "bert.embeddings."
    w[e + "word_embeddings.weight"] = _normal(rng, (spec.vocab_size, H), 0.05)
    w[e + "position_embeddings.weight"] = _normal(rng, (spec.max_pos, H), 0.02)
    w[e + "token_type_embeddings.weight"] = _normal(rng, (spec.type_vocab, H), 0.02)
    w[e + "LayerNorm.weight"] = _norm_weight(rng, H)
    w[e + "LayerNorm.bias"] = _normal(rng, (H,), 0.02)
    for i in range(spec.layers):
        p = f"bert.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            std = 0.05 if nm != "value" else 0.02
            w[p + f"attention.self.{nm}.weight"] = _normal(rng, (H, H), std)
            w[p + f"attention.self.{nm}.bias"] = _normal(rng, (H,), 0.02)
        w[p + "attention.output.dense.weight"] = _normal(rng, (H, H), 0.02)
        w[p + "attention.output.dense.bias"] = _normal(rng, (H,), 0.02)
        w[p + "attention.output.LayerNorm.weight"] = _norm_weight(rng, H)
        w[p + "attention.output.LayerNorm.bias"] = _normal(rng, (H,), 0.02)
        w[p + "intermediate.dense.weight"] = _normal(rng, (I, H), 0.02)
        w[p + "intermediate.dense.bias"] = _normal(rng, (I,), 0.02)
        w[p + "output.dense.weight"] = _normal(rng, (H, I), 0.02)
        w[p + "output.dense.bias"] = _normal(rng, (H,), 0.02)
        w[p + "output.LayerNorm.weight"] = _norm_weight(rng, H)
        w[p + "output.LayerNorm.bias"] = _normal(rng, (H,), 0.02)
    c = "cls.predictions."
    w[c + "transform.dense.weight"] = _normal(rng, (H, H), 0.02)
    w[c + "transform.dense.bias"] = _normal(rng, (H,), 0.02)
    w[c + "transform.LayerNorm.weight"] = _norm_weight(rng, H)
    w[c + "transform.LayerNorm.bias"] = _normal(rng, (H,), 0.02)
    # logit = E[v] . h + b ; h is LayerNorm output (unit variance * gamma) => sigma ~ 0.05*sqrt(H)
    sigma_logit = 0.05 * np.sqrt(H)
    w[c + "bias"] = np.full((spec.vocab_size,), -decoder_bias_sigmas * sigma_logit, dtype=np.float32)
    return w


def make_cross_encoder_weights(seed: int = 1004, spec: BertSpec = BertSpec()) -> Dict[str, np.ndarray]:
    """Seeded ``BertForSequenceClassification`` weights with one label (a cross-encoder reranker such as
    cross-encoder/ms-marco-MiniLM-L-6-v2, verbatim_rag/rerankers.py:112): the BERT stack of ``make_bert_mlm_weights``
    + ``bert.pooler.dense`` + ``classifier`` [1, H]."""
    w = {k: v for k, v in make_bert_mlm_weights(seed, spec).items() if not k.startswith("cls.")}
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    H = spec.hidden
    w["bert.pooler.dense.weight"] = _normal(rng, (H, H), 0.05)
    w["bert.pooler.dense.bias"] = _normal(rng, (H,), 0.02)
    w["classifier.weight"] = _normal(rng, (1, H), 0.2)
    w["classifier.bias"] = np.asarray([0.1], dtype=np.float32)
    return w


def make_qa_model_weights(seed: int = 1001, spec: ModernBertSpec = ModernBertSpec()) -> Dict[str, np.ndarray]:
    """Seeded weights of the legacy ``QAModel`` (packages/core/verbatim_core/extractor_models/model.py:13-57): a
    ModernBERT encoder + ``classifier`` Linear(H, 2) on mean-pooled sentence states -- no prediction head."""
    w = {k: v for k, v in make_modernbert_weights(seed, spec, classifier_scale=10.0,
                                                  classifier_bias=(0.2, -0.2)).items() if not k.startswith("head.")}
    return w


# --------------------------------------------------------------------------------------
# tokenizer + text
# --------------------------------------------------------------------------------------
_SYL = [c + v for c in "bdfgklmnprstvz" for v in "aeiou"]  # 70 syllables


def _word_for(n: int) -> str:
    """Unique pronounceable pseudo-word for integer n >= 0 (base-70 syllables, >= 2 syllables)."""
    out = []
    n0 = n
    while True:
        out.append(_SYL[n % 70])
        n //= 70
        if n == 0:
            break
    if len(out) == 1:
        out.append(_SYL[(n0 * 7 + 3) % 70] + "x")  # keep 1-digit words distinct from 2-digit ones
    return "".join(reversed(out))


class SyntheticTokenizer:
    """Word-level tokenizer over a generated vocabulary (``tokenizers`` library, offsets on).

    Stands in for the HF tokenizer that the reference's remote model owns
    (extractors.py:151-157) / that SparseEncoder loads (embedding_providers.py:125-136).
    Every word is exactly one token, so token counts of synthetic texts are exact.
    """

    def __init__(self, kind: str = "modernbert"):
        from tokenizers import Tokenizer, models, pre_tokenizers, processors

        if kind == "modernbert":
            spec = ModernBertSpec()
            specials = {"[UNK]": spec.unk_id, "[CLS]": spec.cls_id, "[SEP]": spec.sep_id, "[PAD]": spec.pad_id}
            first_word, last_word = 5, 50279
        elif kind == "bert":
            spec = BertSpec()
            specials = {"[PAD]": spec.pad_id, "[UNK]": spec.unk_id, "[CLS]": spec.cls_id, "[SEP]": spec.sep_id}
            first_word, last_word = 1000, spec.vocab_size - 1
        else:
            raise ValueError(f"unknown tokenizer kind {kind!r}")
        self.kind = kind
        self.spec = spec
        self.period_id = first_word - 1
        vocab = dict(specials)
        vocab["."] = self.period_id
        self.first_word, self.last_word = first_word, last_word
        self._words: List[Optional[str]] = [None] * spec.vocab_size
        for i in range(first_word, last_word + 1):
            wd = _word_for(i - first_word)
            vocab[wd] = i
            self._words[i] = wd
        self._words[self.period_id] = "."
        used = set(vocab.values())
        for i in range(spec.vocab_size):  # fill the unused ids so the vocab is dense
            if i not in used:
                vocab[f"[unused{i}]"] = i
        tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="[UNK]"))
        tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.WhitespaceSplit(), pre_tokenizers.Punctuation()])
        tok.post_processor = processors.TemplateProcessing(
            single="[CLS] $A [SEP]",
            pair="[CLS] $A [SEP] $B:1 [SEP]:1",
            special_tokens=[("[CLS]", spec.cls_id), ("[SEP]", spec.sep_id)],
        )
        self.tok = tok
        self.cls_id, self.sep_id, self.pad_id = spec.cls_id, spec.sep_id, spec.pad_id

    def word(self, token_id: int) -> str:
        return self._words[token_id]

    # -- text synthesis ------------------------------------------------------------------
    def make_text(self, rng: np.random.Generator, n_tokens: int, sentence_len=(8, 24)) -> str:
        """Text of exactly ``n_tokens`` tokens: sentences of 8-24 words each ending in '.'."""
        toks: List[str] = []
        remaining = n_tokens
        while remaining > 0:
            if remaining == 1:
                toks.append(self.word(int(rng.integers(self.first_word, self.last_word + 1))))
                break
            n = int(min(rng.integers(sentence_len[0], sentence_len[1] + 1), remaining - 1))
            ids = rng.integers(self.first_word, self.last_word + 1, size=n)
            sent = " ".join(self.word(int(i)) for i in ids)
            toks.append(sent + ".")
            remaining -= n + 1
        return " ".join(toks)

    def make_question(self, rng: np.random.Generator, n_words: int) -> str:
        ids = rng.integers(self.first_word, self.last_word + 1, size=n_words)
        return " ".join(self.word(int(i)) for i in ids)

    # -- encoding ------------------------------------------------------------------------
    def encode_pairs(self, pairs: Sequence[Tuple[str, str]]):
        """Batch-encode (question, context) pairs -> list of tokenizers.Encoding."""
        return self.tok.encode_batch(list(pairs))

    def encode_texts(self, texts: Sequence[str], max_length: Optional[int] = None):
        if max_length is not None:
            self.tok.enable_truncation(max_length=max_length)
        else:
            self.tok.no_truncation()
        out = self.tok.encode_batch(list(texts))
        self.tok.no_truncation()
        return out


# --------------------------------------------------------------------------------------
# corpora for the similarity scan
# --------------------------------------------------------------------------------------
def make_dense_corpus(n: int, dim: int = 768, seed=1004, shard: int = 0) -> np.ndarray:
    """N(0,1) fp32 rows, NOT pre-normalised; per-shard stream ``default_rng([seed, shard])``."""
    rng = np.random.default_rng([seed, shard])
    return rng.standard_normal((n, dim), dtype=np.float32)


def make_dense_queries(q: int, dim: int = 768, seed: int = 2004) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((q, dim), dtype=np.float32)


def make_sparse_rows(n: int, seed: int = 1002, vocab: int = 30522, mean_nnz: float = 160.0,
                     nnz_range: Tuple[int, int] = (16, 512), query: bool = False
                     ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Directly synthesised sparse vectors as CSR (indptr int64, indices int32 ascending, values fp32).

    doc nnz ~ clip(lognormal(ln mean, 0.35), 16, 512); query nnz ~ U{24..48};
    indices Zipf-like over the vocabulary without replacement; values |N(0,1)| + 0.05.
    """
    rng = np.random.default_rng(seed)
    if query:
        nnz = rng.integers(24, 49, size=n)
    else:
        nnz = np.clip(np.round(rng.lognormal(np.log(mean_nnz), 0.35, size=n)), nnz_range[0], nnz_range[1]).astype(np.int64)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nnz, out=indptr[1:])
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    p = ranks ** -1.1
    p /= p.sum()
    perm = np.random.default_rng(seed + 7).permutation(vocab)  # zipf rank -> token id
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    cdf = np.cumsum(p)
    for i in range(n):
        k = int(nnz[i])
        # sample without replacement: draw extra, unique, top up if short
        got: np.ndarray = np.unique(np.searchsorted(cdf, rng.random(2 * k + 16)))
        while got.size < k:
            got = np.unique(np.concatenate([got, np.searchsorted(cdf, rng.random(2 * k + 16))]))
        got = rng.permutation(got)[:k]
        ids = np.sort(perm[np.minimum(got, vocab - 1)]).astype(np.int32)
        indices[indptr[i]:indptr[i + 1]] = ids
    values = (np.abs(rng.standard_normal(indices.size, dtype=np.float32)) + np.float32(0.05)).astype(np.float32)
    return indptr, indices, values


def make_sparse_rows_device(n: int, seed: int = 1002, vocab: int = 30522, mean_nnz: float = 160.0,
                            nnz_range: Tuple[int, int] = (16, 512), device="cuda", chunk: int = 100_000):
    """Same distribution family as ``make_sparse_rows`` for corpora too large for its per-row Python loop (the 1 M-document
    variant of BASELINE configs[1]): generated with torch on ``device`` in chunks, returned as host CSR.  Per row
    ~1.25 x nnz Zipf(1.1) draws are de-duplicated, so the realised nnz is a little below the lognormal target."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    ranks = torch.arange(1, vocab + 1, dtype=torch.float64, device=device)
    cdf = torch.cumsum(ranks ** -1.1 / (ranks ** -1.1).sum(), 0).float()
    perm = torch.from_numpy(np.random.default_rng(seed + 7).permutation(vocab)).to(device)
    K = int(1.25 * nnz_range[1]) + 8
    counts, idx_parts = [], []
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        tgt = torch.exp(torch.randn(m, device=device, generator=g) * 0.35 + float(np.log(mean_nnz)))
        tgt = tgt.round().clamp(nnz_range[0], nnz_range[1])
        cand = (tgt * 1.25).long() + 8
        u = torch.rand(m, K, device=device, generator=g)
        tok = perm[torch.searchsorted(cdf, u).clamp_max(vocab - 1)]
        tok = torch.where(torch.arange(K, device=device)[None, :] < cand[:, None], tok, torch.full_like(tok, vocab))
        tok, _ = torch.sort(tok, dim=1)
        keep = tok < vocab
        keep[:, 1:] &= tok[:, 1:] != tok[:, :-1]
        counts.append(keep.sum(1).cpu())
        idx_parts.append(tok[keep].to(torch.int32).cpu())
    nnz = torch.cat(counts).numpy().astype(np.int64)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nnz, out=indptr[1:])
    indices = torch.cat(idx_parts).numpy()
    values = (np.abs(np.random.default_rng(seed + 11).standard_normal(indices.size, dtype=np.float32))
              + np.float32(0.05)).astype(np.float32)
    return indptr, indices, values


def csr_to_dicts(indptr: np.ndarray, indices: np.ndarray, values: np.ndarray) -> List[Dict[int, float]]:
    """CSR -> the reference's ``List[Dict[int, float]]`` sparse format (embedding_providers.py:161-163)."""
    out = []
    for i in range(len(indptr) - 1):
        a, b = int(indptr[i]), int(indptr[i + 1])
        out.append({int(k): float(v) for k, v in zip(indices[a:b], values[a:b])})
    return out
