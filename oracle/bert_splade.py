"""
Oracle: BERT-MLM forward + SPLADE pooling + the reference's dict conversion.  TEST INFRASTRUCTURE.

Restates ``SpladeProvider`` (verbatim_rag/embedding_providers.py:117-169):

* ``SparseEncoder.encode`` [third-party, sentence-transformers 5.x, not installed]:
  MLMTransformer (``BertForMaskedLM``) -> ``SpladePooling(max, relu)``:
  ``emb = max over tokens of log1p(relu(logits)) * attention_mask``  (SURVEY.md App. B.1);
* embed_text's ``abs(w) > 1e-6`` filter (embedding_providers.py:142-145);
* embed_batch's ``np.nonzero`` filter (embedding_providers.py:157-163).

The BERT forward follows transformers/models/bert/modeling_bert.py (5.5.0; embeddings,
self-attention, post-LN blocks, MLM head :471-511); tests/test_oracle_pin.py checks it
against ``BertForMaskedLM`` on seeded weights.
"""

from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F


def _t(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


@torch.no_grad()
def bert_mlm_hidden(weights: Dict[str, np.ndarray], input_ids, attention_mask, spec=None,
                    emulate_fp16: bool = False) -> torch.Tensor:
    """MLM-head transform output [B, L, H] (input of the tied decoder)."""
    return _bert_forward(weights, input_ids, attention_mask, spec, emulate_fp16, mlm_transform=True)


@torch.no_grad()
def bert_encoder_hidden(weights: Dict[str, np.ndarray], input_ids, attention_mask, spec=None,
                        emulate_fp16: bool = False, token_type_ids=None) -> torch.Tensor:
    """Final hidden states of the encoder stack [B, L, H] (transformers BertModel.last_hidden_state,
    modeling_bert.py BertEncoder) -- the input of the sentence pooling of a dense embedding model."""
    return _bert_forward(weights, input_ids, attention_mask, spec, emulate_fp16, mlm_transform=False,
                         token_type_ids=token_type_ids)


@torch.no_grad()
def dense_encode(weights, seqs: List[np.ndarray], spec=None, pooling: str = "mean", normalize: bool = True) -> np.ndarray:
    """Sentence embeddings the way sentence-transformers composes them for BERT-architecture models (the reference's
    SentenceTransformersProvider, embedding_providers.py:52-80): Transformer -> Pooling (mean over the attended tokens,
    or the [CLS] token) -> Normalize (L2).  One sequence at a time, so no padding enters the mean.  -> [N, H] fp32."""
    out = []
    for s in seqs:
        ids = np.asarray(s, dtype=np.int64)[None]
        h = bert_encoder_hidden(weights, ids, np.ones_like(ids), spec)[0]
        v = h.mean(dim=0) if pooling == "mean" else h[0]
        if normalize:
            v = v / v.norm().clamp_min(1e-12)
        out.append(v.numpy())
    return np.stack(out).astype(np.float32)


def _bert_forward(weights, input_ids, attention_mask, spec, emulate_fp16, mlm_transform, token_type_ids=None) -> torch.Tensor:
    from verbatim_rag_b200.synthetic import BertSpec

    spec = spec or BertSpec()
    ids = _t(input_ids).long()
    am = _t(attention_mask).bool()
    B, L = ids.shape
    H, nh, dh = spec.hidden, spec.heads, spec.head_dim

    def rq(x):
        return x.half().float() if emulate_fp16 else x

    def lin(x, w, b):
        return rq(x) @ rq(_t(w)).t() + _t(b)

    def ln(x, pfx):
        return F.layer_norm(x, (H,), _t(weights[pfx + ".weight"]), _t(weights[pfx + ".bias"]), spec.norm_eps)

    e = "bert.embeddings."
    x = (_t(weights[e + "word_embeddings.weight"])[ids]
         + _t(weights[e + "position_embeddings.weight"])[:L][None]
         + (_t(weights[e + "token_type_embeddings.weight"])[0][None, None] if token_type_ids is None
            else _t(weights[e + "token_type_embeddings.weight"])[_t(token_type_ids).long()]))
    x = ln(x, e + "LayerNorm")
    neg = torch.finfo(torch.float32).min
    mask = torch.zeros(B, 1, 1, L).masked_fill(~am[:, None, None, :], neg)
    for i in range(spec.layers):
        p = f"bert.encoder.layer.{i}."
        q = lin(x, weights[p + "attention.self.query.weight"], weights[p + "attention.self.query.bias"])
        k = lin(x, weights[p + "attention.self.key.weight"], weights[p + "attention.self.key.bias"])
        v = lin(x, weights[p + "attention.self.value.weight"], weights[p + "attention.self.value.bias"])
        q, k, v = (t.view(B, L, nh, dh).transpose(1, 2) for t in (q, k, v))
        s = (rq(q) @ rq(k).transpose(2, 3)) * (dh ** -0.5) + mask
        pr = torch.softmax(s, dim=-1, dtype=torch.float32)
        o = (rq(pr) @ rq(v)).transpose(1, 2).reshape(B, L, H)
        x = ln(x + lin(o, weights[p + "attention.output.dense.weight"], weights[p + "attention.output.dense.bias"]),
               p + "attention.output.LayerNorm")
        f = F.gelu(lin(x, weights[p + "intermediate.dense.weight"], weights[p + "intermediate.dense.bias"]))
        x = ln(x + lin(f, weights[p + "output.dense.weight"], weights[p + "output.dense.bias"]),
               p + "output.LayerNorm")
    if not mlm_transform:
        return x
    c = "cls.predictions."
    t = F.gelu(lin(x, weights[c + "transform.dense.weight"], weights[c + "transform.dense.bias"]))
    return ln(t, c + "transform.LayerNorm")


@torch.no_grad()
def splade_pool(weights, hidden: torch.Tensor, attention_mask, emulate_fp16: bool = False) -> np.ndarray:
    """[B, V] = max_L( log1p(relu(hidden @ E^T + bias)) * mask )  (SpladePooling 'max'/'relu')."""
    am = _t(attention_mask).float()
    E = _t(weights["bert.embeddings.word_embeddings.weight"])
    b = _t(weights["cls.predictions.bias"])
    out = []
    for r in range(hidden.shape[0]):  # one row at a time: [L, V] fp32 is 31 MB at L=256
        h = hidden[r]
        if emulate_fp16:
            logits = h.half().float() @ E.half().float().t() + b
        else:
            logits = h @ E.t() + b
        act = torch.log1p(torch.relu(logits)) * am[r][:, None]
        out.append(act.max(dim=0).values)
    return torch.stack(out).numpy()


@torch.no_grad()
def splade_encode(weights, seqs: List[np.ndarray], spec=None, batch_size: int = 32, **kw) -> np.ndarray:
    """Dense [N, V] SPLADE vectors for unpadded id arrays.

    Mirrors SparseEncoder.encode as the reference calls it (embedding_providers.py:151):
    sort by length (descending), batches of 32, pad, no_grad; results in the original order.
    """
    from verbatim_rag_b200.synthetic import BertSpec

    spec = spec or BertSpec()
    N = len(seqs)
    out = np.zeros((N, spec.vocab_size), dtype=np.float32)
    order = sorted(range(N), key=lambda i: -len(seqs[i]))
    for s in range(0, N, batch_size):
        grp = order[s:s + batch_size]
        L = max(len(seqs[i]) for i in grp)
        ids = np.full((len(grp), L), spec.pad_id, dtype=np.int64)
        am = np.zeros((len(grp), L), dtype=np.int64)
        for r, i in enumerate(grp):
            ids[r, :len(seqs[i])] = seqs[i]
            am[r, :len(seqs[i])] = 1
        hid = bert_mlm_hidden(weights, ids, am, spec, **kw)
        vec = splade_pool(weights, hid, am, **kw)
        for r, i in enumerate(grp):
            out[i] = vec[r]
    return out


def to_dict_embed_text(row: np.ndarray) -> Dict[int, float]:
    """embedding_providers.py:142-145 -- keep entries with abs(weight) > 1e-6."""
    return {int(i): float(row[i]) for i in np.nonzero(np.abs(row) > 1e-6)[0]}


def to_dicts_embed_batch(rows: np.ndarray) -> List[Dict[int, float]]:
    """embedding_providers.py:161-163 -- keep entries != 0 (np.nonzero)."""
    out = []
    for emb in rows:
        idx = np.nonzero(emb)[0]
        out.append({int(i): float(emb[i]) for i in idx})
    return out
