"""
Oracle: ModernBERT token-classifier forward, plain torch fp32 on CPU.  TEST INFRASTRUCTURE.

Restates what the reference runs through HF remote code at
packages/core/verbatim_core/extractors.py:151-157 (load) and :213-221 (``process``):
a ModernBERT-base encoder + prediction head + 2-way classifier.  The architecture
follows the installed third-party executable spec
``transformers/models/modernbert/modeling_modernbert.py`` (5.5.0): embeddings :52-71,
GeGLU MLP :74-91, RoPE :94-172 and :205-228, eager attention :175-194, attention
block :232-310, encoder layer :313-343, model :424-490, head :493-502, token
classification :672-724; sliding mask ``|i-j| <= 64`` per masking_utils.py:121-131.
tests/test_oracle_pin.py checks this file against that class on seeded weights.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _t(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def rope_tables(max_pos: int, head_dim: int, theta: float):
    """cos/sin [max_pos, head_dim] in fp32, 'cat(freqs, freqs)' convention (modeling_modernbert.py:148-172)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(max_pos, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def layer_norm(x, w, eps, b=None):
    return F.layer_norm(x, (x.shape[-1],), _t(w), None if b is None else _t(b), eps)


@torch.no_grad()
def modernbert_forward(weights: Dict[str, np.ndarray], input_ids, attention_mask=None, spec=None,
                       return_hidden: bool = False, emulate_fp16: bool = False, return_final: bool = False):
    """logits [B, L, 2] fp32 for padded ``input_ids`` [B, L] (attention_mask 1 = token, 0 = pad).

    ``emulate_fp16`` rounds every matmul operand to fp16 (fp32 accumulate) -- used only to
    predict the error of a 16-bit tensor-core pipeline on CPU, never as the oracle value.
    """
    from verbatim_rag_b200.synthetic import ModernBertSpec

    spec = spec or ModernBertSpec()
    ids = _t(input_ids).long()
    B, L = ids.shape
    if attention_mask is None:
        attention_mask = torch.ones(B, L, dtype=torch.long)
    am = _t(attention_mask).bool()
    H, nh, dh = spec.hidden, spec.heads, spec.head_dim

    def rq(x):  # fp16 operand rounding (emulation only)
        return x.half().float() if emulate_fp16 else x

    def lin(x, w, b=None):
        y = rq(x) @ rq(_t(w)).t()
        return y if b is None else y + _t(b)

    key_ok = am[:, None, None, :]                                   # [B,1,1,L] keys that exist
    idx = torch.arange(L)
    local_ok = (idx[:, None] - idx[None, :]).abs() <= spec.half_window
    neg = torch.finfo(torch.float32).min
    mask_global = torch.zeros(B, 1, L, L).masked_fill(~key_ok.expand(B, 1, L, L), neg)
    mask_local = torch.zeros(B, 1, L, L).masked_fill(~(key_ok & local_ok[None, None]).expand(B, 1, L, L), neg)
    cos_g, sin_g = rope_tables(L, dh, spec.theta_global)
    cos_l, sin_l = rope_tables(L, dh, spec.theta_local)

    x = layer_norm(_t(weights["model.embeddings.tok_embeddings.weight"])[ids],
                   weights["model.embeddings.norm.weight"], spec.norm_eps)
    hidden: List[torch.Tensor] = [x]
    for i in range(spec.layers):
        p = f"model.layers.{i}."
        h = x if i == 0 else layer_norm(x, weights[p + "attn_norm.weight"], spec.norm_eps)
        qkv = lin(h, weights[p + "attn.Wqkv.weight"]).view(B, L, 3, nh, dh)
        q, k, v = (qkv[:, :, j].transpose(1, 2) for j in range(3))   # [B,nh,L,dh]
        g = spec.is_global(i)
        cos, sin = (cos_g, sin_g) if g else (cos_l, sin_l)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        s = (rq(q) @ rq(k).transpose(2, 3)) * (dh ** -0.5) + (mask_global if g else mask_local)
        pr = torch.softmax(s, dim=-1, dtype=torch.float32)
        o = (rq(pr) @ rq(v)).transpose(1, 2).reshape(B, L, H)
        x = x + lin(o, weights[p + "attn.Wo.weight"])
        h = layer_norm(x, weights[p + "mlp_norm.weight"], spec.norm_eps)
        a, gate = lin(h, weights[p + "mlp.Wi.weight"]).chunk(2, dim=-1)
        x = x + lin(F.gelu(a) * gate, weights[p + "mlp.Wo.weight"])
        hidden.append(x)
    x = layer_norm(x, weights["model.final_norm.weight"], spec.norm_eps)
    if return_final:   # ModernBertModel.last_hidden_state (the input of the legacy QAModel's sentence pooling)
        return x
    hd = layer_norm(F.gelu(lin(x, weights["head.dense.weight"])), weights["head.norm.weight"], spec.norm_eps)
    logits = lin(hd, weights["classifier.weight"], weights["classifier.bias"])
    if return_hidden:
        return logits, hidden
    return logits


@torch.no_grad()
def modernbert_forward_varlen(weights, seqs: List[np.ndarray], spec=None, batch: int = 8, **kw) -> List[np.ndarray]:
    """Forward a list of unpadded id arrays (each its own length); returns per-sequence logits [L_i, 2]."""
    from verbatim_rag_b200.synthetic import ModernBertSpec

    spec = spec or ModernBertSpec()
    out: List[Optional[np.ndarray]] = [None] * len(seqs)
    order = sorted(range(len(seqs)), key=lambda i: -len(seqs[i]))
    for s in range(0, len(order), batch):
        grp = order[s:s + batch]
        L = max(len(seqs[i]) for i in grp)
        ids = np.full((len(grp), L), spec.pad_id, dtype=np.int64)
        am = np.zeros((len(grp), L), dtype=np.int64)
        for r, i in enumerate(grp):
            ids[r, :len(seqs[i])] = seqs[i]
            am[r, :len(seqs[i])] = 1
        lg = modernbert_forward(weights, ids, am, spec, **kw).numpy()
        for r, i in enumerate(grp):
            out[i] = lg[r, :len(seqs[i])].copy()
    return out  # type: ignore[return-value]


def relevant_prob(logits: np.ndarray) -> np.ndarray:
    """P(class 1) = softmax(logits)[..., 1] in fp32 (reference contract: SURVEY.md App. B.2 step 2)."""
    t = torch.from_numpy(np.ascontiguousarray(logits)).float()
    return torch.softmax(t, dim=-1)[..., 1].numpy()
