"""Host-side span logic: the native vrag_spans_from_probs and the window planner against the oracle's restatement
(oracle/highlighter.py steps 1, 4-8), including hypothesis-generated ragged / empty inputs."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import highlighter
from verbatim_rag_b200 import _native
from verbatim_rag_b200.extractor import plan_windows


@given(st.integers(1, 40), st.integers(0, 3000), st.integers(48, 600), st.integers(0, 300))
@settings(max_examples=200, deadline=None)
def test_plan_windows_matches_oracle(n_q, n_ctx, max_length, stride):
    try:
        ref = highlighter.plan_windows(n_q, n_ctx, max_length, stride)
    except ValueError:
        with pytest.raises(ValueError):
            plan_windows(n_q, n_ctx, max_length, stride)
        return
    got = plan_windows(n_q, n_ctx, max_length, stride)
    assert got == ref
    if n_ctx:
        covered = np.zeros(n_ctx, bool)
        for s, e in got:
            covered[s:e] = True
            assert e - s <= max_length - n_q - 3
        assert covered.all()


def _random_contexts(rng, nctx):
    probs, cs, ce, indptr, texts = [], [], [], [0], []
    for _ in range(nctx):
        n = int(rng.integers(0, 60))
        pos, words = 0, []
        for _ in range(n):
            pos += int(rng.integers(0, 4))        # whitespace / punctuation gap
            L = int(rng.integers(1, 9))
            cs.append(pos)
            ce.append(pos + L)
            pos += L
        p = rng.random(n).astype(np.float32)
        if n and rng.random() < 0.3:
            p[:] = 0.2                              # exactly the threshold: never kept (strict >)
        probs.append(p)
        indptr.append(indptr[-1] + n)
        texts.append("x" * (pos + 3))
    return (np.concatenate(probs) if probs else np.zeros(0, np.float32)), np.asarray(cs, np.int32), \
        np.asarray(ce, np.int32), np.asarray(indptr, np.int64), texts


@pytest.mark.parametrize("seed", range(25))
def test_native_spans_match_oracle(seed):
    rng = np.random.default_rng(seed)
    probs, cs, ce, indptr, texts = _random_contexts(rng, int(rng.integers(0, 12)))
    thr = float(rng.choice([0.2, 0.5, 0.05]))
    min_chars, gap = int(rng.integers(0, 40)), int(rng.integers(0, 25))
    got = _native.spans_from_probs(probs, cs, ce, indptr, thr, min_chars, gap)
    exp = []
    for c in range(len(indptr) - 1):
        a, b = indptr[c], indptr[c + 1]
        for sp in highlighter.spans_from_token_probs(texts[c], probs[a:b], list(zip(cs[a:b], ce[a:b])), thr, min_chars, gap):
            exp.append((c, sp["start"], sp["end"], sp["tok_start"], sp["tok_end"], sp["score"]))
    assert len(exp) == len(got["ctx"])
    for i, (c, s, e, ts, te, sc) in enumerate(exp):
        assert (got["ctx"][i], got["start"][i], got["end"][i], got["tok_start"][i], got["tok_end"][i]) == (c, s, e, ts, te)
        assert abs(got["score"][i] - sc) < 1e-6
