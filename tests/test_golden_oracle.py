"""The oracle reproduces the committed golden fixtures (tests/golden/*.npz, made by make_golden.py) -- CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402
from oracle import bert_splade, flat_topk, highlighter, modernbert  # noqa: E402

G = os.path.join(HERE, "golden")


def test_span_cfg1_subset():
    c = cases.span_cfg1()
    g = np.load(os.path.join(G, "span_cfg1.npz"))
    assert bytes(g["input_hash"]).decode() == c["hash"], "synthetic text generator drifted from the fixture"
    cu = g["cu"]
    pick = [0, 31]              # 2 of the 32 pairs (a full 22-layer CPU forward each)
    lg = modernbert.modernbert_forward_varlen(c["weights"], [c["seqs"][i] for i in pick], c["spec"], batch=1)
    for i, l in zip(pick, lg):
        assert np.abs(l - g["logits"][cu[i]:cu[i + 1]]).max() < 2e-4   # batch-size / thread-count reduction order
        p = modernbert.relevant_prob(g["logits"][cu[i]:cu[i + 1]])
        nq = c["n_q"][i]
        sp = highlighter.spans_from_token_probs(c["pairs"][i][1], p[nq + 2:nq + 2 + 128], c["ctx_offsets"][i], 0.2, 30, 20)
        exp = g["spans"][g["spans"][:, 0] == i]
        assert [(s["start"], s["end"], s["tok_start"], s["tok_end"]) for s in sp] == [tuple(r[1:]) for r in exp.tolist()]
        for s in sp:
            assert s["text"] == c["pairs"][i][1][s["start"]:s["end"]] and len(s["text"]) >= 30


def test_splade_small_subset():
    s = cases.splade_small()
    g = np.load(os.path.join(G, "splade_small.npz"))
    assert bytes(g["input_hash"]).decode() == s["hash"]
    pick = [2, 6, 11]
    dense = bert_splade.splade_encode(s["weights"], [s["seqs"][i] for i in pick], s["spec"])
    for r, i in zip(dense, pick):
        a, b = g["indptr"][i], g["indptr"][i + 1]
        ref = np.zeros(30522, np.float32)
        ref[g["indices"][a:b]] = g["values"][a:b]
        assert np.abs(r - ref).max() < 2e-4
        # embed_text's filter (abs > 1e-6) and embed_batch's (!= 0) agree on these vectors
        assert set(bert_splade.to_dict_embed_text(r)) == set(bert_splade.to_dicts_embed_batch(r[None])[0])


def test_topk_goldens():
    d = cases.topk_dense()
    g = np.load(os.path.join(G, "topk_dense.npz"))
    ids, sc = flat_topk.dense_cosine_topk(d["corpus"], d["queries"], d["k"])
    assert np.array_equal(ids, g["ids"]) and np.array_equal(sc, g["scores"])
    assert np.all(np.diff(sc, axis=1) <= 0)
    t = cases.topk_sparse()
    g2 = np.load(os.path.join(G, "topk_sparse.npz"))
    ids2, sc2 = flat_topk.sparse_ip_topk(*t["corpus"], 30522, t["query_dicts"], t["k"])
    assert np.array_equal(ids2, g2["ids"]) and np.array_equal(sc2, g2["scores"])


def test_topk_tie_break_and_edges():
    corpus = np.ones((6, 8), np.float32)
    corpus[3] = 0.0
    ids, sc = flat_topk.dense_cosine_topk(corpus, np.ones((1, 8), np.float32), 10)
    assert ids.tolist() == [[0, 1, 2, 4, 5, 3]]            # ties by insertion order; zero vector scores 0
    assert sc[0, -1] == 0.0
    ip, idx, val = flat_topk.dicts_to_csr([{1: 1.0}, {2: 1.0}, {1: 2.0, 2: 1.0}, {}])
    ids, sc = flat_topk.sparse_ip_topk(ip, idx, val, 10, [{1: 1.0}], 3)
    assert ids.tolist() == [[2, 0, 1]] and sc.tolist() == [[2.0, 1.0, 0.0]]
