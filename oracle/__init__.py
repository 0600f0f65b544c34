"""
CPU oracle for the verbatim-rag hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``verbatim_rag_b200``) never imports it and has no CPU fallback.

PARITY STATUS: *unpinned by the reference*.  The reference's arithmetic on this
path lives in third-party packages that are not vendored in /root/reference and
not installable offline:

* HF remote code of ``KRLabsOrg/verbatim-rag-modern-bert-v2`` (``model.process``),
  called at packages/core/verbatim_core/extractors.py:213-221;
* ``sentence_transformers.SparseEncoder`` 5.x (pin: docker/constraints.txt:409),
  called at verbatim_rag/embedding_providers.py:140,151;
* ``pymilvus`` 2.6.17 + ``milvus-lite`` 2.5.1 (FLAT / SPARSE_INVERTED_INDEX search),
  called at verbatim_rag/vector_stores/milvus_base.py:240-259.

and no reference test holds a golden vector for it.  What the oracle IS pinned
against (tests/test_oracle_pin.py): the installed ``transformers`` 5.5.0
``ModernBertForTokenClassification`` and ``BertForMaskedLM`` classes (the same model
classes the third-party code instantiates) running the same seeded weights, and
the reference's own in-repo post-processing (dict conversion, thresholds, hybrid
RRF merge) imported from /root/reference when present.  Golden fixtures in
tests/golden/ are produced by ``tests/golden/make_golden.py`` from this oracle.
"""
