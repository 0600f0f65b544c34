"""Oracle-backed stand-ins for ``verbatim_rag_b200._native`` handles (TEST INFRASTRUCTURE).

The B200 plugins need a GPU; the reference's Python only exists in the build container, which has none.  To run the
plugins' HOST logic (tokenisation, windowing, payload handling, branch structure, result shaping) through the
reference's real VerbatimIndex / VerbatimRAG on CPU, these fakes answer the C-ABI calls with the CPU oracle.
They are never importable from the product package."""
import numpy as np

from oracle import bert_splade, flat_topk, modernbert


class FakeContext:
    device = 0
    launches = 0

    def sync(self):
        pass


class FakeEncoder:
    def __init__(self, ctx, kind, weights, num_layers, vocab_size, max_tokens=65536, precision="fast"):
        from verbatim_rag_b200.synthetic import BertSpec, ModernBertSpec
        self.kind, self.weights, self.vocab_size = kind, weights, vocab_size
        self.spec = (ModernBertSpec(layers=num_layers, vocab_size=vocab_size) if kind == 0
                     else BertSpec(layers=num_layers, vocab_size=vocab_size))

    @staticmethod
    def _split(ids, cu):
        return [np.asarray(ids[cu[i]:cu[i + 1]], dtype=np.int64) for i in range(len(cu) - 1)]

    def span_forward(self, ids, cu, want_logits=False):
        lg = np.concatenate(modernbert.modernbert_forward_varlen(self.weights, self._split(ids, cu), self.spec), axis=0)
        p = modernbert.relevant_prob(lg)
        return (p, lg) if want_logits else p

    def splade_forward(self, ids, cu, min_abs=0.0, want_dense=False, want_csr=True):
        dense = bert_splade.splade_encode(self.weights, self._split(ids, cu), self.spec)
        indptr = np.zeros(len(cu), np.int64)
        idx, val = [], []
        for i, row in enumerate(dense):
            nz = np.nonzero(row > min_abs)[0]
            idx.append(nz.astype(np.int32))
            val.append(row[nz])
            indptr[i + 1] = indptr[i] + len(nz)
        out = {"indptr": indptr, "indices": np.concatenate(idx), "values": np.concatenate(val)}
        if want_dense:
            out["dense"] = dense
        return out

    def dense_forward(self, ids, cu, pooling=0, normalize=True):
        return bert_splade.dense_encode(self.weights, self._split(ids, cu), self.spec,
                                        pooling="mean" if pooling == 0 else "cls", normalize=bool(normalize))

    def close(self):
        pass


class FakeIndex:
    def __init__(self, ctx, kind, dim):
        self.ctx, self.kind, self.dim = ctx, kind, dim
        self.rows = np.zeros((0, dim), np.float32)
        self.indptr, self.indices, self.values = np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float32)
        self.dead = set()
        self.excluded = None
        self.id_base = 0

    def __len__(self):
        return len(self.rows) if self.kind == 0 else len(self.indptr) - 1

    def set_id_base(self, b):
        self.id_base = b

    def add_dense(self, rows):
        self.excluded = None
        self.rows = np.concatenate([self.rows, np.asarray(rows, np.float32)], axis=0)

    def add_sparse(self, indptr, indices, values):
        self.excluded = None
        a, b = indptr[0], indptr[-1]
        self.indptr = np.concatenate([self.indptr, self.indptr[-1] + (np.asarray(indptr[1:]) - a)])
        self.indices = np.concatenate([self.indices, np.asarray(indices[a:b], np.int32)])
        self.values = np.concatenate([self.values, np.asarray(values[a:b], np.float32)])

    def mark_deleted(self, rows):
        self.dead.update(int(r) for r in rows)
        self.excluded = None

    def set_filter(self, exclude=None):
        if exclude is not None:
            assert len(exclude) == len(self), "mask length must equal the number of rows"
        self.excluded = None if exclude is None else np.asarray(exclude, np.uint8).copy()

    def _finish(self, sc, k):
        if self.dead:
            sc[:, sorted(self.dead)] = -np.inf
        if self.excluded is not None:
            sc[:, self.excluded != 0] = -np.inf
        nq, n = sc.shape
        ids = np.full((nq, k), -1, np.int64)
        out = np.full((nq, k), -np.inf, np.float64)
        for q in range(nq):
            o = np.lexsort((np.arange(n), -sc[q]))[:k]
            o = o[np.isfinite(sc[q][o])]
            ids[q, :len(o)] = o + self.id_base
            out[q, :len(o)] = sc[q][o]
        return ids, out.astype(np.float32), out

    def search_dense(self, queries, k, want64=False):
        q = np.asarray(queries, np.float32).reshape(-1, self.dim)
        if len(self.rows) == 0:
            r = (np.full((len(q), k), -1, np.int64), np.full((len(q), k), -np.inf, np.float32), np.full((len(q), k), -np.inf))
        else:
            r = self._finish(flat_topk.dense_cosine_scores(self.rows, q), k)
        return r if want64 else r[:2]

    def search_sparse(self, q_indptr, q_indices, q_values, k, want64=False):
        qs = [{int(t): float(v) for t, v in zip(q_indices[q_indptr[i]:q_indptr[i + 1]], q_values[q_indptr[i]:q_indptr[i + 1]])}
              for i in range(len(q_indptr) - 1)]
        r = self._finish(flat_topk.sparse_ip_scores(self.indptr, self.indices, self.values, self.dim, qs), k)
        return r if want64 else r[:2]

    def close(self):
        pass


def install(monkeypatch):
    """Patch the native handles used by the plugin modules with the oracle-backed fakes."""
    from verbatim_rag_b200 import _native
    monkeypatch.setattr(_native, "default_context", lambda device=0: FakeContext())
    monkeypatch.setattr(_native, "Encoder", FakeEncoder)
    monkeypatch.setattr(_native, "Index", FakeIndex)
