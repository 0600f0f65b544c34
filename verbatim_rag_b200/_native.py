"""ctypes binding of libvrag_b200.so (the C ABI declared in include/vrag_b200.h).

Loading is lazy and LOUD: a missing / unbuildable library or a missing CUDA device raises
``NativeError`` -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libvrag_b200.so")

VRAG_OK, VRAG_ERR_CUDA, VRAG_ERR_ARG, VRAG_ERR_CAPACITY, VRAG_ERR_WEIGHTS, VRAG_ERR_INTERNAL = range(6)
ENC_MODERNBERT_TOKCLS, ENC_BERT_MLM, ENC_BERT_DENSE, ENC_BERT_CLS, ENC_MODERNBERT_SENT = 0, 1, 2, 3, 4
INDEX_DENSE_COSINE, INDEX_SPARSE_IP = 0, 1
POOL_MEAN, POOL_CLS = 0, 1
PRECISION_FAST, PRECISION_PRECISE = 0, 1
PRECISIONS = {"fast": PRECISION_FAST, "precise": PRECISION_PRECISE}

# every symbol include/vrag_b200.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "vrag_ctx_create", "vrag_ctx_destroy", "vrag_last_error", "vrag_sync", "vrag_stream", "vrag_launch_count",
    "vrag_version", "vrag_profile", "vrag_profile_read", "vrag_encoder_create", "vrag_encoder_destroy", "vrag_span_forward", "vrag_splade_forward",
    "vrag_dense_forward", "vrag_selftest_gemm", "vrag_bench_gemm", "vrag_selftest_attention", "vrag_bench_attention", "vrag_debug_span_hidden", "vrag_spans_from_probs",
    "vrag_index_create", "vrag_index_destroy", "vrag_index_size", "vrag_index_add_dense", "vrag_index_add_sparse",
    "vrag_index_set_id_base", "vrag_index_mark_deleted", "vrag_index_set_filter", "vrag_index_search_dense", "vrag_index_search_sparse",
    "vrag_topk_merge",
    "vrag_topk_publish",
    "vrag_topk_merge_packed",
    "vrag_encoder_create_ex", "vrag_selftest_gemm_split", "vrag_bench_gemm_split", "vrag_selftest_attention_split",
    "vrag_bench_attention_split", "vrag_encoder_hidden", "vrag_rerank_forward", "vrag_sentence_forward",
    "vrag_span_extract",
]


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvrag_b200 error {code}: {msg}")
        self.code = code


class _Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


_lib = None
_lib_lock = threading.Lock()


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library (building it with nvcc first if it is absent)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        stamp = os.path.join(os.path.dirname(LIB_PATH), "build.sha256")
        stale = (os.path.exists(LIB_PATH)
                 and (not os.path.exists(stamp) or open(stamp).read().strip() != _build._digest()))
        if not os.path.exists(LIB_PATH) or stale:
            # a library built from other sources than the ones in the tree would be called with this module's argtypes
            if not build_if_missing:
                raise NativeError(VRAG_ERR_INTERNAL, f"{LIB_PATH} is {'stale' if stale else 'not built'}; "
                                                     "run python -m verbatim_rag_b200.build")
            try:
                _build.build()
            except Exception as exc:
                raise NativeError(VRAG_ERR_INTERNAL, f"{LIB_PATH} is {'stale' if stale else 'missing'} and could not "
                                                     f"be rebuilt: {exc}") from exc
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
        P = C.POINTER
        sig = {
            "vrag_ctx_create": (i32, [i32, P(vp)]),
            "vrag_ctx_destroy": (None, [vp]),
            "vrag_last_error": (C.c_char_p, [vp]),
            "vrag_sync": (i32, [vp]),
            "vrag_stream": (vp, [vp]),
            "vrag_launch_count": (C.c_uint64, [vp]),
            "vrag_version": (C.c_char_p, []),
            "vrag_profile": (i32, [vp, i32]),
            "vrag_profile_read": (i32, [vp, P(f64), P(i64)]),
            "vrag_encoder_create": (i32, [vp, i32, i32, i32, i32, P(_Tensor), i32, P(vp)]),
            "vrag_encoder_create_ex": (i32, [vp, i32, i32, i32, i32, P(_Tensor), i32, i32, P(vp)]),
            "vrag_encoder_destroy": (None, [vp]),
            "vrag_encoder_hidden": (i32, [vp]),
            "vrag_rerank_forward": (i32, [vp, vp, vp, vp, i32, vp, i32]),
            "vrag_sentence_forward": (i32, [vp, vp, vp, i32, vp, vp, vp, vp]),
            "vrag_span_extract": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, f32, i32, i32, vp, vp, vp, vp, vp, vp, i64, P(i64)]),
            "vrag_selftest_gemm_split": (i32, [vp, i32, i32, i32, i32, P(f64), P(f64)]),
            "vrag_bench_gemm_split": (i32, [vp, i32, i32, i32, i32, i32, P(f64)]),
            "vrag_selftest_attention_split": (i32, [vp, vp, vp, vp, i32, i32, vp, vp]),
            "vrag_bench_attention_split": (i32, [vp, i32, i32, i32, i32, vp]),
            "vrag_span_forward": (i32, [vp, vp, vp, i32, vp, vp, i32]),
            "vrag_debug_span_hidden": (i32, [vp, vp, vp, i32, vp, vp, vp]),
            "vrag_splade_forward": (i32, [vp, vp, vp, i32, f32, vp, vp, vp, i64, P(i64), vp, i32]),
            "vrag_dense_forward": (i32, [vp, vp, vp, i32, i32, i32, vp, i32]),
            "vrag_selftest_gemm": (i32, [vp, i32, i32, i32, i32, P(f64), P(f64)]),
            "vrag_bench_gemm": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, P(f64)]),
            "vrag_selftest_attention": (i32, [vp, vp, vp, i32, i32, i32, vp]),
            "vrag_bench_attention": (i32, [vp, i32, i32, i32, i32, vp]),
            "vrag_spans_from_probs": (i32, [vp, vp, vp, vp, i32, f32, i32, i32, vp, vp, vp, vp, vp, vp, i64, P(i64)]),
            "vrag_index_create": (i32, [vp, i32, i32, P(vp)]),
            "vrag_index_destroy": (None, [vp]),
            "vrag_index_size": (i64, [vp]),
            "vrag_index_add_dense": (i32, [vp, vp, i64, i32]),
            "vrag_index_add_sparse": (i32, [vp, vp, vp, vp, i64]),
            "vrag_index_set_id_base": (i32, [vp, i64]),
            "vrag_index_mark_deleted": (i32, [vp, vp, i64]),
            "vrag_index_set_filter": (i32, [vp, vp, i64]),
            "vrag_index_search_dense": (i32, [vp, vp, i32, i32, vp, vp, vp, i32]),
            "vrag_index_search_sparse": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, vp]),
            "vrag_topk_merge": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp]),
            "vrag_topk_publish": (i32, [vp, vp, vp, i32, i32, vp, i32, i32]),
            "vrag_topk_merge_packed": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _ptr(a) -> Optional[int]:
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"cannot take a pointer of {type(a)}")


def _np(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One CUDA context handle = one device + one stream; thread-safe (the library serialises per handle)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.vrag_ctx_create(int(device), C.byref(h))
        if rc != VRAG_OK:
            raise NativeError(rc, (self.lib.vrag_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)

    def check(self, rc: int):
        if rc != VRAG_OK:
            raise NativeError(rc, (self.lib.vrag_last_error(self.h) or b"").decode())

    def sync(self):
        self.check(self.lib.vrag_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.vrag_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.vrag_launch_count(self.h))

    PROF_CLASSES = ("gemm", "attention", "rowops", "scan", "select", "other")

    def profile(self, enable: bool):
        self.check(self.lib.vrag_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        ms = (C.c_double * 6)()
        cnt = (C.c_int64 * 6)()
        self.check(self.lib.vrag_profile_read(self.h, ms, cnt))
        return {n: {"ms": ms[i], "launches": int(cnt[i])} for i, n in enumerate(self.PROF_CLASSES)}

    def selftest_gemm(self, M: int, N: int, K: int, epilogue: int = 10) -> Tuple[float, float]:
        """tcgen05 GEMM vs the SIMT reference GEMM on the device; epilogue 10 = fp32, 0 = fp16, 1 = RoPE-QKV,
        2 = fp32 residual add, 3 = GeGLU, 11 = residual add + fp16 copy + row moments, 12 / 13 = RoPE-QKV / GeGLU
        on rstd-scaled accumulators (deferred LayerNorm)."""
        d, m = C.c_double(), C.c_double()
        self.check(self.lib.vrag_selftest_gemm(self.h, M, N, K, epilogue, C.byref(d), C.byref(m)))
        return d.value, m.value

    def selftest_gemm_split(self, M: int, N: int, K: int, epilogue: int = 10) -> Tuple[float, float]:
        """Split-precision GEMM (hi / lo planes, three MMAs per product) vs the SIMT reference; fp16 outputs are
        compared as hi + lo sums.  Epilogues 10, 0, 1, 2, 3."""
        d, m = C.c_double(), C.c_double()
        self.check(self.lib.vrag_selftest_gemm_split(self.h, M, N, K, epilogue, C.byref(d), C.byref(m)))
        return d.value, m.value

    def bench_gemm_split(self, M: int, N: int, K: int, epilogue: int, iters: int = 10) -> float:
        ms = C.c_double()
        self.check(self.lib.vrag_bench_gemm_split(self.h, M, N, K, epilogue, iters, C.byref(ms)))
        return ms.value

    def bench_attention_split(self, nseq: int, seq_len: int, window: int = -1, iters: int = 10) -> float:
        ms = C.c_double()
        self.check(self.lib.vrag_bench_attention_split(self.h, nseq, seq_len, window, iters, C.byref(ms)))
        return ms.value

    def selftest_attention_split(self, qkv_f32, cu_seqlens, window: int = -1):
        """One split-precision attention launch on fp32 rows qkv [T, 2304] (split into hi / lo fp16 planes here);
        returns the fp32 sum of the output planes [T, 768]."""
        x = np.ascontiguousarray(qkv_f32, dtype=np.float32)
        hi = x.astype(np.float16)
        lo = (x - hi.astype(np.float32)).astype(np.float16)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        assert x.ndim == 2 and x.shape[1] == 2304 and x.shape[0] == int(cu[-1])
        oh = np.empty((x.shape[0], 768), dtype=np.float16)
        ol = np.empty((x.shape[0], 768), dtype=np.float16)
        self.check(self.lib.vrag_selftest_attention_split(self.h, hi.ctypes.data, lo.ctypes.data, cu.ctypes.data,
                                                          len(cu) - 1, int(window), oh.ctypes.data, ol.ctypes.data))
        return oh.astype(np.float32) + ol.astype(np.float32)

    def bench_gemm(self, M: int, N: int, K: int, epilogue: int, stages: int = 0, debug_mode: int = 0,
                   iters: int = 10) -> float:
        """Average launch time (ms) of one encoder GEMM shape on synthetic operands (CUDA events)."""
        ms = C.c_double()
        self.check(self.lib.vrag_bench_gemm(self.h, M, N, K, epilogue, stages, debug_mode, iters, C.byref(ms)))
        return ms.value

    def bench_attention(self, nseq: int, seq_len: int, window: int = -1, iters: int = 10) -> float:
        """Average launch time (ms) of the attention kernel on synthetic rows (CUDA events)."""
        ms = C.c_double()
        self.check(self.lib.vrag_bench_attention(self.h, nseq, seq_len, window, iters, C.byref(ms)))
        return ms.value

    def selftest_attention(self, qkv_f16, cu_seqlens, window: int = -1, legacy: bool = False):
        """One attention launch on host fp16 rows qkv_f16 [T, 2304] (q|k|v, 12 heads x 64); returns fp16 [T, 768]."""
        import numpy as np
        qkv = np.ascontiguousarray(qkv_f16, dtype=np.float16)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        assert qkv.ndim == 2 and qkv.shape[1] == 2304 and qkv.shape[0] == int(cu[-1])
        out = np.empty((qkv.shape[0], 768), dtype=np.float16)
        self.check(self.lib.vrag_selftest_attention(self.h, qkv.ctypes.data, cu.ctypes.data, len(cu) - 1, int(window),
                                                    1 if legacy else 0, out.ctypes.data))
        return out

    def topk_merge(self, scores64, ids, nq: int, m: int, k: int, ids_out, scores_out, scores64_out=None):
        """Device buffers (torch tensors or raw pointers)."""
        self.check(self.lib.vrag_topk_merge(self.h, _ptr(scores64), _ptr(ids), nq, m, k, _ptr(ids_out),
                                            _ptr(scores_out), _ptr(scores64_out)))

    def topk_publish(self, scores64, ids, nq: int, k: int, peer_ptrs, rank: int):
        """Store this rank's [nq, k] (score64, id) block into slot ``rank`` of every peer's exchange buffer
        (``peer_ptrs``: one device address per rank, mapped into this process)."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self.check(self.lib.vrag_topk_publish(self.h, _ptr(scores64), _ptr(ids), nq, k, arr, len(peer_ptrs), rank))

    def topk_merge_packed(self, packed, world: int, nq: int, k: int, ids_out, scores_out, scores64_out=None):
        self.check(self.lib.vrag_topk_merge_packed(self.h, _ptr(packed), world, nq, k, _ptr(ids_out), _ptr(scores_out),
                                                   _ptr(scores64_out)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.vrag_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Dict[int, Context] = {}
_ctx_lock = threading.Lock()


def default_context(device: int = 0) -> Context:
    with _ctx_lock:
        if device not in _default_ctx:
            _default_ctx[device] = Context(device)
        return _default_ctx[device]


class Encoder:
    def __init__(self, ctx: Context, kind: int, weights: Dict[str, np.ndarray], num_layers: int, vocab_size: int,
                 max_tokens: int = 65536, precision: str = "fast"):
        """precision: "fast" = fp16 tensor-core operands (logits within ~3.5e-3 of the fp32 reference);
        "precise" = split-precision operands, three MMAs per product (logits within 1e-3, ~3x the time)."""
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.ctx = ctx
        self.kind = kind
        self.precision = precision
        keep = []
        arr = (_Tensor * len(weights))()
        for i, (name, w) in enumerate(weights.items()):
            a = _np(w, np.float32)
            keep.append(a)
            arr[i].name = name.encode()
            arr[i].data = a.ctypes.data
            arr[i].numel = a.size
        h = C.c_void_p()
        ctx.check(ctx.lib.vrag_encoder_create_ex(ctx.h, kind, num_layers, vocab_size, max_tokens, arr, len(weights),
                                                 PRECISIONS[precision], C.byref(h)))
        self.h = h
        self.num_layers = num_layers
        self.vocab_size = vocab_size
        self.hidden = int(ctx.lib.vrag_encoder_hidden(h))

    @staticmethod
    def _pack(seqs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        cu = np.zeros(len(seqs) + 1, dtype=np.int32)
        if len(seqs):
            np.cumsum([len(s) for s in seqs], out=cu[1:])
        ids = np.concatenate([np.asarray(s, dtype=np.int32) for s in seqs]) if len(seqs) else np.zeros(0, np.int32)
        return np.ascontiguousarray(ids, dtype=np.int32), cu

    def span_forward(self, ids: np.ndarray, cu: np.ndarray, want_logits: bool = False):
        """Host arrays in, host arrays out (the e2e path: H2D of ids and D2H of probs happen inside the call)."""
        ids = _np(ids, np.int32)
        cu = _np(cu, np.int32)
        T = int(cu[-1])
        probs = np.empty(T, dtype=np.float32)
        logits = np.empty((T, 2), dtype=np.float32) if want_logits else None
        self.ctx.check(self.ctx.lib.vrag_span_forward(self.h, _ptr(ids), _ptr(cu), len(cu) - 1, _ptr(probs),
                                                      _ptr(logits), 0))
        return (probs, logits) if want_logits else probs

    def span_forward_device(self, ids_dev, cu: np.ndarray, probs_dev, logits_dev=None):
        """ids / outputs are device buffers (torch CUDA tensors); asynchronous on the context's stream."""
        cu = _np(cu, np.int32)
        self.ctx.check(self.ctx.lib.vrag_span_forward(self.h, _ptr(ids_dev), _ptr(cu), len(cu) - 1, _ptr(probs_dev),
                                                      _ptr(logits_dev), 1))

    def debug_span_hidden(self, ids: np.ndarray, cu: np.ndarray):
        ids = _np(ids, np.int32)
        cu = _np(cu, np.int32)
        T = int(cu[-1])
        probs = np.empty(T, dtype=np.float32)
        logits = np.empty((T, 2), dtype=np.float32)
        hidden = np.empty((self.num_layers + 1, T, 768), dtype=np.float32)
        self.ctx.check(self.ctx.lib.vrag_debug_span_hidden(self.h, _ptr(ids), _ptr(cu), len(cu) - 1, _ptr(probs),
                                                           _ptr(logits), _ptr(hidden)))
        return probs, logits, hidden

    def splade_forward(self, ids: np.ndarray, cu: np.ndarray, min_abs: float = 0.0, want_dense: bool = False,
                       want_csr: bool = True):
        ids = _np(ids, np.int32)
        cu = _np(cu, np.int32)
        n = len(cu) - 1
        dense = np.empty((n, self.vocab_size), dtype=np.float32) if want_dense else None
        cap = max(1024, n * 512)
        nnz = C.c_int64()
        while True:
            indptr = np.zeros(n + 1, dtype=np.int64) if want_csr else None
            indices = np.empty(cap, dtype=np.int32) if want_csr else None
            values = np.empty(cap, dtype=np.float32) if want_csr else None
            rc = self.ctx.lib.vrag_splade_forward(self.h, _ptr(ids), _ptr(cu), n, float(min_abs), _ptr(indptr),
                                                  _ptr(indices), _ptr(values), cap if want_csr else 0, C.byref(nnz),
                                                  _ptr(dense), 0)
            if rc == VRAG_ERR_CAPACITY:
                cap = int(nnz.value)
                continue
            self.ctx.check(rc)
            break
        out = {}
        if want_csr:
            out["indptr"], out["indices"], out["values"] = indptr, indices[:nnz.value], values[:nnz.value]
        if want_dense:
            out["dense"] = dense
        return out

    def splade_forward_device(self, ids_dev, cu: np.ndarray, dense_dev):
        """Device ids -> device dense [nseq, vocab] fp32 (no CSR)."""
        cu = _np(cu, np.int32)
        nnz = C.c_int64()
        self.ctx.check(self.ctx.lib.vrag_splade_forward(self.h, _ptr(ids_dev), _ptr(cu), len(cu) - 1, 0.0, None, None,
                                                        None, 0, C.byref(nnz), _ptr(dense_dev), 1))

    def dense_forward(self, ids: np.ndarray, cu: np.ndarray, pooling: int = POOL_MEAN, normalize: bool = True):
        ids = _np(ids, np.int32)
        cu = _np(cu, np.int32)
        out = np.empty((len(cu) - 1, self.hidden), dtype=np.float32)
        self.ctx.check(self.ctx.lib.vrag_dense_forward(self.h, _ptr(ids), _ptr(cu), len(cu) - 1, pooling,
                                                       1 if normalize else 0, _ptr(out), 0))
        return out

    def span_extract(self, ids, cu, ctx_first, ctx_len, tok_cs, tok_ce, threshold: float, min_span_chars: int,
                     merge_gap_chars: int):
        """Forward + span post-processing on the device (one sequence per context); same dict as ``spans_from_probs``
        with ``ctx`` = sequence index."""
        ids, cu = _np(ids, np.int32), _np(cu, np.int32)
        cf, cl = _np(ctx_first, np.int32), _np(ctx_len, np.int32)
        tcs, tce = _np(tok_cs, np.int32), _np(tok_ce, np.int32)
        nseq = len(cu) - 1
        cap = max(16, int(cl.sum()) // 8 + nseq)
        n = C.c_int64()
        while True:
            sc, cs, ce, ts, te = (np.empty(cap, np.int32) for _ in range(5))
            score = np.empty(cap, np.float32)
            rc = self.ctx.lib.vrag_span_extract(self.h, _ptr(ids), _ptr(cu), nseq, _ptr(cf), _ptr(cl), _ptr(tcs), _ptr(tce),
                                                float(threshold), int(min_span_chars), int(merge_gap_chars), _ptr(sc), _ptr(cs),
                                                _ptr(ce), _ptr(score), _ptr(ts), _ptr(te), cap, C.byref(n))
            if rc == VRAG_ERR_CAPACITY:
                cap = int(n.value)
                continue
            self.ctx.check(rc)
            break
        k = int(n.value)
        return {"ctx": sc[:k], "start": cs[:k], "end": ce[:k], "score": score[:k], "tok_start": ts[:k], "tok_end": te[:k]}

    def rerank_forward(self, ids: np.ndarray, type_ids: np.ndarray, cu: np.ndarray) -> np.ndarray:
        """Cross-encoder relevance logits, one per packed pair sequence (ENC_BERT_CLS encoders)."""
        ids, type_ids, cu = _np(ids, np.int32), _np(type_ids, np.int32), _np(cu, np.int32)
        out = np.empty(len(cu) - 1, dtype=np.float32)
        self.ctx.check(self.ctx.lib.vrag_rerank_forward(self.h, _ptr(ids), _ptr(type_ids), _ptr(cu), len(cu) - 1, _ptr(out), 0))
        return out

    def sentence_forward(self, ids: np.ndarray, cu: np.ndarray, sent_indptr: np.ndarray, sent_start: np.ndarray,
                         sent_end: np.ndarray) -> np.ndarray:
        """Sentence logits [n_sentences, 2] of the legacy QAModel head (ENC_MODERNBERT_SENT encoders)."""
        ids, cu = _np(ids, np.int32), _np(cu, np.int32)
        ip, s0, s1 = _np(sent_indptr, np.int32), _np(sent_start, np.int32), _np(sent_end, np.int32)
        out = np.empty((int(ip[-1]), 2), dtype=np.float32)
        self.ctx.check(self.ctx.lib.vrag_sentence_forward(self.h, _ptr(ids), _ptr(cu), len(cu) - 1, _ptr(ip), _ptr(s0),
                                                          _ptr(s1), _ptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.vrag_encoder_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spans_from_probs(probs: np.ndarray, tok_cs: np.ndarray, tok_ce: np.ndarray, ctx_indptr: np.ndarray,
                     threshold: float, min_span_chars: int, merge_gap_chars: int):
    """Host span post-processing (no GPU needed).  Returns dict of arrays, one entry per span."""
    lib = load_library()
    probs = _np(probs, np.float32)
    tok_cs = _np(tok_cs, np.int32)
    tok_ce = _np(tok_ce, np.int32)
    ctx_indptr = _np(ctx_indptr, np.int64)
    nctx = len(ctx_indptr) - 1
    cap = max(16, len(probs) // 2 + nctx)
    n = C.c_int64()
    while True:
        sc, cs, ce = (np.empty(cap, np.int32) for _ in range(3))
        ts, te = np.empty(cap, np.int32), np.empty(cap, np.int32)
        score = np.empty(cap, np.float32)
        rc = lib.vrag_spans_from_probs(_ptr(probs), _ptr(tok_cs), _ptr(tok_ce), _ptr(ctx_indptr), nctx,
                                       float(threshold), int(min_span_chars), int(merge_gap_chars), _ptr(sc), _ptr(cs),
                                       _ptr(ce), _ptr(score), _ptr(ts), _ptr(te), cap, C.byref(n))
        if rc == VRAG_ERR_CAPACITY:
            cap = int(n.value)
            continue
        if rc != VRAG_OK:
            raise NativeError(rc, "vrag_spans_from_probs: bad argument")
        break
    k = int(n.value)
    return {"ctx": sc[:k], "start": cs[:k], "end": ce[:k], "score": score[:k], "tok_start": ts[:k], "tok_end": te[:k]}


class Index:
    def __init__(self, ctx: Context, kind: int, dim: int):
        self.ctx = ctx
        self.kind = kind
        self.dim = dim
        h = C.c_void_p()
        ctx.check(ctx.lib.vrag_index_create(ctx.h, kind, dim, C.byref(h)))
        self.h = h

    def __len__(self) -> int:
        return int(self.ctx.lib.vrag_index_size(self.h))

    def set_id_base(self, base: int):
        self.ctx.check(self.ctx.lib.vrag_index_set_id_base(self.h, int(base)))

    def add_dense(self, rows):
        if isinstance(rows, np.ndarray) or not hasattr(rows, "data_ptr"):
            a = _np(rows, np.float32)
            assert a.ndim == 2 and a.shape[1] == self.dim, "rows must be [n, dim]"
            self.ctx.check(self.ctx.lib.vrag_index_add_dense(self.h, _ptr(a), a.shape[0], 0))
        else:  # torch CUDA tensor, fp32 contiguous
            assert rows.dim() == 2 and rows.shape[1] == self.dim and rows.is_contiguous()
            self.ctx.check(self.ctx.lib.vrag_index_add_dense(self.h, _ptr(rows), rows.shape[0], 1))

    def add_sparse(self, indptr, indices, values):
        indptr, indices, values = _np(indptr, np.int64), _np(indices, np.int32), _np(values, np.float32)
        self.ctx.check(self.ctx.lib.vrag_index_add_sparse(self.h, _ptr(indptr), _ptr(indices), _ptr(values),
                                                          len(indptr) - 1))

    def mark_deleted(self, rows):
        r = _np(rows, np.int64)
        self.ctx.check(self.ctx.lib.vrag_index_mark_deleted(self.h, _ptr(r), len(r)))

    def set_filter(self, exclude=None):
        """Metadata-filter pushdown: rows with exclude[row] != 0 are skipped by the following searches (None clears)."""
        if exclude is None:
            self.ctx.check(self.ctx.lib.vrag_index_set_filter(self.h, None, 0))
            return
        m = _np(exclude, np.uint8)
        self.ctx.check(self.ctx.lib.vrag_index_set_filter(self.h, _ptr(m), len(m)))

    def search_dense(self, queries: np.ndarray, k: int, want64: bool = False):
        q = _np(queries, np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        ids = np.empty((nq, k), np.int64)
        s32 = np.empty((nq, k), np.float32)
        s64 = np.empty((nq, k), np.float64) if want64 else None
        self.ctx.check(self.ctx.lib.vrag_index_search_dense(self.h, _ptr(q), nq, k, _ptr(ids), _ptr(s32), _ptr(s64), 0))
        return (ids, s32, s64) if want64 else (ids, s32)

    def search_dense_device(self, q_dev, nq: int, k: int, ids_dev, s32_dev, s64_dev=None):
        self.ctx.check(self.ctx.lib.vrag_index_search_dense(self.h, _ptr(q_dev), nq, k, _ptr(ids_dev), _ptr(s32_dev),
                                                            _ptr(s64_dev), 1))

    def search_sparse(self, q_indptr, q_indices, q_values, k: int, want64: bool = False):
        q_indptr, q_indices, q_values = _np(q_indptr, np.int64), _np(q_indices, np.int32), _np(q_values, np.float32)
        nq = len(q_indptr) - 1
        ids = np.empty((nq, k), np.int64)
        s32 = np.empty((nq, k), np.float32)
        s64 = np.empty((nq, k), np.float64) if want64 else None
        self.ctx.check(self.ctx.lib.vrag_index_search_sparse(self.h, _ptr(q_indptr), _ptr(q_indices), _ptr(q_values),
                                                             nq, k, _ptr(ids), _ptr(s32), _ptr(s64)))
        return (ids, s32, s64) if want64 else (ids, s32)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.vrag_index_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
