"""Tokenisation workers for the plugins' host pipeline.

The ``tokenizers`` library encodes batches in parallel without the GIL, but pulling ids and character offsets out of
its ``Encoding`` objects is Python-object work (two lists of several hundred ints / tuples per text) that holds the
GIL -- at 512-token chunks that alone caps a single process near 6 k texts/s, below what one B200 extracts.  Large
batches are therefore split over a few worker PROCESSES that return flat numpy arrays.  This module imports nothing
but numpy and tokenizers so that a spawned worker starts fast."""
from __future__ import annotations

from itertools import chain
from typing import List, Sequence, Tuple

import numpy as np

_TOK = None


def init(tok_json: str, rayon_threads: int = 0) -> None:
    global _TOK
    if rayon_threads > 0:   # the workers share the host cores: keep each one's encode_batch pool to its share
        import os
        os.environ["RAYON_NUM_THREADS"] = str(rayon_threads)
    from tokenizers import Tokenizer
    _TOK = Tokenizer.from_str(tok_json)
    _TOK.no_truncation()
    _TOK.no_padding()


def encode_with(tok, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (lens int32 [n], ids int32 [sum lens], offsets int32 [sum lens, 2]) without special tokens."""
    encs = tok.encode_batch(list(texts), add_special_tokens=False)
    lens = np.fromiter((len(e) for e in encs), np.int32, count=len(encs))
    total = int(lens.sum())
    ids = np.fromiter(chain.from_iterable(e.ids for e in encs), np.int32, count=total)
    offs = np.fromiter(chain.from_iterable(chain.from_iterable(e.offsets for e in encs)), np.int32, count=2 * total)
    return lens, ids, offs.reshape(-1, 2)


def encode(texts: List[str]):
    return encode_with(_TOK, texts)


class TokenizerWorkers:
    """Lazy pool of ``n`` spawned worker processes sharing one tokenizer definition (its JSON)."""

    def __init__(self, tok, n: int):
        self.n = int(n)
        self._tok = tok
        self._pool = None

    def _ensure(self):
        if self._pool is None:
            import multiprocessing as mp
            from concurrent.futures import ProcessPoolExecutor
            import os
            try:
                cores = len(os.sched_getaffinity(0))
            except AttributeError:
                cores = os.cpu_count() or 1
            self._pool = ProcessPoolExecutor(self.n, mp_context=mp.get_context("spawn"), initializer=init,
                                             initargs=(self._tok.to_str(), max(1, cores // self.n)))
        return self._pool

    def encode(self, texts: Sequence[str]):
        """Same result as ``encode_with(tok, texts)``, computed by the workers on contiguous slices."""
        texts = list(texts)
        if self.n <= 0 or len(texts) < 2 * self.n:
            return encode_with(self._tok, texts)
        step = (len(texts) + self.n - 1) // self.n
        try:
            pool = self._ensure()
            parts = [f.result() for f in [pool.submit(encode, texts[a:a + step]) for a in range(0, len(texts), step)]]
        except Exception as exc:  # noqa: BLE001 -- the pool is an optimisation, never a reason to lose a batch
            # Typical cause: the calling script spawns at import time (no ``if __name__ == "__main__":`` guard), or the
            # host forbids new processes.  Tokenise in-process from now on -- same result, one warning.
            import logging
            logging.getLogger(__name__).warning(
                "tokenizer worker processes unavailable (%s: %s); tokenising in-process from now on",
                type(exc).__name__, str(exc).strip().splitlines()[0] if str(exc).strip() else "")
            self.close()
            self.n = 0
            return encode_with(self._tok, texts)
        return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
                np.concatenate([p[2] for p in parts], axis=0))

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=False, cancel_futures=True)
            self._pool = None
