"""Seeded synthetic workloads of BASELINE.json's configs.

Counter-based RNG (numpy Philox) so that every harness -- tests, bench, the reference arm --
regenerates identical bytes from (config, seed) without shipping data files.  Generation is
chunked so that the 10 M-row configs do not need a second full-size temporary.
"""
from __future__ import annotations

import numpy as np


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=seed))


def uniform(n: int, dim: int, seed: int) -> np.ndarray:
    """C1: i.i.d. U[0,1) float32."""
    return _rng(seed).random((n, dim), dtype=np.float32)


def sift_like_u8(n: int, seed: int, dim: int = 128, chunk: int = 1 << 18) -> np.ndarray:
    """C4 (and the integer core of C2): SIFT-histogram-like bytes -- a clipped geometric
    distribution, mean about 27, support 0..218, most mass near zero."""
    rng = _rng(seed)
    out = np.empty((n, dim), np.uint8)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        v = rng.geometric(1.0 / 28.0, size=(e - s, dim)) - 1
        np.minimum(v, 218, out=v)
        out[s:e] = v.astype(np.uint8)
    return out


def sift_like_f32(n: int, seed: int, dim: int = 128) -> np.ndarray:
    """C2: the same integers cast to float32 ("SIFT-shaped synthetic float")."""
    return sift_like_u8(n, seed, dim).astype(np.float32)


def gist_like(n: int, dim: int, seed: int, clusters: int = 64, centroid_seed: int = 5) -> np.ndarray:
    """C3: U[0,1) * exp-decaying per-dimension scale around `clusters` latent centroids
    (centroid + N(0, 0.05)) so that the HNSW graph is navigable.  Data and queries share the
    centroids (centroid_seed) and differ in `seed`."""
    crng = _rng(centroid_seed)
    scale = np.exp(-np.arange(dim, dtype=np.float32) / (dim / 3.0)).astype(np.float32)
    cent = (crng.random((clusters, dim), dtype=np.float32) * scale).astype(np.float32)
    rng = _rng(seed)
    which = rng.integers(0, clusters, size=n)
    x = cent[which] + (0.05 * rng.standard_normal((n, dim), dtype=np.float32)) * scale
    return np.ascontiguousarray(x, np.float32)


def embedding_like(n: int, dim: int, seed: int, chunk: int = 1 << 17, rows=None) -> np.ndarray:
    """C5: N(0,1) rows, L2-normalised, scaled by lognormal(0, 0.1).  Every chunk of 131 072 rows has its own
    counter-based stream (Philox key (seed, chunk)), so `rows=(lo, hi)` regenerates just those rows of the same
    n-row data set -- a rank of the 8-way sharded 10 M x 768 run makes its 3.8 GB shard, not the 30.7 GB whole."""
    lo, hi = rows if rows is not None else (0, n)
    out = np.empty((hi - lo, dim), np.float32)
    for c in range(lo // chunk, (hi + chunk - 1) // chunk):
        s, e = c * chunk, min(n, (c + 1) * chunk)
        rng = _rng(seed) if c == 0 else np.random.Generator(np.random.Philox(key=[seed, c]))
        x = rng.standard_normal((e - s, dim), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x *= rng.lognormal(0.0, 0.1, size=(e - s, 1)).astype(np.float32)
        a, b = max(lo, s), min(hi, e)
        out[a - lo:b - lo] = x[a - s:b - s]
    return out


CONFIGS = {
    # name: (space, method, data_type, dist_type, n, dim, nq, k, db_seed, q_seed)
    "c1": ("l2", "seq_search", "DenseVector", "Float", 10_000, 128, 1_000, 10, 1, 2),
    "c2": ("l2sqr", "seq_search", "DenseVector", "Float", 1_000_000, 128, 10_000, 10, 3, 4),
    "c3": ("cosinesimil", "hnsw", "DenseVector", "Float", 1_000_000, 960, 10_000, 10, 5, 6),
    "c4": ("l2sqr_sift", "seq_search", "DenseUInt8Vector", "Int", 10_000_000, 128, 10_000, 10, 7, 8),
    "c5": ("negdotprod", "seq_search", "DenseVector", "Float", 10_000_000, 768, 100_000, 100, 9, 10),
}


def make(config: str, n: int | None = None, nq: int | None = None, rows=None):
    """(database, queries) of a BASELINE config, optionally at a reduced row / query count.  rows=(lo, hi): only
    that shard of the database (generated without the rest where the generator allows it: c5)."""
    space, method, dtype, dist, n0, dim, nq0, k, s_db, s_q = CONFIGS[config]
    n = n or n0
    nq = nq or nq0
    if rows is not None:
        if config == "c5":
            return embedding_like(n, dim, s_db, rows=rows), embedding_like(nq, dim, s_q)
        data, queries = make(config, n, nq)
        return np.ascontiguousarray(data[rows[0]:rows[1]]), queries
    if config == "c1":
        return uniform(n, dim, s_db), uniform(nq, dim, s_q)
    if config == "c2":
        return sift_like_f32(n, s_db, dim), sift_like_f32(nq, s_q, dim)
    if config == "c3":
        return gist_like(n, dim, s_db), gist_like(nq, dim, s_q)
    if config == "c4":
        return sift_like_u8(n, s_db, dim), sift_like_u8(nq, s_q, dim)
    if config == "c5":
        return embedding_like(n, dim, s_db), embedding_like(nq, dim, s_q)
    raise KeyError(config)
