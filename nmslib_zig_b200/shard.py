"""Row-wise database sharding across ranks -- the host-side bookkeeping.

One process per GPU.  Rank r owns the contiguous rows [lo, hi) of the database and reports
positions as `lo + local row` (Index.setShard), so that the 64-bit result keys
`ordered(distance) << 32 | position` sort globally exactly like the reference's single
KNNQueue would: by distance, ties by insertion position.  Per batch every rank
scans its shard, the per-shard (key, id) lists are exchanged with ONE all-gather over
NCCL/NVLink, and the k-way merge runs on device (csrc/topk_merge.cu, Index.mergeTopk).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced row range of `rank` (the reference's own thread split is contiguous
    as well, seqsearch.cc:73-85)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def ordered_f32(d: np.ndarray) -> np.ndarray:
    """Order-preserving uint32 image of float32 (csrc/common.cuh f32_ordered); NaN sorts last."""
    d = np.ascontiguousarray(d, np.float32) + np.float32(0.0)
    b = d.view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)) != 0
    out = np.where(neg, b ^ np.uint64(0xFFFFFFFF), b ^ np.uint64(0x80000000))
    out = np.where(np.isnan(d), np.uint64(0xFFFFFFFF), out)
    return out.astype(np.uint64)


def make_keys(dists: np.ndarray, positions: np.ndarray) -> np.ndarray:
    """(distance, global position) -> uint64 keys; missing entries (position < 0) become KEY_MAX."""
    keys = (ordered_f32(dists) << np.uint64(32)) | positions.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    return np.where(positions < 0, np.uint64(0xFFFFFFFFFFFFFFFF), keys)


def merge_gathered_host(keys: np.ndarray, ids: np.ndarray, k: int):
    """Host model of the device k-way merge: keys / ids are [world, q, k] as all_gather lays them
    out; returns ([q, k] ids, [q, k] keys) of the k smallest keys per query.  Used by the CPU-side
    multi-rank tests as the specification of csrc/topk_merge.cu."""
    world, q, kk = keys.shape
    flat_k = np.transpose(keys, (1, 0, 2)).reshape(q, world * kk)
    flat_i = np.transpose(ids, (1, 0, 2)).reshape(q, world * kk)
    order = np.argsort(flat_k, axis=1, kind="stable")[:, :k]
    out_k = np.take_along_axis(flat_k, order, axis=1)
    out_i = np.take_along_axis(flat_i, order, axis=1)
    out_i = np.where(out_k == np.uint64(0xFFFFFFFFFFFFFFFF), -1, out_i)
    return out_i, out_k
