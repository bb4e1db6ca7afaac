"""Shared checkers for the parity tests."""
from __future__ import annotations

import numpy as np

RTOL = 1e-5   # BASELINE.json north_star: ids exact except at ties within 1e-5 relative distance
ATOL = 1e-6   # +1e-6 absolute for values near 0 (the reference's cosine clamps at 0)
# cosinesimil is d = 1 - nsp with nsp a unit-scale normalised scalar product: both the reference
# and any other fp32 evaluation carry ~sqrt(D)*2^-24 of absolute error in nsp (1.8e-6 at D = 960),
# so the 1e-5 relative tolerance applies to the unit scale of nsp, i.e. 1e-5 absolute on d.
ATOL_COSINE = 1e-5


def comparable(space: str, d):
    """Distances in the domain where the tolerance applies.  angulardist = acos(nsp) is ill-conditioned near
    0 (an nsp of 1 - 1 ulp is already 3.5e-4 rad, and the reference itself returns 0 or 3.5e-4 for identical
    vectors depending on how S / sqrt(S) / sqrt(S) rounds), so it is compared as 1 - cos(d), i.e. as the cosine
    distance it was computed from."""
    if space == "angulardist":
        return 1.0 - np.cos(np.asarray(d, np.float64))
    return d


def close(a, b, rtol=RTOL, atol=ATOL):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        return both_inf | (np.abs(a - b) <= rtol * np.abs(b) + atol)


def assert_knn_matches(ids, dists, counts, ref_ids, ref_dists, ref_counts, *, exact=False, dist_of=None,
                       what="", atol=None):
    """Tie-aware comparison of two kNN answers.

    exact=True  (integer spaces): distances bit-equal; ids equal except inside groups of EQUAL
                distance (the reference orders ties by heap address).
    exact=False (float spaces): |d - d_ref| <= RTOL*|d_ref| + ATOL; ids equal except inside groups
                of distances that are within that tolerance of each other.
    dist_of(q, id) -> distance, optional: used to validate an id that the reference list does not
                contain at all (possible only in the tie group cut by k).
    """
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    dists, ref_dists = np.asarray(dists), np.asarray(ref_dists)
    nq = ref_ids.shape[0]
    assert np.array_equal(np.asarray(counts).reshape(-1), np.asarray(ref_counts).reshape(-1)), f"{what}: counts differ"
    rtol, atol = (0.0, 0.0) if exact else (RTOL, ATOL if atol is None else atol)
    for q in range(nq):
        c = int(ref_counts[q])
        d, r = dists[q, :c], ref_dists[q, :c]
        ok = close(d, r, rtol, atol)
        assert ok.all(), f"{what}: query {q} distances differ: ours {d[~ok][:4]} ref {r[~ok][:4]}"
        assert np.all(np.diff(d.astype(np.float64)) >= 0), f"{what}: query {q} distances not ascending"
        i_ours, i_ref = ids[q, :c], ref_ids[q, :c]
        for j in np.nonzero(i_ours != i_ref)[0]:
            # (a) a swap inside a tie group of the reference list
            where = np.nonzero(i_ref == i_ours[j])[0]
            if where.size and close(r[where[0]], r[j], rtol, atol):
                continue
            # (b) a different member of the tie group that k cuts through
            if c > 0 and close(r[j], r[c - 1], rtol, atol) and not where.size:
                if dist_of is not None:
                    assert close(dist_of(q, int(i_ours[j])), r[j], rtol, atol), \
                        f"{what}: query {q} rank {j}: id {i_ours[j]} is not tied with the reference's k-th"
                continue
            raise AssertionError(f"{what}: query {q} rank {j}: id {i_ours[j]} != reference {i_ref[j]} "
                                 f"(d={d[j]!r}, ref d={r[j]!r})")


def recall(ids, exact_ids):
    """|approx ∩ exact| / |exact| (eval_metrics.h:112-128), averaged over queries."""
    hits = 0
    for a, e in zip(np.asarray(ids), np.asarray(exact_ids)):
        hits += len(set(a[a >= 0].tolist()) & set(e[e >= 0].tolist()))
    return hits / float(np.sum(np.asarray(exact_ids) >= 0))


def count_mismatches(ids, dists, counts, ref_ids, ref_dists, ref_counts, *, exact=False, atol=None):
    """Number of queries on which assert_knn_matches would fail (bench.py's `parity` record)."""
    bad = 0
    for q in range(np.asarray(ref_ids).shape[0]):
        try:
            assert_knn_matches(np.asarray(ids)[q:q + 1], np.asarray(dists)[q:q + 1], np.asarray(counts).reshape(-1)[q:q + 1],
                               np.asarray(ref_ids)[q:q + 1], np.asarray(ref_dists)[q:q + 1],
                               np.asarray(ref_counts).reshape(-1)[q:q + 1], exact=exact, atol=atol)
        except AssertionError:
            bad += 1
    return bad
