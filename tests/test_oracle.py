"""Pin the oracle (oracle/knn_oracle.c, the C restatement of the reference's query path):
  * against every golden fixture produced by the UNMODIFIED reference (tests/golden/make_golden.py);
  * against the assertions the reference's own tests hold for this path (lib.zig:1292-1299, :1419-1424);
  * and, where oracle/_ref exists (this container and, via the prebuilt .so, the GPU box), live against
    the reference itself on fresh random inputs.
CPU only."""
import glob
from pathlib import Path

import numpy as np
import pytest

from helpers import ATOL, ATOL_COSINE, assert_knn_matches, recall
from oracle import oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
SEQ_CASES = sorted(Path(p).stem for p in glob.glob(str(GOLDEN / "seq_*.npz")))
HNSW_CASES = sorted(Path(p).stem for p in glob.glob(str(GOLDEN / "hnsw_*.npz")))


def test_golden_fixtures_exist():
    assert len(SEQ_CASES) >= 10 and len(HNSW_CASES) >= 3


@pytest.mark.parametrize("case", SEQ_CASES)
def test_seq_oracle_matches_reference_golden(case):
    g = np.load(GOLDEN / f"{case}.npz")
    space, k = str(g["space"]), int(g["k"])
    ids, d, c = O.seq_knn(str(g["ref_space"]), g["data"], g["queries"], k, g["ids"])
    dist_of = lambda q, i: O.pair_distance(str(g["ref_space"]),
                                           g["data"][np.nonzero(g["ids"] == i)[0][0]], g["queries"][q])
    assert_knn_matches(ids, d, c, g["ref_ids"], g["ref_dists"], g["ref_counts"],
                       exact=(space == "l2sqr_sift"), dist_of=dist_of, what=case)
    if space == "l2sqr":   # the new space: same ids as l2, distances = l2 squared
        ids2, d2, c2 = O.seq_knn("l2sqr", g["data"], g["queries"], k, g["ids"])
        sq = (g["ref_dists"].astype(np.float64) ** 2).astype(np.float32)
        assert_knn_matches(ids2, d2, c2, g["ref_ids"], sq, g["ref_counts"], what=case + "/l2sqr")


def test_reference_own_assertions():
    """lib.zig:1292-1299: l2, 3 unit vectors with ids 10/20/30, query = point 0, k = 2."""
    g = np.load(GOLDEN / "seq_libzig_unit3.npz")
    ids, d, c = O.seq_knn("l2", g["data"], g["queries"], 2, g["ids"])
    assert c[0] == 2 and ids[0, 0] == 10 and abs(d[0, 0]) < 1e-4
    assert abs(d[0, 1] - np.sqrt(2.0)) < 1e-6                      # lib.zig:1419-1424
    assert abs(O.pair_distance("l2", g["data"][0], g["data"][1]) - np.sqrt(2.0)) < 1e-6


@pytest.mark.parametrize("case", HNSW_CASES)
def test_hnsw_oracle_matches_reference_golden(case):
    g = np.load(GOLDEN / f"{case}.npz")
    h = O.PortHnsw(GOLDEN / f"{case}.hnsw")
    k = int(g["k"])
    for ef in g["efs"]:
        ids, d, c, evals = h.knn(g["queries"], k, int(ef))
        assert_knn_matches(ids, d, c, g[f"ids_ef{ef}"], g[f"dists_ef{ef}"], g[f"counts_ef{ef}"],
                           what=f"{case} ef={ef}")
        assert recall(ids, g["exact_ids"]) == recall(g[f"ids_ef{ef}"], g["exact_ids"])
        assert evals.min() > 0
    h.close()


def test_pair_distance_edge_cases():
    z = np.zeros(8, np.float32)
    x = np.arange(8, dtype=np.float32)
    assert O.pair_distance("cosinesimil", z, x) == 1.0               # distcomp_scalar.cc:150-155 -> nsp 0
    assert O.pair_distance("cosinesimil", x, x) <= 1e-6
    assert O.pair_distance("cosinesimil", x, -x) == 2.0
    assert O.pair_distance("negdotprod", x, x) == -140.0
    assert O.pair_distance("l2sqr", x, z) == 140.0
    a = np.full(128, 255, np.uint8)
    assert O.pair_distance("l2sqr_sift", a, np.zeros(128, np.uint8)) == 128 * 255 * 255   # SURVEY 0.9 max


needs_ref = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("space,dim", [("l2", 128), ("l2", 19), ("cosinesimil", 64), ("cosinesimil", 5),
                                       ("negdotprod", 768), ("negdotprod", 31)])
def test_seq_oracle_matches_live_reference(space, dim):
    rng = np.random.default_rng(100 + dim)
    x = rng.standard_normal((1500, dim)).astype(np.float32)
    q = rng.standard_normal((40, dim)).astype(np.float32)
    ids = rng.permutation(5000)[:1500].astype(np.int32)
    ri, rd, rc = O.RefIndex(space, "seq_search").add(x, ids).build("").knn(q, 25)
    pi, pd, pc = O.seq_knn(space, x, q, 25, ids)
    assert_knn_matches(pi, pd, pc, ri, rd, rc, what=f"{space}/{dim}")


@needs_ref
def test_sift_oracle_bit_exact_vs_live_reference():
    rng = np.random.default_rng(7)
    x = rng.integers(0, 256, (3000, 128), dtype=np.uint8)
    q = rng.integers(0, 256, (32, 128), dtype=np.uint8)
    ri, rd, rc = O.RefIndex("l2sqr_sift", "seq_search").add(x).build("").knn(q, 10)
    pi, pd, pc = O.seq_knn("l2sqr_sift", x, q, 10)
    assert np.array_equal(rd, pd) and np.array_equal(rc, pc)
    assert_knn_matches(pi, pd, pc, ri, rd, rc, exact=True, what="sift")


@needs_ref
@pytest.mark.parametrize("space,dim,params", [("l2", 48, "M=10,efConstruction=120"),
                                              ("cosinesimil", 40, "M=8,efConstruction=100"),
                                              ("negdotprod", 33, "M=12,efConstruction=100")])
def test_hnsw_oracle_matches_live_reference(space, dim, params, tmp_path):
    rng = np.random.default_rng(dim)
    cent = rng.standard_normal((12, dim)).astype(np.float32)
    x = (cent[rng.integers(0, 12, 4000)] + 0.3 * rng.standard_normal((4000, dim))).astype(np.float32)
    q = (cent[rng.integers(0, 12, 100)] + 0.3 * rng.standard_normal((100, dim))).astype(np.float32)
    r = O.RefIndex(space, "hnsw").add(x).build(params + ",indexThreadQty=4")
    path = tmp_path / "g.hnsw"
    r.save(path)
    h = O.PortHnsw(path)
    for ef in (10, 64, 300, 1200):            # 1200 -> SearchOld in hybrid mode (hnsw.cc:724)
        r.set_query_params(f"efSearch={ef}")
        ri, rd, rc = r.knn(q, 10)
        pi, pd, pc, _ = h.knn(q, 10, ef)
        assert_knn_matches(pi, pd, pc, ri, rd, rc, what=f"{space} ef={ef}")
    h.close()


def test_int_space_hnsw_golden_is_the_flat_search_over_exact_integer_distances(tmp_path):
    """SURVEY a18 pinned on the CPU: the reference's answers for l2sqr_sift + hnsw (Hnsw<int>: pointer graph,
    baseSearchAlgorithmV1Merge / Old, hnsw.cc:1076-1300; golden made by the reference itself) equal the oracle's port
    of the FLAT search run over the same links with the rows widened to float -- ids (ties aside) and int32 distances,
    at every efSearch incl. the SearchOld regime.  That equivalence is what the device path relies on."""
    g = np.load(GOLDEN / "regular_hnsw_sift.npz")
    flat = O.regular_to_flat_hnsw(GOLDEN / "regular_hnsw_sift.hnsw", g["data"], g["ids"], tmp_path / "flat.hnsw")
    h = O.PortHnsw(flat)
    assert (h.total, h.dim) == (3000, 128)
    k = int(g["k"])
    q = g["queries"].astype(np.float32)
    for ef in g["efs"]:
        ids, d, c, _ = h.knn(q, k, int(ef))
        assert_knn_matches(ids, d, c, g[f"ids_ef{ef}"], g[f"dists_ef{ef}"], g[f"counts_ef{ef}"], exact=True,
                           what=f"int-space hnsw ef={ef}")
    h.close()
