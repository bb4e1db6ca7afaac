"""Parity of the CUDA HNSW beam search against the reference on the SAME graph: the graphs in
tests/golden/*.hnsw were built and saved by the unmodified reference (Hnsw::SaveIndex); the
expected answers are the reference's own at each efSearch.  Gate (BASELINE.json north_star):
recall@10 at or above the reference's at the same efSearch on the same graph; in practice the
answers agree id for id up to float rounding."""
import glob
from pathlib import Path

import numpy as np
import pytest

import nmslib_zig_b200 as nb
from helpers import ATOL, ATOL_COSINE, assert_knn_matches, recall
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
HNSW_CASES = sorted(Path(p).stem for p in glob.glob(str(GOLDEN / "hnsw_*.npz")))


def agreement(ids, ref_ids):
    return float(np.mean(np.asarray(ids) == np.asarray(ref_ids)))


@pytest.mark.parametrize("case", HNSW_CASES)
def test_same_graph_matches_reference_golden(case):
    g = np.load(GOLDEN / f"{case}.npz")
    k = int(g["k"])
    idx = nb.Index(str(g["space"]), None, "hnsw")
    idx.importHnsw(GOLDEN / f"{case}.hnsw")
    for ef in g["efs"]:
        idx.setQueryTimeParams(nb.Params({"efSearch": int(ef)}))
        r = idx.knnQueryBatch(g["queries"], k)
        ref_ids, ref_d = g[f"ids_ef{ef}"], g[f"dists_ef{ef}"]
        rec_gpu, rec_ref = recall(r.ids, g["exact_ids"]), recall(ref_ids, g["exact_ids"])
        assert rec_gpu >= rec_ref - 1e-9, f"{case} ef={ef}: recall {rec_gpu} < reference {rec_ref}"
        assert agreement(r.ids, ref_ids) >= 0.995, f"{case} ef={ef}: id agreement {agreement(r.ids, ref_ids)}"
        assert_knn_matches(r.ids, r.distances, r.sizes, ref_ids, ref_d, g[f"counts_ef{ef}"], what=f"{case} ef={ef}",
                           atol=ATOL_COSINE if str(g["space"]).startswith("cos") else ATOL)
    s = idx.stats()
    assert s["distance_evals"] > 0 and s["hnsw_expansions"] > 0
    idx.deinit()


def test_default_ef_is_200_like_the_c_abi():
    """nmslib_c.cpp:330/:986 force efSearch=200; without an explicit setQueryTimeParams we match that."""
    g = np.load(GOLDEN / "hnsw_l2_d32.npz")
    idx = nb.Index.load(str(GOLDEN / "hnsw_l2_d32.hnsw"))
    r = idx.knnQueryBatch(g["queries"], 10)
    assert_knn_matches(r.ids, r.distances, r.sizes, g["ids_ef200"], g["dists_ef200"], g["counts_ef200"], what="ef200")
    one = idx.knnQuery(g["queries"][3], 10)
    assert np.array_equal(one.ids, r.ids[3])
    idx.deinit()


def test_l2_hnsw_reports_squared_distances():
    """l2 + hnsw returns the squared distance (hnsw.cc:374-385)."""
    g = np.load(GOLDEN / "hnsw_l2_d32.npz")
    idx = nb.Index.load(str(GOLDEN / "hnsw_l2_d32.hnsw"))
    idx.setQueryTimeParams(nb.Params({"efSearch": 200}))
    r = idx.knnQueryBatch(g["queries"], 10)
    exact_l2 = g["exact_dists"]                      # seq_search l2 = sqrt
    hit = r.ids[:, 0] == g["exact_ids"][:, 0]
    assert hit.mean() > 0.9
    assert np.allclose(r.distances[hit, 0], exact_l2[hit, 0].astype(np.float64) ** 2, rtol=1e-5, atol=1e-6)
    idx.deinit()


def test_hnsw_without_graph_fails_loudly():
    idx = nb.Index("l2", None, "hnsw")
    idx.addDenseBatch(np.eye(4, dtype=np.float32))
    idx.buildIndex()
    try:
        r = idx.knnQuery(np.ones(4, np.float32), 2)   # once the device builder exists this returns results
        assert len(r.ids) == 2
    except nb.NmslibError as e:
        assert e.name == "IndexBuildFailed"
    idx.deinit()


needs_ref = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("space,n,dim,params", [("l2", 20_000, 64, "M=16,efConstruction=200"),
                                                ("cosinesimil", 10_000, 200, "M=16,efConstruction=200"),
                                                ("negdotprod", 8_000, 48, "M=12,efConstruction=100")])
def test_same_graph_live_reference_ef_sweep(space, n, dim, params, tmp_path):
    """Build with the reference here and now (default M=16 -> maxM0=32, the shape of BASELINE config 3),
    export with SaveIndex, search the same graph on the GPU across the efSearch sweep."""
    from nmslib_zig_b200 import synth
    data = synth.gist_like(n, dim, 5, clusters=32)
    q = synth.gist_like(300, dim, 6, clusters=32)
    if space == "negdotprod":
        data, q = synth.embedding_like(n, dim, 9), synth.embedding_like(300, dim, 10)
    ref = O.RefIndex(space, "hnsw").add(data).build(params + ",indexThreadQty=8")
    path = tmp_path / "live.hnsw"
    ref.save(path)
    exact_ids, _, _ = O.seq_knn(space, data, q, 10)
    idx = nb.Index(space, None, "hnsw")
    idx.importHnsw(path)
    for ef in (50, 100, 200, 400, 1000):
        ref.set_query_params(f"efSearch={ef}")
        ri, rd, rc = ref.knn(q, 10, threads=8)
        idx.setQueryTimeParams(nb.Params({"efSearch": ef}))
        r = idx.knnQueryBatch(q, 10)
        rec_gpu, rec_ref = recall(r.ids, exact_ids), recall(ri, exact_ids)
        assert rec_gpu >= rec_ref - 2e-3, f"{space} ef={ef}: recall {rec_gpu} < reference {rec_ref}"
        assert agreement(r.ids, ri) >= 0.99, f"{space} ef={ef}: agreement {agreement(r.ids, ri)}"
    idx.deinit()
