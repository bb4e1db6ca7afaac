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


def test_lib_zig_dense_workflow_through_hnsw(tmp_path):
    """The reference's own flagship test, call for call (lib.zig:1273-1313): hnsw over three unit
    vectors with ids 10/20/30, k = 2 -> ids[0] == 10, distances[0] ~ 0, getDistance(0,1) ~ sqrt 2,
    save -> reset -> load -> 3 points."""
    idx = nb.Index.init("l2", nb.Params({"dim": 4}), "hnsw", "DenseVector", "Float")
    data = np.eye(4, dtype=np.float32)[:3]
    idx.addDenseBatch(data, [10, 20, 30])
    idx.buildIndex(None, False)
    assert idx.dataQty() == 3 and idx.getSpaceType() == "l2" and idx.getMethod() == "hnsw"
    r = idx.knnQuery(np.array([1, 0, 0, 0], np.float32), 2)
    assert len(r.ids) == 2 and r.ids[0] == 10 and abs(r.distances[0]) < 1e-4
    assert abs(idx.getDistance(0, 1) - np.sqrt(2.0)) < 1e-4
    assert np.array_equal(idx.getDataPoint(0), data[0])
    idx.save(str(tmp_path / "test_index"), True)
    idx.reset()
    assert idx.dataQty() == 0
    loaded = nb.Index.load(str(tmp_path / "test_index"), "DenseVector", "Float", True)
    assert loaded.dataQty() == 3 and np.array_equal(loaded.getDataPoint(0), data[0])
    r2 = loaded.knnQuery(np.array([0, 1, 0, 0], np.float32), 2)
    assert r2.ids[0] == 20
    loaded.deinit()
    idx.deinit()


def test_lib_zig_uint8_workflow_through_hnsw():
    """lib.zig:1357-1380: l2sqr_sift + hnsw over uint8 vectors (the reference searches the pointer graph,
    hnsw.cc:1174-1300; here the same beam runs on float-widened rows, distances stay exact integers)."""
    from nmslib_zig_b200 import synth
    from oracle import oracle as O
    data = synth.sift_like_u8(5000, 7)
    q = synth.sift_like_u8(64, 8)
    idx = nb.Index("l2sqr_sift", None, "hnsw", "DenseUInt8Vector", "Int")
    idx.addUInt8Batch(data)
    idx.buildIndex(nb.Params({"M": 16, "efConstruction": 100}))
    r = idx.knnQueryBatch(q, 10)
    ei, ed, _ = O.seq_knn("l2sqr_sift", data, q, 10)
    assert recall(r.ids, ei) >= 0.9
    assert np.all(r.distances == np.round(r.distances))
    for i in range(len(q)):                      # every reported distance is the exact integer distance of its id
        for j in range(10):
            assert r.distances[i, j] == O.pair_distance("l2sqr_sift", data[r.ids[i, j]], q[i])
    two = idx.knnQuery(data[3], 2)
    assert two.ids[0] == 3 and two.distances[0] == 0
    idx.deinit()


@pytest.mark.parametrize("space", ["l2", "cosinesimil", "negdotprod"])
def test_self_built_graph_recall(space):
    from nmslib_zig_b200 import synth
    from oracle import oracle as O
    data = synth.gist_like(20_000, 48, 41, clusters=16)
    q = synth.gist_like(300, 48, 42, clusters=16)
    idx = nb.Index(space, None, "hnsw")
    idx.addDenseBatch(data)
    idx.buildIndex(nb.Params({"M": 16, "efConstruction": 200}))
    ei, _, _ = O.seq_knn(space, data, q, 10)
    last = 0.0
    for ef in (20, 100, 400):
        idx.setQueryTimeParams(nb.Params({"efSearch": ef}))
        rec = recall(idx.knnQueryBatch(q, 10).ids, ei)
        assert rec >= last - 0.01
        last = rec
    assert last >= 0.95
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
    for ef in (50, 100, 200, 400, 1000, 4096):         # (>= 1000: the reference's SearchOld; 4096: above the round-1 cap)
        ref.set_query_params(f"efSearch={ef}")
        ri, rd, rc = ref.knn(q, 10, threads=8)
        idx.setQueryTimeParams(nb.Params({"efSearch": ef}))
        r = idx.knnQueryBatch(q, 10)
        rec_gpu, rec_ref = recall(r.ids, exact_ids), recall(ri, exact_ids)
        assert rec_gpu >= rec_ref - 2e-3, f"{space} ef={ef}: recall {rec_gpu} < reference {rec_ref}"
        assert agreement(r.ids, ri) >= 0.99, f"{space} ef={ef}: agreement {agreement(r.ids, ri)}"
    idx.deinit()


def test_team_kernel_answers_are_bit_identical_to_the_one_warp_kernel(tmp_path):
    """Small batches run a team of 2 or 4 warps per query (hnsw_search_team_kernel): same expansions, same arithmetic
    per pair, so the same keys.  The switch is read once per process, hence one child process per kernel."""
    import subprocess
    import sys
    script = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {str(Path(__file__).resolve().parents[1])!r})\n"
        "import nmslib_zig_b200 as nb\n"
        "from nmslib_zig_b200 import synth\n"
        "nb.set_option('hnsw_team', int(sys.argv[2]))\n"
        "data, q = synth.gist_like(20000, 96, 5, clusters=16), synth.gist_like(700, 96, 6, clusters=16)\n"
        "idx = nb.Index('cosinesimil', None, 'hnsw'); idx.addDenseBatch(data)\n"
        "idx.buildIndex(nb.Params({'M': 16, 'efConstruction': 100, 'b200_build': 'host', 'indexThreadQty': 1}))\n"
        "out = []\n"
        "for ef in (10, 64, 300, 1000):\n"
        "    idx.setQueryTimeParams(nb.Params({'efSearch': ef}))\n"
        "    r = idx.knnQueryBatch(q, 10); out += [r.ids, r.distances.view(np.int32)]\n"
        "np.save(sys.argv[1], np.stack(out))\n")
    outs = []
    for mode in ("0", "2", "4"):
        path = tmp_path / f"team{mode}.npy"
        subprocess.run([sys.executable, "-c", script, str(path), mode], check=True, timeout=600)
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_int_space_hnsw_same_graph_parity_with_the_reference():
    """l2sqr_sift + hnsw is Hnsw<int> in the reference -- the pointer graph, saved by SaveRegularIndexBin
    (hnsw.cc:810-842) and searched by baseSearchAlgorithmV1Merge / Old (hnsw.cc:1076-1300).  The golden file is that
    index as the reference wrote it; the kernel must return the reference's ids on it (ties aside) with the exact
    int32 distances, at every efSearch incl. the SearchOld regime."""
    g = np.load(GOLDEN / "regular_hnsw_sift.npz")
    data, ids, q = g["data"], g["ids"], g["queries"]
    idx = nb.Index("l2sqr_sift", None, "hnsw", "DenseUInt8Vector", "Int")
    idx.addUInt8Batch(data, ids)
    idx.importHnsw(GOLDEN / "regular_hnsw_sift.hnsw")
    exact = g["exact_ids"]
    pos_of = {int(e): i for i, e in enumerate(ids)}
    for ef in g["efs"]:
        idx.setQueryTimeParams(nb.Params({"efSearch": int(ef)}))
        r = idx.knnQueryBatch(q, 10)
        ri, rd = g[f"ids_ef{ef}"], g[f"dists_ef{ef}"]
        assert recall(r.ids, exact) >= recall(ri, exact) - 1e-9, f"ef={ef}"
        assert agreement(r.ids, ri) >= 0.99, f"ef={ef}: agreement {agreement(r.ids, ri)}"
        assert np.array_equal(np.sort(r.distances, axis=1), r.distances)
        for qi in range(0, len(q), 7):                       # reported distances are the exact integer distances
            for j in range(10):
                x = data[pos_of[int(r.ids[qi, j])]].astype(np.int64)
                assert int(r.distances[qi, j]) == int(np.sum((x - q[qi].astype(np.int64)) ** 2))
        same = r.ids == ri
        assert np.array_equal(r.distances[same], rd[same])
    idx.deinit()
    # a regular index needs its data set first (LoadRegularIndexBin CHECKs the counts, hnsw.cc:956-959)
    empty = nb.Index("l2sqr_sift", None, "hnsw", "DenseUInt8Vector", "Int")
    empty.addUInt8Batch(data[:10])
    with pytest.raises(nb.NmslibError):
        empty.importHnsw(GOLDEN / "regular_hnsw_sift.hnsw")
    empty.deinit()
