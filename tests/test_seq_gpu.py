"""Parity of the CUDA sequential-search path (through the C ABI, host buffers in / host buffers
out) against the golden fixtures produced by the reference and against the oracle on seeded
inputs.  Integer space: bit-exact.  Float spaces: ids exact except within ties of 1e-5 relative
(+1e-6 absolute) distance -- helpers.RTOL / ATOL."""
import glob
from pathlib import Path

import numpy as np
import pytest

import nmslib_zig_b200 as nb
from helpers import ATOL, ATOL_COSINE, RTOL, assert_knn_matches, comparable
from nmslib_zig_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
SEQ_CASES = sorted(Path(p).stem for p in glob.glob(str(GOLDEN / "seq_*.npz")))


def make_index(space, data, ids=None, method="seq_search"):
    u8 = space == "l2sqr_sift"
    idx = nb.Index(space, None, method, "DenseUInt8Vector" if u8 else "DenseVector", "Int" if u8 else "Float")
    (idx.addUInt8Batch if u8 else idx.addDenseBatch)(data, ids)
    idx.buildIndex()
    return idx


def check_against_oracle(space, data, queries, k, ids=None, what=""):
    idx = make_index(space, data, ids)
    r = idx.knnQueryBatch(queries, k)
    oi, od, oc = O.seq_knn(space, data, queries, k, ids)
    ids_arr = np.arange(len(data)) if ids is None else np.asarray(ids)
    pos_of = {int(v): i for i, v in enumerate(ids_arr)}
    dist_of = lambda q, i: comparable(space, O.pair_distance(space, data[pos_of[i]], queries[q]))
    assert_knn_matches(r.ids, comparable(space, r.distances), r.sizes, oi, comparable(space, od), oc,
                       exact=(space == "l2sqr_sift"), dist_of=dist_of,
                       what=what or space, atol=ATOL_COSINE if space.startswith(("cos", "angular")) else ATOL)
    idx.deinit()
    return r


@pytest.mark.parametrize("case", SEQ_CASES)
def test_matches_reference_golden(case):
    g = np.load(GOLDEN / f"{case}.npz")
    space, k = str(g["space"]), int(g["k"])
    idx = make_index(space, g["data"], g["ids"])
    r = idx.knnQueryBatch(g["queries"], k)
    ref_d = g["ref_dists"]
    if space == "l2sqr":  # golden distances come from the reference's l2
        ref_d = (ref_d.astype(np.float64) ** 2).astype(np.float32)
    pos_of = {int(v): i for i, v in enumerate(g["ids"])}
    dist_of = lambda q, i: comparable(space, O.pair_distance(space, g["data"][pos_of[i]], g["queries"][q]))
    assert_knn_matches(r.ids, comparable(space, r.distances), r.sizes, g["ref_ids"], comparable(space, ref_d),
                       g["ref_counts"], exact=(space == "l2sqr_sift"), dist_of=dist_of, what=case,
                       atol=ATOL_COSINE if space.startswith(("cos", "angular")) else ATOL)
    # the single-query entry (lib.zig knnQuery -> get_size + fill) is a batch of one
    one = idx.knnQuery(g["queries"][0], k)
    assert np.array_equal(one.ids, r.ids[0, : r.sizes[0]]) and np.array_equal(one.distances, r.distances[0, : r.sizes[0]])
    idx.deinit()


def test_reference_own_test_vectors_through_seq_search():
    """lib.zig:1292-1299 assertions (ids[0] == 10, distances[0] ~ 0, k = 2)."""
    idx = make_index("l2", np.eye(4, dtype=np.float32)[:3], [10, 20, 30])
    r = idx.knnQuery(np.array([1, 0, 0, 0], np.float32), 2)
    assert len(r.ids) == 2 and r.ids[0] == 10 and abs(r.distances[0]) < 1e-4
    assert abs(r.distances[1] - np.sqrt(2.0)) < 1e-5
    idx.deinit()


@pytest.mark.parametrize("space,n,dim,nq,k", [
    ("l2", 10_000, 128, 1_000, 10),          # BASELINE config 1 at full size
    ("l2", 3_001, 19, 77, 7),                # ragged: dim % 4 != 0, n % 128 != 0, nq % 128 != 0
    ("l2sqr", 20_000, 128, 300, 10),
    ("cosinesimil", 5_000, 960, 64, 10),     # GIST-shaped dim
    ("cosinesimil", 4_000, 33, 130, 25),
    ("negdotprod", 6_000, 768, 96, 100),     # config-5 shape: k = 100
    ("negdotprod", 2_000, 5, 40, 3),
    ("l2", 9_000, 160, 40, 10),              # long rows, more live candidates than the re-rank buffer holds
    ("l2", 40_000, 128, 2_600, 10),          # several query blocks x several pieces, shared thresholds
    ("l2sqr", 70_000, 64, 700, 12),          # register top-16 list with margin 4; short rows
    ("l2sqr", 30_000, 96, 500, 40),          # k > 12: append buffer + deferred compaction
    ("l1", 9_000, 128, 200, 10),             # exact CUDA-core scan (no dot-product form)
    ("l1", 2_000, 19, 50, 7),
    ("linf", 9_000, 64, 200, 10),
    ("angulardist", 20_000, 128, 300, 10),   # cosine ranking on the tensor cores, acos in the re-rank
    ("angulardist", 3_000, 200, 64, 25),
])
def test_float_spaces_match_oracle(space, n, dim, nq, k):
    if space == "negdotprod" and dim == 768:
        data, q = synth.embedding_like(n, dim, 9), synth.embedding_like(nq, dim, 10)
    elif space == "l2sqr":
        data, q = synth.sift_like_f32(n, 3, dim), synth.sift_like_f32(nq, 4, dim)
    elif dim == 960:
        data, q = synth.gist_like(n, dim, 5), synth.gist_like(nq, dim, 6)
    else:
        data, q = synth.uniform(n, dim, 1) - 0.25, synth.uniform(nq, dim, 2) - 0.25
    ids = (np.arange(n, dtype=np.int32) * 3 + 11)
    check_against_oracle(space, data, q, k, ids, what=f"{space} {n}x{dim} q{nq} k{k}")


@pytest.mark.parametrize("n,nq,k", [(50_000, 256, 10), (1_000, 3, 1), (129, 129, 100)])
def test_sift_uint8_bit_exact(n, nq, k):
    data, q = synth.sift_like_u8(n, 7), synth.sift_like_u8(nq, 8)
    r = check_against_oracle("l2sqr_sift", data, q, k, what=f"sift {n} q{nq} k{k}")
    assert r.distances.dtype == np.float32 and np.all(r.distances == np.round(r.distances))


def test_sift_extreme_values_do_not_overflow():
    """max distance 128*255^2 = 8 323 200 < 2^24 stays exact as float."""
    data = np.zeros((300, 128), np.uint8)
    data[1::2] = 255
    q = np.stack([np.zeros(128, np.uint8), np.full(128, 255, np.uint8)])
    r = check_against_oracle("l2sqr_sift", data, q, 140)
    assert r.distances[1, -1] == 0 and r.distances[0, 0] == 0
    far = make_index("l2sqr_sift", data[1::2])               # only the all-255 rows
    assert far.knnQuery(q[0], 3).distances[0] == 128 * 255 * 255
    far.deinit()


def test_duplicates_zero_rows_and_tie_order():
    """Exact ties resolve by insertion position, like the reference does in practice:
    with position ids the answer must equal the oracle's id for id, not just up to ties."""
    d = synth.uniform(1000, 16, 21)
    d[500:520] = d[10:30]
    d[700] = 0.0
    q = np.concatenate([d[10:20], np.zeros((1, 16), np.float32)])
    for space in ("l2", "l2sqr", "negdotprod", "cosinesimil"):
        idx = make_index(space, d)
        r = idx.knnQueryBatch(q, 8)
        oi, od, oc = O.seq_knn(space, d, q, 8)
        assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, what=f"ties/{space}",
                           atol=ATOL_COSINE if space.startswith(("cos", "angular")) else ATOL)
        if space in ("l2", "l2sqr"):
            assert np.array_equal(r.ids[:10, 0], np.arange(10, 20))          # the earlier duplicate wins
            assert np.array_equal(r.ids[:10, 1], np.arange(500, 510))
        idx.deinit()
    du = synth.sift_like_u8(600, 23)
    du[300:340] = du[0:40]
    idx = make_index("l2sqr_sift", du)
    r = idx.knnQueryBatch(du[:16], 5)
    oi, od, oc = O.seq_knn("l2sqr_sift", du, du[:16], 5)
    assert np.array_equal(r.ids, oi) and np.array_equal(r.distances, od)      # bit-exact incl. tie order
    idx.deinit()


def test_k_larger_than_n_and_tiny_indexes():
    d = synth.uniform(7, 5, 31)
    q = synth.uniform(3, 5, 32)
    r = check_against_oracle("l2", d, q, 10)
    assert np.all(r.sizes == 7)
    assert np.all(r.ids[:, 7:] == -1)
    check_against_oracle("cosinesimil", d[:1], q, 4)


def test_incremental_add_reset_and_repeat_batches():
    d = synth.uniform(2_000, 24, 41)
    q = synth.uniform(50, 24, 42)
    idx = make_index("l2", d[:1000])
    r1 = idx.knnQueryBatch(q, 5)
    oi, od, oc = O.seq_knn("l2", d[:1000], q, 5)
    assert_knn_matches(r1.ids, r1.distances, r1.sizes, oi, od, oc, what="first half")
    idx.addDenseBatch(d[1000:], np.arange(1000, 2000, dtype=np.int32))       # re-upload on next query
    r2 = idx.knnQueryBatch(q, 5)
    oi, od, oc = O.seq_knn("l2", d, q, 5)
    assert_knn_matches(r2.ids, r2.distances, r2.sizes, oi, od, oc, what="both halves")
    r3 = idx.knnQueryBatch(q[:7], 5)                                         # smaller batch reuses scratch
    assert np.array_equal(r3.ids, r2.ids[:7])
    idx.reset()
    assert idx.dataQty() == 0
    idx.addDenseBatch(d[:10])
    idx.buildIndex()
    assert idx.knnQueryBatch(q[:2], 3).sizes.tolist() == [3, 3]
    idx.deinit()


def test_query_errors_like_the_reference():
    d = synth.uniform(100, 8, 51)
    idx = make_index("l2", d)
    with pytest.raises(nb.NmslibError) as e:
        idx.knnQueryBatch(synth.uniform(4, 9, 52), 3)        # length mismatch: CHECK in space_lp.cc:29 -> error 9
    assert e.value.name == "QueryExecutionFailed"
    with pytest.raises(nb.NmslibError) as e:
        idx.knnQuery(d[0], 0)
    assert e.value.name == "InvalidArgument"
    idx.deinit()
    unbuilt = nb.Index("l2", None, "seq_search")
    unbuilt.addDenseBatch(d)
    unbuilt.built = True                                     # bypass the host-side auto build (lib.zig:890)
    with pytest.raises(nb.NmslibError) as e:
        unbuilt.knnQueryBatch(d[:2], 3)
    assert e.value.name == "IndexBuildFailed"                # nmslib_c.cpp:963-967
    unbuilt.deinit()


def test_stats_report_our_kernels_ran():
    d = synth.uniform(3000, 32, 61)
    idx = make_index("l2", d)
    idx.knnQueryBatch(d[:100], 5)
    s = idx.stats()
    assert s["kernel_launches"] >= 2 and s["queries"] == 100 and s["distance_evals"] == 100 * 3000
    assert s["last_kernel_ms"] > 0 and s["last_total_ms"] >= s["last_kernel_ms"] and s["device_bytes"] > 0
    idx.deinit()


def test_config2_full_size_properties():
    """BASELINE config 2 at full size (1 M x 128, 10 K queries, k = 10, l2sqr): size-independent
    properties + an oracle check on a query sample."""
    data, q = synth.make("c2")
    q[:64] = data[1000:1064]                                  # planted exact matches
    idx = make_index("l2sqr", data)
    r = idx.knnQueryBatch(q, 10)
    assert np.all(r.sizes == 10)
    assert np.all(np.diff(r.distances, axis=1) >= 0)          # sortedness
    assert np.all(r.distances[:64, 0] == 0) and np.array_equal(r.ids[:64, 0], np.arange(1000, 1064))
    assert np.all(r.distances == np.round(r.distances))       # integer-valued inputs -> integer distances
    for row in r.ids[::97]:
        assert len(set(row.tolist())) == 10                   # no id twice
    sample = np.arange(0, 10_000, 313)
    oi, od, oc = O.seq_knn("l2sqr", data, q[sample], 10)
    assert_knn_matches(r.ids[sample], r.distances[sample], r.sizes[sample], oi, od, oc, what="c2 sample")
    perm = np.random.default_rng(0).permutation(2048)         # permutation invariance of the batch
    r2 = idx.knnQueryBatch(q[:2048][perm], 10)
    assert np.array_equal(r2.ids, r.ids[:2048][perm]) and np.array_equal(r2.distances, r.distances[:2048][perm])
    idx.deinit()


@pytest.mark.parametrize("space,n,dim,nq,k", [("l2sqr", 60_000, 256, 20_000, 10), ("negdotprod", 40_000, 768, 20_000, 100),
                                              ("cosinesimil", 30_000, 960, 19_200, 10)])
def test_long_rows_many_query_blocks_pair_kernel(space, n, dim, nq, k):
    """Rows of more than 128 floats run on CTA pairs (tc_scan_pair_kernel, cta_group::2).  >= 74 query blocks make
    the piece table use whole waves of blocks AND the aligned-segments + left-over tail, and the re-rank is launched
    once per list count.  Size-independent properties + the oracle on a query sample spread over all blocks."""
    if space == "negdotprod":
        data, q = synth.embedding_like(n, dim, 9), synth.embedding_like(nq, dim, 10)
    elif space == "cosinesimil":
        data, q = synth.gist_like(n, dim, 5), synth.gist_like(nq, dim, 6)
    else:
        data, q = synth.sift_like_f32(n, 3, dim), synth.sift_like_f32(nq, 4, dim)
    planted = np.arange(0, nq, 997)
    q[planted] = data[(planted * 7) % n]
    idx = make_index(space, data)
    r = idx.knnQueryBatch(q, k)
    assert np.all(r.sizes == k)
    assert np.all(np.diff(r.distances, axis=1) >= 0)
    if space != "negdotprod":                                   # a row is its own nearest neighbour (distance ~ 0)
        hit = [(planted[i] * 7) % n in r.ids[planted[i], :3] for i in range(len(planted))]
        assert np.mean(hit) == 1.0
    for row in r.ids[::211]:
        assert len(set(row.tolist())) == k
    sample = np.concatenate([np.arange(0, nq, 389), np.arange(nq - 40, nq)])   # every block range incl. the tail blocks
    oi, od, oc = O.seq_knn(space, data, q[sample], k)
    dist_of = lambda qi, i: O.pair_distance(space, data[i], q[sample[qi]])
    assert_knn_matches(r.ids[sample], r.distances[sample], r.sizes[sample], oi, od, oc, dist_of=dist_of,
                       atol=ATOL_COSINE if space == "cosinesimil" else ATOL, what=f"pair kernel {space}")
    perm = np.random.default_rng(1).permutation(nq)[:4096]      # the same queries in another batch shape
    r2 = idx.knnQueryBatch(q[perm], k)
    same = np.mean(r2.ids == r.ids[perm])
    # (cosine on clustered 960-D rows: neighbours closer than fp32 resolves swap places when a query is certified in
    # one batch and re-run by the exact kernel -- another summation order -- in the other; both pass the oracle check)
    assert same >= (0.995 if space == "cosinesimil" else 0.9999), f"batch-shape dependence: {same}"
    assert np.allclose(r2.distances, r.distances[perm], rtol=RTOL, atol=ATOL_COSINE)
    st = idx.stats()
    # concentrated 960-D clusters sit inside the error band of truncated TF32 operands: the first batch fails its
    # certificates, the engine switches to split (3xTF32) operands and certifies (nearly) all of them
    assert st["fallback_queries"] <= 0.02 * (nq + 4096), f"{st['fallback_queries']} queries re-run by the exact scan"
    if space == "cosinesimil":
        assert st["split_queries"] >= nq // 2
    idx.deinit()


def test_tensor_core_path_is_the_one_that_runs_and_certifies():
    """Float seq_search goes through the tcgen05 candidate pass + exact re-rank; the certificate
    must hold for (nearly) every query on ordinary data, i.e. the exact-scan re-run stays idle."""
    for space, data, q, k in [("l2sqr", synth.sift_like_f32(60_000, 3), synth.sift_like_f32(1_500, 4), 10),
                              ("l2", synth.uniform(20_000, 128, 1), synth.uniform(700, 128, 2), 10),
                              ("negdotprod", synth.embedding_like(30_000, 96, 9), synth.embedding_like(300, 96, 10), 100)]:
        idx = make_index(space, data)
        r = idx.knnQueryBatch(q, k)
        st = idx.stats()
        oi, od, oc = O.seq_knn(space, data, q, k)
        assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, what=f"tc/{space}")
        assert st["fallback_queries"] <= 0.02 * len(q), f"{space}: {st['fallback_queries']} uncertified queries"
        assert st["scan_count"] >= 1 and st["last_scan_ms"] > 0
        idx.deinit()


def test_uncertifiable_queries_are_rerun_exactly_and_the_margin_adapts():
    """Rows that differ by less than the pass-1 error bound cannot be certified: those queries must come back
    through the exact re-run (same answers as the oracle), and the engine must widen its candidate margin."""
    rng = np.random.default_rng(11)
    base = rng.random((1, 128), dtype=np.float32)
    # 6000 near-duplicates of one point (differences ~1e-6 << the TF32 bound) + ordinary rows
    dup = base + rng.normal(0, 1e-6, (6_000, 128)).astype(np.float32)
    data = np.concatenate([dup, synth.uniform(14_000, 128, 1)]).astype(np.float32)
    rng.shuffle(data, axis=0)
    q = np.concatenate([base + rng.normal(0, 1e-6, (40, 128)).astype(np.float32), synth.uniform(300, 128, 2)])
    idx = make_index("l2", data)
    for rep in range(3):
        r = idx.knnQueryBatch(q, 10)
        oi, od, oc = O.seq_knn("l2", data, q, 10)
        dist_of = lambda qi, i: O.pair_distance("l2", data[i], q[qi])
        assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, dist_of=dist_of, what=f"near-duplicates rep {rep}")
    st = idx.stats()
    assert st["fallback_queries"] >= 40, "the near-duplicate queries cannot have been certified"
    idx.deinit()


def test_adversarial_order_every_row_is_a_new_best():
    """Rows sorted by decreasing distance to the queries: every scanned value beats the running threshold
    (worst case for the survivor path and its buffers).  Results must still be exact."""
    q = np.zeros((70, 64), np.float32)
    q[:, 0] = np.linspace(0, 1, 70)
    n = 30_000
    data = np.zeros((n, 64), np.float32)
    data[:, 1] = np.linspace(500.0, 1.0, n)       # distance to every query decreases with the position
    data[:, 2] = (np.arange(n) % 7).astype(np.float32)
    check_against_oracle("l2sqr", data, q, 10, what="decreasing distances")


@pytest.mark.parametrize("space,dim", [("l2", 48), ("l2sqr", 128), ("cosinesimil", 33), ("negdotprod", 20),
                                       ("l2sqr_sift", 128), ("l1", 40), ("linf", 24), ("angulardist", 33)])
def test_range_query_matches_a_position_ordered_scan(space, dim):
    """nmslib_range_query_fill (nmslib_c.cpp:1051-1153) over seq_search: every object with d <= radius, in
    position order, truncated to the capacity; distances as IndexTimeDistance reports them."""
    n = 6_000
    if space == "l2sqr_sift":
        data, q = synth.sift_like_u8(n, 7), synth.sift_like_u8(3, 8)
    else:
        data, q = synth.uniform(n, dim, 1) - 0.3, synth.uniform(3, dim, 2) - 0.3
    ids = (np.arange(n, dtype=np.int32) * 3 + 7)
    idx = make_index(space, data, ids)
    oi, od, oc = O.seq_knn(space, data, q, n, ids)            # all objects, ascending distance
    for qi in range(3):
        d_sorted = od[qi]
        ref_d = {int(i): float(d) for i, d in zip(oi[qi], od[qi])}
        for want in (1, 57, 300, n // 2):
            lo, hi = float(d_sorted[want - 1]), float(d_sorted[want])
            # a radius strictly between two neighbouring distances (integer space: exactly on a value)
            radius = lo if space == "l2sqr_sift" else (lo + hi) / 2
            if radius < 0 or (space != "l2sqr_sift" and not hi > lo):
                continue  # the ABI rejects negative radii (nmslib_c.cpp:1038); negdotprod has many
            for cap in (200, 25):
                r = idx.rangeQuery(q[qi], radius, capacity=cap)
                inside = np.sort(oi[qi][d_sorted <= np.float32(radius)])     # ids grow with the position
                assert np.array_equal(r.ids, inside[:cap]), f"{space} q{qi} radius {radius} cap {cap}"
                got = np.array([ref_d[int(i)] for i in r.ids], np.float32)
                assert np.allclose(comparable(space, r.distances), comparable(space, got), rtol=RTOL * 4,
                                   atol=ATOL_COSINE), f"{space}: distances differ"
    idx.deinit()


def test_range_query_errors_like_the_reference():
    data = synth.uniform(500, 16, 1)
    idx = make_index("l2", data)
    with pytest.raises(nb.NmslibError) as e:
        idx.rangeQuery(data[0], -1.0)
    assert e.value.code == 2                      # INVALID_ARGUMENT, nmslib_c.cpp:1038-1042
    with pytest.raises(nb.NmslibError) as e:
        idx.rangeQuery(data[0, :8], 1.0)
    assert e.value.code == 9                      # length mismatch -> QUERY_EXECUTION_FAILED
    idx.deinit()


def test_pair_kernel_and_single_cta_kernel_agree(tmp_path):
    """Option tc_pair=0 keeps the single-CTA long-row kernel (tc_scan_kernel) for A/B runs: on the same data both
    nominate candidates for the same exact re-rank, so ids and distances are equal."""
    import os
    import subprocess
    import sys
    script = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {str(Path(__file__).resolve().parents[1])!r})\n"
        "import nmslib_zig_b200 as nb\n"
        "from nmslib_zig_b200 import synth\n"
        "nb.set_option('tc_pair', int(sys.argv[2]))\n"
        "out = []\n"
        "for space, dim, k in (('negdotprod', 768, 100), ('l2', 200, 10)):\n"
        "    data = synth.embedding_like(30000, dim, 9) if space == 'negdotprod' else synth.uniform(30000, dim, 1)\n"
        "    q = synth.embedding_like(3000, dim, 10) if space == 'negdotprod' else synth.uniform(3000, dim, 2)\n"
        "    idx = nb.Index(space, None, 'seq_search'); idx.addDenseBatch(data); idx.buildIndex()\n"
        "    r = idx.knnQueryBatch(q, k)\n"
        "    out += [r.ids[:, :10].copy(), r.distances[:, :10].view(np.int32).copy()]\n"
        "    idx.deinit()\n"
        "np.save(sys.argv[1], np.stack(out))\n")
    outs = []
    for mode in ("1", "0"):
        path = tmp_path / f"pair{mode}.npy"
        subprocess.run([sys.executable, "-c", script, str(path), mode], check=True, timeout=600)
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1])


def test_reset_then_shorter_rows_with_the_same_padded_length():
    """the staged-query buffer keeps its allocation across nmslib_reset_index; going from 128 to 100 floats
    (both pad to 128 words) must not leave the old queries' columns [100, 128) in the padding, which feeds the exact
    re-rank, the cosine norms and the range scan."""
    for space in ("l2", "cosinesimil", "negdotprod"):
        d1, q1 = synth.uniform(3000, 128, 71) + 5.0, synth.uniform(300, 128, 72) + 5.0     # large stale values
        idx = make_index(space, d1)
        idx.knnQueryBatch(q1, 10)
        idx.reset()
        d2, q2 = synth.uniform(3000, 100, 73), synth.uniform(300, 100, 74)
        idx.addDenseBatch(d2)
        idx.buildIndex()
        r = idx.knnQueryBatch(q2, 10)
        oi, od, oc = O.seq_knn(space, d2, q2, 10)
        assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, what=f"reset/{space}",
                           atol=ATOL_COSINE if space == "cosinesimil" else None)
        if space == "l2":
            rr = idx.rangeQuery(q2[0], float(od[0, 4]) * 1.0001, capacity=64)
            assert set(oi[0, :5].tolist()) <= set(rr.ids.tolist())
        idx.deinit()


def test_two_half_shards_on_one_gpu_merge_to_the_unsharded_answer():
    """the cross-shard step (keys carrying global positions -> k-way merge, with and without id lists) on a
    box with ONE GPU: two half-shard indexes on the same device play the two ranks."""
    import torch
    dev = torch.device("cuda", 0)
    n, dim, nq, k = 20_001, 64, 257, 10
    data, q = synth.uniform(n, dim, 81), synth.uniform(nq, dim, 82)
    data[n // 2 + 3] = data[5]
    q[0] = data[5]
    ext = np.arange(n, dtype=np.int32) * 2 + 1
    d_q = torch.from_numpy(q).to(dev)
    keys = torch.empty((2, nq, k), dtype=torch.int64, device=dev)
    ids = torch.empty((2, nq, k), dtype=torch.int32, device=dev)
    dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
    shards = []
    for r, (lo, hi) in enumerate(((0, n // 2), (n // 2, n))):
        idx = nb.Index("l2", None, "seq_search")
        idx.setShard(lo)
        idx.addDenseBatch(data[lo:hi], ext[lo:hi])
        idx.buildIndex()
        idx.knnDevice(d_q.data_ptr(), nq, dim, k, ids[r].data_ptr(), dd.data_ptr(), keys[r].data_ptr(), 0)
        shards.append(idx)
    torch.cuda.synchronize(dev)
    o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    o_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    p_ids, p_d = torch.empty_like(o_ids), torch.empty_like(o_d)
    shards[0].mergeTopk(keys.data_ptr(), ids.data_ptr(), 2, nq, k, o_ids.data_ptr(), o_d.data_ptr(), 0)
    shards[0].mergeTopk(keys.data_ptr(), 0, 2, nq, k, p_ids.data_ptr(), p_d.data_ptr(), 0)
    torch.cuda.synchronize(dev)
    assert torch.equal(p_ids * 2 + 1, o_ids) and torch.equal(p_d, o_d)   # no id lists: ids = global positions
    whole = make_index("l2", data, ext)
    w = whole.knnQueryBatch(q, k)
    assert np.array_equal(o_ids.cpu().numpy(), w.ids) and np.array_equal(o_d.cpu().numpy(), w.distances)
    assert w.ids[0, 0] == 5 * 2 + 1 and w.ids[0, 1] == (n // 2 + 3) * 2 + 1      # cross-shard tie: lower position first
    oi, od, oc = O.seq_knn("l2", data, q, k, ext)
    assert_knn_matches(w.ids, w.distances, w.sizes, oi, od, oc, what="two half shards")
    for idx in shards + [whole]:
        idx.deinit()


def test_device_entry_reruns_uncertified_queries_without_the_host():
    """nmslib_b200_knn_device never synchronises: the re-rank lists the uncertified queries on the device and gather ->
    exact scan -> merge -> scatter run predicated on that count.  Near-duplicate rows (differences far below the TF32
    bound) force the path; the count reaches the statistics afterwards; answers equal the host entry's and the oracle's."""
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    base = rng.random((1, 128), dtype=np.float32)
    dup = base + rng.normal(0, 1e-6, (6_000, 128)).astype(np.float32)
    data = np.concatenate([dup, synth.uniform(14_000, 128, 1)]).astype(np.float32)
    rng.shuffle(data, axis=0)
    q = np.concatenate([base + rng.normal(0, 1e-6, (40, 128)).astype(np.float32), synth.uniform(300, 128, 2)])
    nq, k = len(q), 10
    idx = make_index("l2", data)
    d_q = torch.from_numpy(q).to(dev)
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    st0 = idx.stats()["fallback_queries"]
    for _ in range(3):
        idx.knnDevice(d_q.data_ptr(), nq, 128, k, d_ids.data_ptr(), d_d.data_ptr(), 0, 0)
    idx.prepare()
    torch.cuda.synchronize(dev)
    oi, od, oc = O.seq_knn("l2", data, q, k)
    dist_of = lambda qi, i: O.pair_distance("l2", data[i], q[qi])
    assert_knn_matches(d_ids.cpu().numpy(), d_d.cpu().numpy(), oc, oi, od, oc, dist_of=dist_of, what="device entry, near-duplicates")
    import time
    time.sleep(0.05)
    assert idx.stats()["fallback_queries"] - st0 >= 3 * 40, "the near-duplicate queries cannot have been certified"
    r = idx.knnQueryBatch(q, k)
    assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, dist_of=dist_of, what="host entry, near-duplicates")
    idx.deinit()


def test_uint8_on_the_integer_tensor_pipe_equals_the_widened_path_bit_for_bit():
    """l2sqr_sift runs on tcgen05.mma.kind::i8 by default (byte rows, norm digits folded into the MMA); option
    u8_imma=0 keeps the round-1 path (rows widened to TF32 operands).  Same ids, same int32 distances, ties included;
    rows whose norms exceed the digit block (all-255 vectors) fall back to the widened path on their own."""
    rng = np.random.default_rng(5)
    n, nq = 70_001, 700
    data, q = synth.sift_like_u8(n, 7), synth.sift_like_u8(nq, 8)
    data[1000:1010] = data[10:20]                      # exact ties across tiles
    q[:10] = data[10:20]
    q[10] = 0
    data[5] = 0
    outs = {}
    for mode in (1, 0):
        nb.set_option("u8_imma", mode)
        idx = make_index("l2sqr_sift", data)
        for k in (10, 100):
            r = idx.knnQueryBatch(q, k)
            outs[(mode, k)] = (r.ids.copy(), r.distances.copy())
        assert idx.stats()["u8_imma"] == mode
        assert idx.stats()["fallback_queries"] <= 0.02 * nq
        idx.deinit()
    nb.set_option("u8_imma", 1)
    for k in (10, 100):
        assert np.array_equal(outs[(1, k)][0], outs[(0, k)][0]) and np.array_equal(outs[(1, k)][1], outs[(0, k)][1])
        oi, od, oc = O.seq_knn("l2sqr_sift", data, q, k)
        assert_knn_matches(outs[(1, k)][0], outs[(1, k)][1], oc, oi, od, oc, exact=True, what=f"imma k={k}")
    # the extreme: rows of 255s (|x|^2 = 128 * 255^2 is more than the digit block carries) -> widened path, still exact
    ext = np.concatenate([data[:5000], np.full((8, 128), 255, np.uint8)])
    qe = np.concatenate([q[:50], np.full((2, 128), 255, np.uint8), np.zeros((1, 128), np.uint8)])
    idx = make_index("l2sqr_sift", ext)
    r = idx.knnQueryBatch(qe, 10)
    assert idx.stats()["u8_imma"] == 0
    oi, od, oc = O.seq_knn("l2sqr_sift", ext, qe, 10)
    assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, exact=True, what="255-extreme")
    assert r.distances[52].max() <= 128 * 255 ** 2
    idx.deinit()


@pytest.mark.parametrize("space,k", [("l2", 200), ("l2", 500), ("negdotprod", 1000), ("l2sqr_sift", 1000), ("cosinesimil", 300)])
def test_large_k_has_no_cap_like_the_reference_queue(space, k):
    """KNNQueue grows to any k (knnqueue.h:55-64).  k <= 256 stays on the tensor-core path (the exact re-run of
    uncertified queries keeps its lists in a global scratch array above 144); larger k runs the exact scan with
    global lists.  k > n returns n results."""
    u8 = space == "l2sqr_sift"
    n, nq = 6_000, 40
    data = synth.sift_like_u8(n, 7) if u8 else synth.uniform(n, 48, 1)
    q = synth.sift_like_u8(nq, 8) if u8 else synth.uniform(nq, 48, 2)
    idx = make_index(space, data)
    r = idx.knnQueryBatch(q, k)
    oi, od, oc = O.seq_knn(space, data, q, k)
    dist_of = lambda qi, i: O.pair_distance(space, data[i], q[qi])
    assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, exact=u8, dist_of=dist_of, what=f"{space} k={k}",
                       atol=ATOL_COSINE if space == "cosinesimil" else None)
    small = make_index(space, data[:300])
    r = small.knnQueryBatch(q[:5], k)
    assert np.all(r.sizes == min(k, 300))
    idx.deinit()
    small.deinit()


def test_cosine_error_against_float64_is_no_worse_than_the_reference():
    """the cosine tolerance (1e-5 absolute on d = 1 - nsp) is justified by measuring both implementations
    against a float64 ground truth on 960-D rows: the device's error is bounded by the reference's own fp32 error
    (+1e-6), for every reported (query, neighbour) pair in aggregate and in the mean."""
    n, nq, dim, k = 20_000, 200, 960, 10
    data, q = synth.gist_like(n, dim, 5), synth.gist_like(nq, dim, 6)
    idx = make_index("cosinesimil", data)
    r = idx.knnQueryBatch(q, k)
    idx.deinit()
    x64, q64 = data.astype(np.float64), q.astype(np.float64)
    err_gpu, err_ref = [], []
    for qi in range(nq):
        for j in range(k):
            i = int(r.ids[qi, j])
            nsp = float(np.dot(x64[i], q64[qi]) / np.sqrt(np.dot(x64[i], x64[i])) / np.sqrt(np.dot(q64[qi], q64[qi])))
            d64 = max(0.0, 1.0 - max(-1.0, min(1.0, nsp)))
            err_gpu.append(abs(float(r.distances[qi, j]) - d64))
            err_ref.append(abs(O.pair_distance("cosinesimil", data[i], q[qi]) - d64))
    err_gpu, err_ref = np.array(err_gpu), np.array(err_ref)
    assert err_gpu.max() <= err_ref.max() + 1e-6, (err_gpu.max(), err_ref.max())
    assert err_gpu.mean() <= err_ref.mean() + 2e-7, (err_gpu.mean(), err_ref.mean())
    assert err_gpu.max() <= ATOL_COSINE / 2      # the tolerance the parity tests use has a factor of two to spare


@pytest.mark.parametrize("space", ["l2", "cosinesimil", "negdotprod", "l2sqr_sift"])
def test_appended_rows_are_uploaded_alone_and_answers_equal_a_fresh_index(space):
    """adding rows to an index that already lives in HBM uploads only the new rows (buffers grow keeping
    their contents, operands of the tensor-core scan are prepared for the new rows only); the answers are those of an
    index built from all rows at once, bit for bit."""
    u8 = space == "l2sqr_sift"
    n1, n2, nq, k = 5_003, 2_222, 200, 10
    data = synth.sift_like_u8(n1 + n2, 7) if u8 else synth.uniform(n1 + n2, 40, 1) - 0.2
    q = synth.sift_like_u8(nq, 8) if u8 else synth.uniform(nq, 40, 2) - 0.2
    ids = np.arange(n1 + n2, dtype=np.int32) + 100
    idx = make_index(space, data[:n1], ids[:n1])
    idx.knnQueryBatch(q, k)
    assert idx.stats()["uploaded_rows"] == n1
    (idx.addUInt8Batch if u8 else idx.addDenseBatch)(data[n1:], ids[n1:])
    r = idx.knnQueryBatch(q, k)
    assert idx.stats()["uploaded_rows"] == n1 + n2                     # not n1 + (n1 + n2)
    fresh = make_index(space, data, ids)
    f = fresh.knnQueryBatch(q, k)
    assert np.array_equal(r.ids, f.ids) and np.array_equal(r.distances.view(np.int32), f.distances.view(np.int32))
    oi, od, oc = O.seq_knn(space, data, q, k, ids)
    assert_knn_matches(r.ids, r.distances, r.sizes, oi, od, oc, exact=u8, what=f"append/{space}",
                       atol=ATOL_COSINE if space == "cosinesimil" else None)
    idx.deinit()
    fresh.deinit()


@pytest.mark.parametrize("u8", [False, True])
def test_large_upload_through_pinned_staging_lands_every_chunk_in_place(u8):
    """Uploads of 256 MB and more travel in 4 MB chunks through pinned staging buffers, several host threads and copy
    streams (engine.cu upload_rows): every row must arrive at its own position -- rows on both sides of
    every chunk boundary are queried for themselves."""
    if u8:
        n, dim = 2_200_000, 128                      # 282 MB of byte rows (32 768 rows per chunk)
        rng = np.random.default_rng(5)
        data = rng.integers(0, 256, (n, dim), dtype=np.uint8)
        idx = nb.Index("l2sqr_sift", None, "seq_search", "DenseUInt8Vector", "Int")
        idx.addUInt8Batch(data)
        per_chunk = (4 << 20) // dim
    else:
        n, dim = 600_000, 120                        # 288 MB; rows of 480 bytes padded to 512 on the device
        data = synth.uniform(n, dim, 9)
        idx = nb.Index("l2", None, "seq_search")
        idx.addDenseBatch(data)
        per_chunk = (4 << 20) // (dim * 4)
    idx.buildIndex()
    pos = np.unique(np.concatenate([np.arange(0, n, per_chunk), np.arange(per_chunk - 1, n, per_chunk), [n - 1],
                                    np.random.default_rng(1).integers(0, n, 64)]))
    r = idx.knnQueryBatch(data[pos], 1)
    assert np.array_equal(r.ids[:, 0], pos.astype(np.int32))
    assert np.all(r.distances[:, 0] == 0)
    assert idx.stats()["uploaded_rows"] == n
    idx.deinit()
