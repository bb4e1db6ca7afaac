"""Row-sharded multi-GPU query path on real GPUs (needs >= 2 devices; skipped otherwise):
per-rank CUDA scan -> NCCL all-gather of (key, id) lists -> device k-way merge must equal the
oracle's unsharded answer, and equal the single-GPU answer bit for bit."""
import os
import socket
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, space, n, dim, nq, k, out_path):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import nmslib_zig_b200 as nb
    from nmslib_zig_b200 import shard, synth
    nb.set_device(rank)
    u8 = space == "l2sqr_sift"
    data = synth.sift_like_u8(n, 7) if u8 else synth.uniform(n, dim, 1)
    data[n // 2 + 3] = data[5]
    q = synth.sift_like_u8(nq, 8) if u8 else synth.uniform(nq, dim, 2)
    q[0] = data[5]
    lo, hi = shard.shard_bounds(n, rank, world)
    idx = nb.Index(space, None, "seq_search", "DenseUInt8Vector" if u8 else "DenseVector", "Int" if u8 else "Float")
    idx.setShard(lo)
    (idx.addUInt8Batch if u8 else idx.addDenseBatch)(data[lo:hi], np.arange(lo, hi, dtype=np.int32) * 2 + 1)
    idx.buildIndex()
    dev = torch.device("cuda", rank)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    d_q = torch.from_numpy(q).to(dev)
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    idx.knnDevice(d_q.data_ptr(), nq, q.shape[1], k, d_ids.data_ptr(), d_d.data_ptr(), d_keys.data_ptr(),
                  stream.cuda_stream)
    g_keys = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    g_ids = torch.empty((world, nq, k), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(g_keys, d_keys)
    dist.all_gather_into_tensor(g_ids, d_ids)
    o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    o_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx.mergeTopk(g_keys.data_ptr(), g_ids.data_ptr(), world, nq, k, o_ids.data_ptr(), o_d.data_ptr(),
                  stream.cuda_stream)
    # without id lists the merge reports the global POSITION in the key: ids here are 2 * position + 1
    p_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    p_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    idx.mergeTopk(g_keys.data_ptr(), 0, world, nq, k, p_ids.data_ptr(), p_d.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize(dev)
    assert torch.equal(p_ids * 2 + 1, o_ids) and torch.equal(p_d, o_d)
    # the library's own exchange (include/nmslib_b200.h mode B): windows exported, blobs all-gathered over the host
    # channel, then ONE call per rank returns the global answer -- on the device API and through nmslib_knn_query_batch
    blobs = [None] * world
    dist.all_gather_object(blobs, idx.shardExport(nq, k))
    idx.shardConnect(rank, world, b"".join(blobs))
    x_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    x_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(3):                                            # (several steps: the double-buffered windows turn over)
        idx.knnDevice(d_q.data_ptr(), nq, q.shape[1], k, x_ids.data_ptr(), x_d.data_ptr(), 0, stream.cuda_stream)
    torch.cuda.synchronize(dev)
    assert torch.equal(x_ids, o_ids) and torch.equal(x_d, o_d), "peer-memory exchange != NCCL all-gather + merge"
    r = idx.knnQueryBatch(q, k)
    assert np.array_equal(r.ids, o_ids.cpu().numpy()) and np.array_equal(r.distances, o_d.cpu().numpy())
    dist.barrier()
    idx.shardDisconnect()
    if rank == 0:
        np.savez(out_path, ids=o_ids.cpu().numpy(), dists=o_d.cpu().numpy())
    idx.deinit()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("space,n,dim,k", [("l2", 40_001, 128, 10), ("negdotprod", 20_000, 96, 100),
                                           ("l2sqr_sift", 30_000, 128, 10)])
def test_two_gpu_sharded_equals_oracle(tmp_path, space, n, dim, k):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from helpers import assert_knn_matches
    from nmslib_zig_b200 import synth
    from oracle import oracle as O
    nq = 300
    out = tmp_path / "res.npz"
    mp.spawn(_worker, args=(2, _free_port(), space, n, dim, nq, k, str(out)), nprocs=2, join=True)
    r = np.load(out)
    u8 = space == "l2sqr_sift"
    data = synth.sift_like_u8(n, 7) if u8 else synth.uniform(n, dim, 1)
    data[n // 2 + 3] = data[5]
    q = synth.sift_like_u8(nq, 8) if u8 else synth.uniform(nq, dim, 2)
    q[0] = data[5]
    oi, od, oc = O.seq_knn(space, data, q, k, np.arange(n, dtype=np.int32) * 2 + 1)
    assert_knn_matches(r["ids"], r["dists"], oc, oi, od, oc, exact=u8, what=f"2gpu/{space}")
    assert r["ids"][0, 0] == 5 * 2 + 1 and r["ids"][0, 1] == (n // 2 + 3) * 2 + 1   # cross-shard tie: lower position first


def _group_matches_single(space, n, dim, nq, k, devices):
    import nmslib_zig_b200 as nb
    from nmslib_zig_b200 import synth
    u8 = space == "l2sqr_sift"
    data = synth.sift_like_u8(n, 7) if u8 else synth.uniform(n, dim, 1)
    data[n // 2 + 3] = data[5]
    q = synth.sift_like_u8(nq, 8) if u8 else synth.uniform(nq, dim, 2)
    q[0] = data[5]
    ext = np.arange(n, dtype=np.int32) * 3 + 7
    args = (space, None, "seq_search", "DenseUInt8Vector" if u8 else "DenseVector", "Int" if u8 else "Float")
    one = nb.Index(*args)
    (one.addUInt8Batch if u8 else one.addDenseBatch)(data, ext)
    one.buildIndex()
    grp = nb.Index(*args)
    (grp.addUInt8Batch if u8 else grp.addDenseBatch)(data, ext)
    grp.buildIndex(nb.Params({"b200_devices": devices}))
    for batch in (q, q[:37], q):                          # batch shapes change: windows and slices follow
        a, b = one.knnQueryBatch(batch, k), grp.knnQueryBatch(batch, k)
        assert np.array_equal(a.ids, b.ids) and np.array_equal(a.distances.view(np.int32), b.distances.view(np.int32))
        assert np.array_equal(a.sizes, b.sizes)
    assert b.ids[0, 0] == 5 * 3 + 7 and b.ids[0, 1] == (n // 2 + 3) * 3 + 7     # cross-shard tie: lower position first
    single = grp.knnQuery(q[1], k)                        # knn_query_fill = a batch of one
    assert np.array_equal(single.ids, a.ids[1][:len(single.ids)])
    grp.addUInt8Batch(data[:100], ext[:100] + 1) if u8 else grp.addDenseBatch(data[:100], ext[:100] + 1)
    one.addUInt8Batch(data[:100], ext[:100] + 1) if u8 else one.addDenseBatch(data[:100], ext[:100] + 1)
    a, b = one.knnQueryBatch(q, k), grp.knnQueryBatch(q, k)                 # rows added: shards are cut again
    assert np.array_equal(a.ids, b.ids) and np.array_equal(a.distances.view(np.int32), b.distances.view(np.int32))
    assert grp.stats()["kernel_launches"] > 0
    one.deinit()
    grp.deinit()


@pytest.mark.parametrize("space,n,dim,k", [("l2", 20_001, 128, 10), ("cosinesimil", 9_000, 200, 10),
                                           ("negdotprod", 12_000, 96, 100), ("l2sqr_sift", 30_000, 128, 10)])
def test_shard_group_two_shards_on_one_gpu_equals_the_single_index(space, n, dim, k):
    """include/nmslib_b200.h mode (A) through the unmodified query entry points, C ABI only (no torch): an index built
    with b200_devices=0,0 (two row shards on ONE GPU -- legal because the group orders publish -> merge with events,
    not with in-kernel waits) answers bit-identically to the plain index."""
    _group_matches_single(space, n, dim, 257, k, "0,0")


def test_shard_group_tiny_index_and_bad_device_lists():
    import nmslib_zig_b200 as nb
    idx = nb.Index("l2", None, "seq_search")
    idx.addDenseBatch(np.eye(4, dtype=np.float32)[:3], [10, 20, 30])          # 3 rows, 2 shards: served by one device
    idx.buildIndex(nb.Params({"b200_devices": "0,0"}))
    r = idx.knnQuery(np.array([1, 0, 0, 0], np.float32), 2)
    assert r.ids[0] == 10 and abs(r.distances[0]) < 1e-6
    idx.deinit()
    for bad in ("0,,1", "a", "3-1", "999"):
        idx = nb.Index("l2", None, "seq_search")
        idx.addDenseBatch(np.eye(4, dtype=np.float32))
        with pytest.raises(nb.NmslibError):
            idx.buildIndex(nb.Params({"b200_devices": bad}))
        idx.deinit()


def test_shard_group_on_two_gpus_equals_the_single_index():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _group_matches_single("l2", 50_001, 128, 1000, 10, "0,1")
    _group_matches_single("l2sqr_sift", 40_000, 128, 300, 10, "0-1")


def _hnsw_group_matches_single(space, devices):
    """method hnsw + b200_devices: every device holds a replica of the graph and takes a slice of the batch; the answers
    are those of the one-device index (same graph: the group's host index builds it once, the plain index imports it)."""
    import nmslib_zig_b200 as nb
    from nmslib_zig_b200 import synth
    n, dim, nq, k = 6000, 48, 301, 10
    data, q = synth.uniform(n, dim, 1), synth.uniform(nq, dim, 2)
    ext = np.arange(n, dtype=np.int32) * 2 + 11
    grp = nb.Index(space, None, "hnsw")
    grp.addDenseBatch(data, ext)
    grp.buildIndex(nb.Params({"M": 12, "efConstruction": 80, "b200_build": "host", "b200_devices": devices}))
    grp.setQueryTimeParams(nb.Params({"efSearch": 64}))
    b = grp.knnQueryBatch(q, k)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "g.hnsw")
        grp.save(path)                                     # the graph the replicas search
        one = nb.Index.load(path)
    one.setQueryTimeParams(nb.Params({"efSearch": 64}))
    a = one.knnQueryBatch(q, k)
    assert np.array_equal(a.ids, b.ids) and np.array_equal(a.distances.view(np.int32), b.distances.view(np.int32))
    assert np.array_equal(a.sizes, b.sizes)
    grp.setQueryTimeParams(nb.Params({"efSearch": 200}))   # reaches every replica
    one.setQueryTimeParams(nb.Params({"efSearch": 200}))
    a, b = one.knnQueryBatch(q[:50], k), grp.knnQueryBatch(q[:50], k)
    assert np.array_equal(a.ids, b.ids)
    single = grp.knnQuery(q[1], k)
    assert np.array_equal(single.ids, a.ids[1][:len(single.ids)])
    st = grp.stats()
    assert st["kernel_launches"] > 0 and st["queries"] >= nq
    one.deinit()
    grp.deinit()


@pytest.mark.gpu
@pytest.mark.parametrize("space", ["l2", "cosinesimil"])
def test_hnsw_replicas_on_one_device_match_plain_index(space):
    _hnsw_group_matches_single(space, "0,0")


@pytest.mark.gpu
def test_hnsw_replicas_on_two_devices_match_plain_index():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _hnsw_group_matches_single("l2", "0,1")
