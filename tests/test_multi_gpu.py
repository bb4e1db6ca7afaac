"""Row-sharded multi-GPU query path on real GPUs (needs >= 2 devices; skipped otherwise):
per-rank CUDA scan -> NCCL all-gather of (key, id) lists -> device k-way merge must equal the
oracle's unsharded answer, and equal the single-GPU answer bit for bit."""
import os
import socket
import sys
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
