"""world_size-2 (and 3) gloo tests of the multi-rank path on CPU: row sharding, global-position
keys, the all-gather layout and the k-way merge specification reproduce the unsharded answer
exactly.  The per-shard scan is played by the oracle here; on GPUs it is the CUDA
engine (bench.py --gpus N, tests/test_multi_gpu.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, space, n, dim, nq, k, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nmslib_zig_b200 import shard, synth
    from oracle import oracle as O
    u8 = space == "l2sqr_sift"
    data = synth.sift_like_u8(n, 7) if u8 else synth.uniform(n, dim, 1)
    data[n // 2 + 3] = data[5]                       # a duplicate across the shard boundary: tie by position
    q = (synth.sift_like_u8(nq, 8) if u8 else synth.uniform(nq, dim, 2))
    q[0] = data[5]
    lo, hi = shard.shard_bounds(n, rank, world)
    ext = (np.arange(n, dtype=np.int32) * 2 + 1)
    ids, d, c = O.seq_knn(space, data[lo:hi], q, k, ext[lo:hi], threads=2)
    # local positions -> global positions, exactly what Index.setShard(lo) makes the engine emit
    pos = np.full((nq, k), -1, np.int64)
    for i in range(nq):
        pos[i, : c[i]] = (ids[i, : c[i]].astype(np.int64) - 1) // 2
    keys = shard.make_keys(d, pos)
    g_keys = [torch.empty((nq, k), dtype=torch.int64) for _ in range(world)]
    g_ids = [torch.empty((nq, k), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(g_keys, torch.from_numpy(keys.view(np.int64)))
    dist.all_gather(g_ids, torch.from_numpy(ids))
    if rank == 0:
        gk = torch.stack(g_keys).numpy().view(np.uint64)
        gi = torch.stack(g_ids).numpy()
        m_ids, m_keys = shard.merge_gathered_host(gk, gi, k)
        full_ids, full_d, full_c = O.seq_knn(space, data, q, k, ext, threads=2)
        np.savez(out_path, m_ids=m_ids, full_ids=full_ids, m_keys=m_keys,
                 full_keys=shard.make_keys(full_d, (full_ids.astype(np.int64) - 1) // 2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,space,n,k", [(2, "l2", 2001, 10), (3, "negdotprod", 1500, 7), (2, "l2sqr_sift", 1203, 12)])
def test_sharded_scan_allgather_merge_equals_unsharded(tmp_path, world, space, n, k):
    out = tmp_path / "res.npz"
    mp.spawn(_worker, args=(world, _free_port(), space, n, 24, 33, k, str(out)), nprocs=world, join=True)
    r = np.load(out)
    assert np.array_equal(r["m_ids"], r["full_ids"])       # identical, including the cross-shard tie order
    assert np.array_equal(r["m_keys"], r["full_keys"])


def test_shard_bounds_cover_every_row_once():
    from nmslib_zig_b200 import shard
    for n in (1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_key_order_is_distance_then_position():
    from nmslib_zig_b200 import shard
    d = np.array([[0.0, -0.0, 1.5, -2.0, np.inf, 1.5]], np.float32)
    p = np.array([[3, 1, 9, 4, 2, 7]], np.int64)
    keys = shard.make_keys(d, p)[0]
    order = np.argsort(keys, kind="stable")
    assert order.tolist() == [3, 1, 0, 5, 2, 4]            # -2 < (+-0: pos 1 < pos 3) < 1.5 (pos 7 < 9) < inf


def test_shard_generation_matches_the_whole_data_set():
    """bench.py lets every rank generate only its own rows (config 5: 3.8 GB instead of 30.7 GB per rank): a shard made
    with rows=(lo, hi) must be byte-identical to the same rows of the whole data set, across chunk boundaries too."""
    from nmslib_zig_b200 import synth
    from nmslib_zig_b200.shard import shard_bounds
    n, dim = 300_000, 8                       # chunks of 131 072 rows: three chunks, shards cut through them
    whole = synth.embedding_like(n, dim, 9)
    parts = [synth.embedding_like(n, dim, 9, rows=shard_bounds(n, r, 3)) for r in range(3)]
    assert np.array_equal(np.concatenate(parts), whole)
    d2, _ = synth.make("c2", 5000, 8, rows=(1000, 2500))
    full2, _ = synth.make("c2", 5000, 8)
    assert np.array_equal(d2, full2[1000:2500])
