"""One process, several devices behind the plain C ABI (index parameter b200_devices; include/nmslib_b200.h mode A):
end-to-end queries/s of nmslib_knn_query_batch with host buffers, against the same index on one device.
usage: group_bench.py <devices e.g. 0-3> [c2|c3] [n]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth
import torch

devices = sys.argv[1] if len(sys.argv) > 1 else "0,1"
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
space, method, dtype, dist, n0, dim, nq, k, _, _ = synth.CONFIGS[cfg]
n = int(sys.argv[3]) if len(sys.argv) > 3 else n0
data, q = synth.make(cfg, n, nq)
q = torch.from_numpy(q).pin_memory().numpy()


def run(dev_param):
    idx = nb.Index(space, None, method, dtype, dist)
    idx.addDenseBatch(data)
    params = {"b200_devices": dev_param} if dev_param else {}
    if method == "hnsw":
        params.update({"M": 16, "efConstruction": 200})
    t0 = time.perf_counter()
    idx.buildIndex(nb.Params(params))
    if method == "hnsw":
        idx.setQueryTimeParams(nb.Params({"efSearch": 400}))
    r = idx.knnQueryBatch(q, k)
    t_first = time.perf_counter() - t0
    for _ in range(2):
        idx.knnQueryBatch(q, k)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        r = idx.knnQueryBatch(q, k)
        ts.append(time.perf_counter() - t0)
    idx.deinit()
    return r, min(ts), t_first


one, t1, f1 = run(None)
grp, tg, fg = run(devices)
same = np.array_equal(one.ids, grp.ids) and np.array_equal(one.distances.view(np.int32), grp.distances.view(np.int32))
print(f"{cfg} n={n} nq={nq} k={k}: one device {nq / t1:,.0f} q/s ({t1 * 1e3:.3f} ms; build + first batch {f1:.2f} s) | "
      f"b200_devices={devices} {nq / tg:,.0f} q/s ({tg * 1e3:.3f} ms; build + shards + first batch {fg:.2f} s) | "
      f"speed-up {t1 / tg:.2f} | identical answers: {same}", flush=True)
