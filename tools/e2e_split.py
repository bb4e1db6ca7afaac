"""Where the end-to-end call spends its time: python wrapper / C ABI / device (events inside knn_host)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth
import torch

n, dim, nq, k = 1_000_000, 128, 10_000, 10
data = synth.uniform(n, dim, 1)
q = torch.from_numpy(synth.uniform(nq, dim, 2)).pin_memory().numpy()
idx = nb.Index("l2sqr", None, "seq_search")
idx.addDenseBatch(data)
idx.buildIndex()
for _ in range(3):
    idx.knnQueryBatch(q, k)
ts, tot, ker = [], [], []
for _ in range(20):
    t0 = time.perf_counter()
    idx.knnQueryBatch(q, k)
    ts.append((time.perf_counter() - t0) * 1e3)
    st = idx.stats()
    tot.append(st["last_total_ms"]); ker.append(st["last_kernel_ms"])
print("python call ms  min/med", min(ts), sorted(ts)[10])
print("device total ms min/med", min(tot), sorted(tot)[10])
print("kernels ms      min/med", min(ker), sorted(ker)[10])
