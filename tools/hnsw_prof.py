"""One HNSW graph (built by the reference, test infrastructure) and three query batches at one efSearch:
the ncu target for hnsw_search_kernel (`-k regex:hnsw_search -s 1 -c 1`).
usage: hnsw_prof.py [n] [dim] [ef] [nq]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 960
ef = int(sys.argv[3]) if len(sys.argv) > 3 else 200
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 9_472
data, q = synth.gist_like(n, dim, 5), synth.gist_like(nq, dim, 6)
path = f"/tmp/nb200_prof_{n}_{dim}.hnsw"
if not os.path.exists(path):
    ref = O.RefIndex("cosinesimil", "hnsw").add(data).build(f"M=16,efConstruction=200,indexThreadQty={os.cpu_count()}")
    ref.save(path)
idx = nb.Index("cosinesimil", None, "hnsw")
idx.importHnsw(path)
idx.setQueryTimeParams(nb.Params({"efSearch": ef}))
import time
try:
    import torch
    q = torch.from_numpy(q).pin_memory().numpy()  # pinned host queries, as bench.py's e2e leg uses
except Exception:
    pass
for _ in range(3):
    s0 = idx.stats()
    t0 = time.perf_counter()
    idx.knnQueryBatch(q, 10)
    dt = time.perf_counter() - t0
    s = idx.stats()
    ms = s["scan_ms_sum"] - s0["scan_ms_sum"]
    print("kernel_ms", ms, "q/s", nq / (ms * 1e-3), "e2e_ms", dt * 1e3, "e2e q/s", nq / dt, flush=True)
idx.deinit()
