"""Timing of the exact CUDA-core scan (K2): spaces without a dot-product form (l1, linf) and force_exact runs.
usage: exact_prof.py [space] [n] [nq] [k] [dim]"""
import sys, time
sys.path.insert(0, ".")
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth

space = sys.argv[1] if len(sys.argv) > 1 else "l1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dim = int(sys.argv[5]) if len(sys.argv) > 5 else 128
if space in ("l2_exact", "cos_exact"):
    nb.set_option("force_exact", 1)
    space = {"l2_exact": "l2", "cos_exact": "cosinesimil"}[space]
data, q = synth.sift_like_f32(n, 3, dim), synth.sift_like_f32(nq, 4, dim)
idx = nb.Index(space, None, "seq_search")
idx.addDenseBatch(data)
idx.buildIndex()
best = 1e9
for _ in range(4):
    idx.knnQueryBatch(q, k)
    best = min(best, idx.stats()["last_kernel_ms"])
print(f"{space} n={n} dim={dim} nq={nq} k={k}: kernels {best:.2f} ms = {2.0 * n * nq * dim / best / 1e9:.1f} TFLOP/s-equivalent (2*Q*N*D)", flush=True)
