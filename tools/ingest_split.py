"""Ingest split: host slab copy (add) / first upload (one-time costs inside) / second upload of the same
rows after a reset (buffers and kernels warm).  usage: ingest_split.py [n] [dim] [u8]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")   # the CUDA context exists before the clock starts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
u8 = len(sys.argv) > 3 and sys.argv[3] == "u8"
data = synth.sift_like_u8(n, 7) if u8 else synth.uniform(n, dim, 1)
gb = data.nbytes / 1e9
idx = nb.Index("l2sqr_sift", None, "seq_search", "DenseUInt8Vector", "Int") if u8 else nb.Index("l2sqr", None, "seq_search")
L = nb.index.lib()
for rnd in range(3):
    t0 = time.perf_counter()
    (idx.addUInt8Batch if u8 else idx.addDenseBatch)(data)
    t1 = time.perf_counter()
    idx.buildIndex()
    L.nmslib_initialize_pool(idx.handle)
    t2 = time.perf_counter()
    print(f"round {rnd}: add {t1 - t0:.3f} s ({gb / (t1 - t0):.1f} GB/s)  upload+prep {t2 - t1:.3f} s ({gb / (t2 - t1):.1f} GB/s)  "
          f"rows/s {n / (t2 - t0):.3g}", flush=True)
    idx.reset() if hasattr(idx, "reset") else L.nmslib_reset_index(idx.handle)
