"""Three config-2 query batches (the ncu target: `-k regex:tc_scan -s 1 -c 1` captures a warm launch).
usage: tc_prof.py [dim] [n] [nq] [k] [space]   (space negdotprod: the embedding-shaped rows of config 5; l2sqr_sift: uint8 rows)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
space = sys.argv[5] if len(sys.argv) > 5 else "l2sqr"
if space == "negdotprod":
    data, q = synth.embedding_like(n, dim, 9), synth.embedding_like(nq, dim, 10)
elif space == "cosinesimil":
    data, q = synth.gist_like(n, dim, 5), synth.gist_like(nq, dim, 6)
else:
    data, q = synth.sift_like_f32(n, 3, dim), synth.sift_like_f32(nq, 4, dim)
if space == "l2sqr_sift":
    data, q = synth.sift_like_u8(n, 7), synth.sift_like_u8(nq, 8)
    idx = nb.Index(space, None, "seq_search", "DenseUInt8Vector", "Int")
    idx.addUInt8Batch(data)
else:
    idx = nb.Index(space, None, "seq_search")
    idx.addDenseBatch(data)
idx.buildIndex()
for _ in range(3):
    idx.knnQueryBatch(q, k)
    st = idx.stats()
    print("scan_ms", st["last_scan_ms"], "total_ms", st["last_total_ms"], "fallback", st["fallback_queries"], "split",
          st["split_queries"], flush=True)
idx.deinit()
