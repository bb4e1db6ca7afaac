"""Timing experiments on the tensor-core scan: environment switches -> scan time.
usage: tc_exp.py [dim] [n] [nq]   (SIFT-shaped l2sqr data, k = 10)"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
k = 10
data, q = synth.sift_like_f32(n, 3, dim), synth.sift_like_f32(nq, 4, dim)
settings = [{}, {"NB200_TC_DEBUG": "1"}, {"NB200_TC_DEBUG": "3"}, {"NB200_TC_DEBUG": "5"}]
if os.environ.get("TC_EXP_SETTINGS"):
    import json
    settings = json.loads(os.environ["TC_EXP_SETTINGS"])
KNOBS = ("NB200_TC_DEBUG", "NB200_TC_HWM", "NB200_TC_WARM", "NB200_TC_REFRESH", "NB200_TC_MARGIN", "NB200_TC_NO_TS")
for s in settings:
    for kk in KNOBS:
        os.environ.pop(kk, None)
    os.environ.update(s)
    idx = nb.Index("l2sqr", None, "seq_search")  # (some knobs are read when the index is created)
    idx.addDenseBatch(data)
    idx.buildIndex()
    ms = []
    for _ in range(4):
        idx.knnQueryBatch(q, k)
        ms.append(idx.stats()["last_scan_ms"])
    tf = 2.0 * nq * n * dim / (min(ms) * 1e-3) / 1e12
    print(f"dim={dim} n={n} nq={nq} {s}: scan_ms min={min(ms):.3f} -> {tf:.0f} TFLOP/s  fallback={idx.stats()['fallback_queries']}",
          flush=True)
    idx.deinit()
