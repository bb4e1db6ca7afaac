"""Timing experiments on the tensor-core scan: environment switches -> scan time (needs the experiments library:
python -m nmslib_zig_b200.build --experiments; this script selects it through NB200_LIB).
usage: tc_exp.py [dim] [n] [nq] [space]   (SIFT-shaped data, k = 10; space l2sqr or l2sqr_sift)"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
os.environ.setdefault("NB200_LIB", str(ROOT / "nmslib_zig_b200" / "lib" / "libnmslib_b200_exp.so"))
sys.path.insert(0, str(ROOT))
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
space = sys.argv[4] if len(sys.argv) > 4 else "l2sqr"
k = 10
u8 = space == "l2sqr_sift"
if u8:
    data, q = synth.sift_like_u8(n, 7), synth.sift_like_u8(nq, 8)
else:
    data, q = synth.sift_like_f32(n, 3, dim), synth.sift_like_f32(nq, 4, dim)
settings = [{}, {"NB200_TC_COUNT": "1"}, {"NB200_TC_DEBUG": "1"}, {"NB200_TC_DEBUG": "3"}]
if os.environ.get("TC_EXP_SETTINGS"):
    settings = json.loads(os.environ["TC_EXP_SETTINGS"])
KNOBS = ("NB200_TC_DEBUG", "NB200_TC_HWM", "NB200_TC_WARM", "NB200_TC_REFRESH", "NB200_TC_MARGIN", "NB200_TC_NO_TS",
         "NB200_TC_COUNT", "NB200_TC_NO_REG", "NB200_U8_IMMA")
for s in settings:
    for kk in KNOBS:
        os.environ.pop(kk, None)
    os.environ.update({a: b for a, b in s.items() if not a.startswith("opt:")})
    for a, b in s.items():       # "opt:<name>": a documented variant selector (nmslib_b200_set_option)
        if a.startswith("opt:"):
            nb.set_option(a[4:], int(b))
    idx = (nb.Index(space, None, "seq_search", "DenseUInt8Vector", "Int") if u8 else nb.Index(space, None, "seq_search"))
    (idx.addUInt8Batch if u8 else idx.addDenseBatch)(data)
    idx.buildIndex()
    ms = []
    for _ in range(4):
        idx.knnQueryBatch(q, k)
        ms.append(idx.stats()["last_scan_ms"])
    tf = 2.0 * nq * n * dim / (min(ms) * 1e-3) / 1e12
    print(f"dim={dim} n={n} nq={nq} {space} {s}: scan_ms min={min(ms):.3f} -> {tf:.0f} T(FL)OP/s  "
          f"fallback={idx.stats()['fallback_queries']}", flush=True)
    idx.deinit()
