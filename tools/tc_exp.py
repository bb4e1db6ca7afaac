"""Timing experiments on the tensor-core scan (config 2): environment switches -> scan time."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
data, q = synth.make(cfg, None if cfg == "c2" else 200_000, None if cfg == "c2" else 2_000)
space, k = synth.CONFIGS[cfg][0], synth.CONFIGS[cfg][7]
idx = nb.Index(space, None, "seq_search")
idx.addDenseBatch(data)
idx.buildIndex()
settings = [{}, {"NB200_TC_L2AHEAD": "0"}, {"NB200_TC_L2AHEAD": "4"}, {"NB200_TC_DEBUG": "1"},
            {"NB200_TC_DEBUG": "1", "NB200_TC_L2AHEAD": "0"}]
for s in settings:
    for kk in ("NB200_TC_DEBUG", "NB200_TC_L2AHEAD"):
        os.environ.pop(kk, None)
    os.environ.update(s)
    ms = []
    for _ in range(4):
        idx.knnQueryBatch(q, k)
        ms.append(idx.stats()["last_scan_ms"])
    print(f"{cfg} {s}: scan_ms min={min(ms):.3f} all={[round(m, 3) for m in ms]} fallback={idx.stats()['fallback_queries']}",
          flush=True)
idx.deinit()
