"""Quick A/B of the tensor-core scan against the exact scan on the same inputs (GPU box only)."""
import os as _os
from pathlib import Path as _Path
# the NB200_* switches exist only in the experiments build (python -m nmslib_zig_b200.build --experiments)
_os.environ.setdefault("NB200_LIB", str(_Path(__file__).resolve().parents[1] / "nmslib_zig_b200" / "lib" / "libnmslib_b200_exp.so"))

import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth


def run(space, data, q, k, force_exact):
    nb.set_option("force_exact", 1 if force_exact else 0)
    idx = nb.Index(space, None, "seq_search")
    idx.addDenseBatch(data)
    idx.buildIndex()
    r = idx.knnQueryBatch(q, k)
    t0 = time.perf_counter()
    r = idx.knnQueryBatch(q, k)
    dt = time.perf_counter() - t0
    st = idx.stats()
    idx.deinit()
    return r, st, dt


cases = [("l2", synth.uniform(10_000, 128, 1), synth.uniform(1_000, 128, 2), 10),
         ("l2sqr", synth.sift_like_f32(100_000, 3), synth.sift_like_f32(2_000, 4), 10),
         ("negdotprod", synth.embedding_like(20_000, 768, 9), synth.embedding_like(512, 768, 10), 100),
         ("cosinesimil", synth.gist_like(20_000, 960, 5), synth.gist_like(300, 960, 6), 10),
         ("l2", synth.uniform(3_001, 19, 1) - .25, synth.uniform(77, 19, 2) - .25, 7)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cases = [("l2sqr", *synth.make("c2"), 10)]
for space, data, q, k in cases:
    rt, st, dt = run(space, data, q, k, False)
    re, se, de = run(space, data, q, k, True)
    same_ids = float(np.mean(rt.ids == re.ids))
    maxrel = float(np.max(np.abs(rt.distances - re.distances) / np.maximum(np.abs(re.distances), 1e-6)))
    print(f"{space:12s} n={data.shape[0]} d={data.shape[1]} q={q.shape[0]} k={k}: ids_equal={same_ids:.6f} "
          f"max_rel_dist={maxrel:.2e} fallback={st['fallback_queries']}/{2 * q.shape[0]} "
          f"tc_scan_ms={st['last_scan_ms']:.3f} tc_call_ms={dt * 1e3:.2f} exact_scan_ms={se['last_scan_ms']:.3f} "
          f"exact_call_ms={de * 1e3:.2f}", flush=True)
