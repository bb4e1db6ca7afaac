import os as _os
from pathlib import Path as _Path
# the NB200_* switches exist only in the experiments build (python -m nmslib_zig_b200.build --experiments)
_os.environ.setdefault("NB200_LIB", str(_Path(__file__).resolve().parents[1] / "nmslib_zig_b200" / "lib" / "libnmslib_b200_exp.so"))
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import nmslib_zig_b200 as nb
from nmslib_zig_b200 import synth
data, q = synth.embedding_like(30_000, 96, 9), synth.embedding_like(300, 96, 10)
for env in [{}, {"NB200_TC_NO_TS": "1"}]:
    os.environ.update(env)
    for k in (100, 40, 20, 10):
        idx = nb.Index("negdotprod", None, "seq_search"); idx.addDenseBatch(data); idx.buildIndex()
        r = idx.knnQueryBatch(q, k); st = idx.stats()
        print(env, "k", k, "fallback", st["fallback_queries"], "launches", st["kernel_launches"], flush=True)
        idx.deinit()
