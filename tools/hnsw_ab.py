"""A/B timing of the HNSW search kernel under environment switches (NB200_HNSW_PF, NB200_HNSW_PFADJ, ...).
    python tools/hnsw_ab.py build [n] [dim]     # device-built cosine graph -> /tmp/nb200_ab.hnsw (+ queries)
    NB200_HNSW_PF=0 python tools/hnsw_ab.py run  # loads it, efSearch 100 / 400, kernel ms over 10 K queries
The switches are read once per process, hence one process per setting."""
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

P = "/tmp/nb200_ab.hnsw"
if sys.argv[1] == "build":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
    dim = int(sys.argv[3]) if len(sys.argv) > 3 else 960
    data = synth.gist_like(n, dim, 5)
    idx = nb.Index("cosinesimil", None, "hnsw")
    idx.addDenseBatch(data)
    idx.buildIndex(nb.Params({"M": 16, "efConstruction": 200, "b200_build": "device"}))
    idx.prepare()
    idx.save(P, False)
    np.save("/tmp/nb200_ab_q.npy", synth.gist_like(10_000, dim, 6))
    print("built", n, dim, idx.stats()["build_total_ms"], "ms")
else:
    q = np.load("/tmp/nb200_ab_q.npy")
    if len(sys.argv) > 2:
        q = q[:int(sys.argv[2])]
    idx = nb.Index("cosinesimil", None, "hnsw")
    idx.importHnsw(P)
    idx.prepare()
    tag = {k: v for k, v in os.environ.items() if k.startswith("NB200_HNSW")}
    for ef in (100, 400):
        idx.setQueryTimeParams(nb.Params({"efSearch": ef}))
        idx.knnQueryBatch(q, 10)
        s0 = idx.stats()
        r = idx.knnQueryBatch(q, 10)
        s1 = idx.stats()
        ms = s1["scan_ms_sum"] - s0["scan_ms_sum"]
        ev = s1["distance_evals"] - s0["distance_evals"]
        print(tag, "nq", len(q), "ef", ef, "kernel_ms %.2f" % ms, "GB/s %.0f" % (ev * 4.0 * q.shape[1] / 1e6 / ms), "ids_sum", int(r.ids.sum()))
