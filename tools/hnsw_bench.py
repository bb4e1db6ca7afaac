"""HNSW efSearch sweep on a B200 box: GPU beam search vs the reference on the SAME graph.

The graph is built here by the unmodified reference (oracle/_ref, test infrastructure -- index
build stays on the reference CPU code), exported with Hnsw::SaveIndex and imported into the
engine.  Reports, per efSearch: recall@k of both (ground truth = exact scan), id agreement,
GPU queries/s (kernel and end-to-end), reference queries/s on all host cores, and the
gather bandwidth the kernel achieved (bytes = evals * 4 * D + expansions * 4 * maxM0).

    python tools/hnsw_bench.py [--n 200000] [--dim 128] [--space l2] [--nq 10000] [--shape gist|sift|emb]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import nmslib_zig_b200 as nb
from helpers import recall
from nmslib_zig_b200 import synth
from oracle import oracle as O


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--space", default="l2")
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--shape", default="gist")
    ap.add_argument("--efs", default="50,100,200,400")
    ap.add_argument("--cpu-queries", type=int, default=2000)
    ap.add_argument("--build", default="ref", choices=["ref", "device"],
                    help="ref: graph built by the unmodified reference and imported (same-graph parity run); "
                         "device: graph built by csrc/hnsw_build_gpu.cu (recall-parity run, no reference timing)")
    args = ap.parse_args()
    threads = os.cpu_count() or 1
    if args.shape == "gist":
        data, q = synth.gist_like(args.n, args.dim, 5), synth.gist_like(args.nq, args.dim, 6)
    elif args.shape == "sift":
        data, q = synth.sift_like_f32(args.n, 3, args.dim), synth.sift_like_f32(args.nq, 4, args.dim)
    else:
        data, q = synth.embedding_like(args.n, args.dim, 9), synth.embedding_like(args.nq, args.dim, 10)

    ref = None
    if args.build == "ref":
        t0 = time.perf_counter()
        ref = O.RefIndex(args.space, "hnsw").add(data).build(f"M=16,efConstruction=200,indexThreadQty={threads}")
        build_s = time.perf_counter() - t0
        path = "/tmp/nb200_bench.hnsw"
        ref.save(path)

    # ground truth from the exact GPU scan (itself parity-tested against the oracle)
    ex = nb.Index(args.space, None, "seq_search")
    ex.addDenseBatch(data)
    ex.buildIndex()
    exact_ids = ex.knnQueryBatch(q, args.k).ids
    ex.deinit()

    idx = nb.Index(args.space, None, "hnsw")
    build_info = {}
    if ref is not None:
        idx.importHnsw(path)
        idx.prepare()
    else:
        idx.addDenseBatch(data)
        idx.buildIndex(nb.Params({"M": 16, "efConstruction": 200, "b200_build": "device"}))
        t0 = time.perf_counter()
        idx.prepare()                      # upload + device build
        build_s = time.perf_counter() - t0
        st = idx.stats()
        build_info = {k_: st[k_] for k_ in st if k_.startswith("build_")}
    out = []
    for ef in [int(e) for e in args.efs.split(",")]:
        idx.setQueryTimeParams(nb.Params({"efSearch": ef}))
        idx.knnQueryBatch(q[:256], args.k)  # warm
        idx.knnQueryBatch(q, args.k)
        s0 = idx.stats()
        t0 = time.perf_counter()
        r = idx.knnQueryBatch(q, args.k)
        e2e_s = time.perf_counter() - t0
        s1 = idx.stats()
        kern_ms = s1["scan_ms_sum"] - s0["scan_ms_sum"]  # (a large batch is searched in several chunks)
        evals = s1["distance_evals"] - s0["distance_evals"]
        exps = s1["hnsw_expansions"] - s0["hnsw_expansions"]
        gbytes = (evals * 4.0 * args.dim + exps * 4.0 * 32) / 1e9
        nqc = min(args.cpu_queries, args.nq)
        if ref is not None:
            ref.set_query_params(f"efSearch={ef}")
            t0 = time.perf_counter()
            ri, rd, rc = ref.knn(q[:nqc], args.k, threads=threads)
            cpu_s = time.perf_counter() - t0
        else:
            line = {"space": args.space, "n": args.n, "dim": args.dim, "nq": args.nq, "k": args.k, "ef": ef,
                    "graph": "device-built", "recall_gpu": recall(r.ids, exact_ids),
                    "gpu_kernel_ms": kern_ms, "gpu_kernel_qps": args.nq / (kern_ms * 1e-3),
                    "gpu_e2e_qps": args.nq / e2e_s, "evals_per_query": evals / args.nq,
                    "expansions_per_query": exps / args.nq, "gather_GBps": gbytes / (kern_ms * 1e-3),
                    "build_s_device_incl_upload": build_s, **build_info}
            print(json.dumps(line), flush=True)
            continue
        line = {"space": args.space, "n": args.n, "dim": args.dim, "nq": args.nq, "k": args.k, "ef": ef,
                "recall_gpu": recall(r.ids, exact_ids), "recall_ref": recall(ri, exact_ids[:nqc]),
                "recall_gpu_same_queries": recall(r.ids[:nqc], exact_ids[:nqc]),
                "id_agreement": float(np.mean(r.ids[:nqc] == ri)),
                "gpu_kernel_ms": kern_ms, "gpu_kernel_qps": args.nq / (kern_ms * 1e-3),
                "gpu_e2e_qps": args.nq / e2e_s, "ref_cpu_qps": nqc / cpu_s, "ref_cores": threads,
                "evals_per_query": evals / args.nq, "expansions_per_query": exps / args.nq,
                "gather_GBps": gbytes / (kern_ms * 1e-3), "build_s_ref": build_s}
        print(json.dumps(line), flush=True)
        out.append(line)
    idx.deinit()


if __name__ == "__main__":
    main()
