"""Build libnmslib_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch dependency).

    python -m nmslib_zig_b200.build [--force] [--verbose] [--experiments]

--experiments builds a second library, lib/libnmslib_b200_exp.so, with -DNB200_EXPERIMENTS: the NB200_* timing /
debugging environment switches and the kernels' cycle counters exist only there (tools/ load it through
NB200_LIB=...; the release library never reads them).

The shared library links cudart statically, so it dlopen()s on a machine without a GPU
(the CPU-side tests check that it loads and exports every symbol of include/nmslib_b200.h).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libnmslib_b200.so"
STAMP = LIBDIR / ".build_stamp"

SOURCES = ["scan_exact.cu", "scan_tc.cu", "topk_merge.cu", "range_scan.cu", "hnsw_search.cu", "hnsw_build_gpu.cu", "exchange.cu", "shard_group.cu", "engine.cu", "hnsw_format.cpp", "hnsw_build.cpp",
           "c_abi.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (cand == "nvcc" or Path(cand).exists()):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    for f in sorted(CSRC.iterdir()) + [PKG.parent / "include" / "nmslib_b200.h", Path(__file__)]:
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    lib = LIBDIR / "libnmslib_b200_exp.so" if experiments else LIB
    stamp = LIBDIR / ".build_stamp_exp" if experiments else STAMP
    flags = NVCC_FLAGS + (["-DNB200_EXPERIMENTS"] if experiments else [])
    digest = _digest("exp" if experiments else "")
    if not force and lib.exists() and stamp.exists() and stamp.read_text() == digest:
        return lib
    nvcc = _nvcc()
    objs = []
    procs = []
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's CXX wrapper is not a host compiler nvcc can use
    env.pop("CC", None)
    for src in SOURCES:
        obj = LIBDIR / (("exp_" if experiments else "") + src.rsplit(".", 1)[0] + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *flags, "-x", "cu", "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode(errors="replace")
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *objs,
            "-cudart", "static", "-Xcompiler", "-fPIC"]
    subprocess.run(link, check=True, env=env)
    stamp.write_text(digest)
    return lib


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, experiments="--experiments" in sys.argv)
    print(p)
