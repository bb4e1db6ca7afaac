"""nmslib_zig_b200 -- B200-native batched dense-vector k-NN query engine.

A drop-in for the knnQuery / knnQueryBatch path of B-R-P/NMSLIB-ZIG: hand-written sm_100a
CUDA kernels behind the reference's own C ABI (include/nmslib_b200.h), plus a host-side
mirror of lib.zig's `Index` (nmslib_zig_b200.index).  See DESIGN.md.
"""
from .index import (BatchResult, Index, NmslibError, Params, QueryResult, device_available, lib,  # noqa: F401
                    live_allocations, scan_plan, scan_plan_pairs, set_device, set_option, version)
