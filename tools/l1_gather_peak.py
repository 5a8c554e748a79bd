#!/usr/bin/env python
"""Measure the L1-data-stage ceiling for the MSDeformAttn access pattern (4 x 128-byte rows per warp-level LDG.128).

  table 32 KB   -> every request hits L1            (the ceiling of any fp32 gather formulation on this part)
  table 22.8 MB -> L1 misses served by L2           (a C2 frame's `value`)
Writes profiles/l1_gather_peak.json; bench.py reports the gather kernels against it as `l1_roofline`."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from trackformer_b200 import ext
    lib = ctypes.CDLL(ext.library_path())
    lib.msda_b200_l1_gather_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p]
    dev = torch.device("cuda:0")
    sink = torch.zeros(1 << 22, device=dev)
    out = {}
    for name, rows in (("l1_resident_32KB", 256), ("l2_resident_22.8MB", 177784), ("l1_resident_128KB", 1024)):
        table = torch.randn(rows, 32, device=dev)
        ctas, iters = 148 * 8, 2048
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.msda_b200_l1_gather_probe(table.data_ptr(), sink.data_ptr(), rows, iters, ctas,
                                               torch.cuda.current_stream().cuda_stream)
            e1.record()
            e1.synchronize()
            assert rc == 0
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        nbytes = ctas * 256 * iters * 16
        out[name] = {"rows": rows, "bytes": nbytes, "ms": round(best, 4), "gbs": round(nbytes / best / 1e6, 1)}
        print(name, out[name], flush=True)
    out["note"] = ("LDG.128 requests shaped like the MSDeformAttn gather (4 groups x 8 lanes, one 128-byte row per group); "
                   "best of 5 launches, CUDA events, 148x8 CTAs x 2048 requests per thread")
    path = os.path.join(ROOT, "gpurun_out", "l1_gather_peak.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
