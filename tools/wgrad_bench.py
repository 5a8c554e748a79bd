#!/usr/bin/env python
"""Weight-gradient GEMM (dW = gy^T @ x over a long token axis): library call vs split-K slabs, per shape and split."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters=30):
    for _ in range(5):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    from trackformer_b200.fused_linear import weight_grad
    out = []
    for k, c_out, c_in in ((22223, 256, 256), (22223, 384, 256), (22223, 1024, 256), (22223, 256, 1024), (44446, 256, 256)):
        gy = torch.randn(k, c_out, device="cuda")
        x = torch.randn(k, c_in, device="cuda")
        rec = {"k": k, "c_out": c_out, "c_in": c_in, "library_us": timed(lambda: gy.t() @ x),
               "auto_us": timed(lambda: weight_grad(gy, x))}
        for s in (2, 4, 8, 16, 32):
            slab = (k // s) & ~7
            main_ = slab * s

            def split():
                gw = torch.bmm(gy[:main_].view(s, slab, c_out).transpose(1, 2), x[:main_].view(s, slab, c_in)).sum(0)
                return gw.addmm_(gy[main_:].t(), x[main_:])
            rec[f"split{s}_us"] = timed(split)
        rec["tflops_library"] = 2 * k * c_out * c_in / rec["library_us"] / 1e6
        rec["tflops_auto"] = 2 * k * c_out * c_in / rec["auto_us"] / 1e6
        out.append(rec)
        print(json.dumps({a: (round(b, 1) if isinstance(b, float) else b) for a, b in rec.items()}))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "wgrad_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
