#!/usr/bin/env python
"""clip + AdamW over the detector's 40.6 M trainable parameters: flat one-pass kernel vs torch.optim.AdamW(fused=True)
preceded by the flat-buffer clip (norm, clamp, mul_), timed back to back in one process with CUDA events."""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from trackformer_b200.flat_adamw import reference_param_groups
    from trackformer_b200.model_factory import build_model, default_args
    from trackformer_b200.train_step import TrainStep
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model, criterion, _ = build_model(default_args(device="cuda:0"))
    model.to(dev).train()
    model_t = copy.deepcopy(model)
    flat = TrainStep(model, criterion, None, max_norm=0.1, use_graphs=False,
                     flat_adamw={"groups": reference_param_groups(model)})
    groups_t = reference_param_groups(model_t)
    ref = TrainStep(model_t, criterion, lambda ps: torch.optim.AdamW(groups_t, lr=2e-4, weight_decay=1e-4, fused=True),
                    max_norm=0.1, use_graphs=False)
    for s in (flat, ref):
        s.flat_grad.normal_(generator=torch.Generator(device=dev).manual_seed(1))
    scrub = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flat_update():
        norm = torch.linalg.vector_norm(flat.flat_grad)
        flat.flat_optimizer.step(norm, 0.1)

    def torch_update():
        norm = torch.linalg.vector_norm(ref.flat_grad)
        ref.flat_grad.mul_(torch.clamp(0.1 / (norm + 1e-6), max=1.0))
        ref.optimizer.step()

    def timed(fn, iters=30):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            scrub.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in ev)
        return t[len(t) // 2] * 1e3

    n = flat.flat_param.numel()
    out = {"parameters": n, "flat_clip_adamw_us": timed(flat_update), "torch_clip_fused_adamw_us": timed(torch_update),
           "flat_clip_adamw_us_again": timed(flat_update)}
    out["flat_gbs"] = (7 * 4 * n + 4 * n) / out["flat_clip_adamw_us"] / 1e3       # norm read + 4 reads + 3 writes
    print(json.dumps(out))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "optimizer_bench.json"), "w"))


if __name__ == "__main__":
    main()
