#!/usr/bin/env python
"""Where does a training step spend its time?  Phase timers (host wall-clock with device syncs) and a
torch.profiler kernel table for the CUDA-graph TrainStep.  Diagnostic tool, not part of the benchmark."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    from trackformer_b200.model_factory import build_model, default_args
    from trackformer_b200.train_step import TrainStep
    torch.manual_seed(0)
    model, criterion, _ = build_model(default_args(device=str(dev)))
    model.to(dev).train()
    criterion.to(dev).train()
    frames = torch.randn(args.batch, 3, bench.H, bench.W, device=dev)
    targets = bench.make_targets(args.batch, dev, 2)
    step = TrainStep(model, criterion, lambda ps: torch.optim.AdamW(ps, lr=2e-4, weight_decay=1e-4, fused=True),
                     use_graphs=not args.no_graphs, example_frames=frames)
    for _ in range(3):
        step(frames, targets)
    torch.cuda.synchronize()

    sync = torch.cuda.synchronize
    sync(); t0 = time.perf_counter()
    for _ in range(args.steps):
        step(frames, targets)
    sync()
    print("PIPELINED_MS_PER_STEP", round((time.perf_counter() - t0) * 1e3 / args.steps, 3))

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step(frames, targets)
        sync()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))


if __name__ == "__main__":
    main()
