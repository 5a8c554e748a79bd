#!/usr/bin/env python
"""Eager-mode profile of one training step grouped by (op, input shapes): which tensors do the memory-bound
elementwise / normalisation kernels run over?  Diagnostic tool."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
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
    frames = torch.randn(1, 3, bench.H, bench.W, device=dev)
    targets = bench.make_targets(1, dev, 2)
    step = TrainStep(model, criterion, None, use_graphs=False)
    for _ in range(3):
        step(frames, targets)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step(frames, targets)
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=70,
                                                             max_name_column_width=40, max_shapes_column_width=70))


if __name__ == "__main__":
    main()
