#!/usr/bin/env python
"""Online-tracking latency on one GPU: Tracker.step per frame, eager detector vs CUDA-graph replay.

Workload: tracking model (deformable DETR R50, 20-class head), one 3x800x1333 frame per
step, `--tracks` live track queries (default 100) + 300 object queries, random-init weights, synthetic frames.  Each
timed step is the whole `Tracker.step(blob)`: host frame -> device, forward, fused post-processing, one read-back, host
bookkeeping.  Random weights give flat scores, so the live-track count is pinned by thresholds: the first frame starts
exactly `--tracks` tracks, afterwards every track is kept and no new ones start.

Prints one JSON line; tools/README.md lists it among the measurement aids.  Not a bench.py arm.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=100)
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--multi-frame", action="store_true")
    ap.add_argument("--bucket", type=int, default=32)
    ap.add_argument("--no-tf32", action="store_true")
    ap.add_argument("--skip-eager", action="store_true", help="only the CUDA-graph detector")
    args = ap.parse_args()
    from trackformer_b200.deformable_detr import DeformablePostProcess
    from trackformer_b200.graphed_detector import GraphedDetector
    from trackformer_b200.model_factory import build_model, default_args
    from trackformer_b200.device_tracker import DeviceTracker
    from trackformer_b200.tracker import Tracker

    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = not args.no_tf32
    torch.manual_seed(0)
    model, _, _ = build_model(default_args(True, args.multi_frame, device="cuda:0"))
    with torch.no_grad():
        for name, p in model.named_parameters():               # make class 0 the winning label (random weights)
            if "class_embed" in name and name.endswith("bias"):
                p[0] += 8.0
    model.to(dev).eval()
    model.tracking()
    g = torch.Generator().manual_seed(1)
    frames = [torch.randn(1, 3, args.height, args.width, generator=g).pin_memory() for _ in range(4)]
    cfg = dict(public_detections=False, detection_obj_score_thresh=2.0, track_obj_score_thresh=-1.0,
               detection_nms_thresh=0.0, track_nms_thresh=0.0, steps_termination=1, prev_frame_dist=1,
               inactive_patience=-1, reid_sim_threshold=0.0, reid_sim_only=False, reid_score_thresh=2.0,
               reid_greedy_matching=False)
    size = torch.tensor([[args.height, args.width]])

    def run(detector, tracker_cls=Tracker):
        post = DeformablePostProcess()
        tr = tracker_cls(detector, {"bbox": post}, cfg, False)
        tr.reset()
        # frame 0: start exactly --tracks tracks
        with torch.no_grad():
            out = detector(frames[0].to(dev), None, None)[0]
        rows = post.packed(out, size.to(dev))[0].cpu().numpy()
        s = np.sort(rows[rows[:, 1] == 0, 0])[::-1]
        assert len(s) > args.tracks, "not enough class-0 queries to start the requested number of tracks"
        tr.detection_obj_score_thresh = float((s[args.tracks - 1] + s[args.tracks]) / 2)
        tr.step({"img": frames[0], "orig_size": size, "dets": torch.zeros(1, 0, 4)})
        tr.detection_obj_score_thresh = 2.0
        assert len(tr.tracks) == args.tracks, len(tr.tracks)
        times = []
        for i in range(args.warmup + args.frames):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tr.step({"img": frames[(i + 1) % 4], "orig_size": size, "dets": torch.zeros(1, 0, 4)})
            torch.cuda.synchronize()
            if i >= args.warmup:
                times.append((time.perf_counter() - t0) * 1e3)
        assert len(tr.tracks) == args.tracks
        return times

    det = GraphedDetector(model, bucket=args.bucket)
    graphed = run(det)
    eager = graphed if args.skip_eager else run(model)
    device = run(GraphedDetector(model, bucket=args.bucket), DeviceTracker)      # decisions in csrc/track_step.cu
    line = {
        "what": "Tracker.step latency, host frame in -> results out", "frames": args.frames, "warmup": args.warmup,
        "config": {"workload": f"tracking{' multi-frame' if args.multi_frame else ''} 1x3x{args.height}x{args.width}, "
                               f"{args.tracks} track + {model.num_queries} object queries", "bucket": args.bucket,
                   "tf32": not args.no_tf32},
        "eager_ms": {"median": float(np.median(eager)), "p10": float(np.percentile(eager, 10)), "p90": float(np.percentile(eager, 90))},
        "graph_ms": {"median": float(np.median(graphed)), "p10": float(np.percentile(graphed, 10)), "p90": float(np.percentile(graphed, 90))},
        "graph_device_tracker_ms": {"median": float(np.median(device)), "p10": float(np.percentile(device, 10)),
                                    "p90": float(np.percentile(device, 90))},
        "eager_fps": 1e3 / float(np.median(eager)), "graph_fps": 1e3 / float(np.median(graphed)),
        "graph_device_tracker_fps": 1e3 / float(np.median(device)),
        "graph_captures": det.captures, "graph_replays": det.replays,
        "h2d_bytes_per_frame": int(frames[0].numel() * 4), "d2h_bytes_per_frame": int((args.tracks + model.num_queries) * 6 * 4),
        "d2h_bytes_per_frame_device_tracker": int((8 + 8 * (args.tracks + model.num_queries)) * 4),
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
