"""Scripted detector + frame blobs that drive BOTH the reference Tracker (tests/golden/make_golden_tracker.py, build
container only) and trackformer_b200.tracker.Tracker through the same synthetic scene.

The detector stands in for DeformableDETRTracking in tracking mode: it takes (img, targets, prev_features) and returns
(outputs, None, features, None, None) with `pred_logits`, `pred_boxes`, `hs_embed`; track queries arrive through
targets[0]['track_query_hs_embeds' / 'track_query_boxes'] and are answered first, object queries last -- the contract
of src/trackformer/models/tracker.py:306-330.  Its answers depend only on the frame number and on WHICH objects the
incoming track queries follow (decoded from their embeddings), so two trackers that make the same decisions see the
same detector outputs, and any divergence in the bookkeeping shows up in the results.

The scene exercises: births and deaths, occlusions (track -> inactive -> back through its re-identification query, or
through a fresh detection matched by embedding / centre distance), duplicate detections of tracked objects (both NMS
passes), label flips (class != 0), a zero-area box, and public-detection gating.
"""
import numpy as np
import torch

IMG_H, IMG_W = 480, 640
N_FRAMES = 16

BASE_CFG = dict(public_detections=False, detection_obj_score_thresh=0.4, track_obj_score_thresh=0.4,
                detection_nms_thresh=0.9, track_nms_thresh=0.9, steps_termination=1, prev_frame_dist=1,
                inactive_patience=-1, reid_sim_threshold=0.0, reid_sim_only=False, reid_score_thresh=0.4,
                reid_greedy_matching=False)        # cfgs/track.yaml:27-50

CASES = {
    "default": dict(cfg={}, overflow=True),
    "reid_embedding": dict(cfg=dict(inactive_patience=5, reid_sim_threshold=1.0, detection_nms_thresh=0.6,
                                    track_nms_thresh=0.6, steps_termination=2), overflow=True),
    "reid_greedy": dict(cfg=dict(inactive_patience=5, reid_greedy_matching=True, detection_nms_thresh=0.7), overflow=True),
    "public_center": dict(cfg=dict(public_detections="center_distance", inactive_patience=2), overflow=False),
    "public_iou": dict(cfg=dict(public_detections="min_iou_0_5", inactive_patience=3, reid_score_thresh=0.3,
                                prev_frame_dist=2), overflow=False),
    "reid_sim_only": dict(cfg=dict(reid_sim_only=True, inactive_patience=5, reid_sim_threshold=1.0), overflow=True),
}


def tracker_cfg(case):
    cfg = dict(BASE_CFG)
    cfg.update(CASES[case]["cfg"])
    return cfg


class Scene:
    """Object trajectories and per-frame scripts, all derived from one seed (numpy, host)."""

    def __init__(self, n_objects=9, hidden_dim=32, seed=0):
        rng = np.random.RandomState(seed)
        self.n, self.c = n_objects, hidden_dim
        self.start = rng.uniform(0.15, 0.85, (n_objects, 2)).astype(np.float32)
        self.vel = rng.uniform(-0.012, 0.012, (n_objects, 2)).astype(np.float32)
        self.wh = rng.uniform(0.08, 0.22, (n_objects, 2)).astype(np.float32)
        self.birth = rng.randint(0, 4, n_objects)
        self.birth[:3] = 0
        self.death = rng.randint(10, N_FRAMES + 4, n_objects)
        emb = rng.standard_normal((n_objects, hidden_dim)).astype(np.float32)
        self.emb = 4.0 * emb / np.linalg.norm(emb, axis=1, keepdims=True)
        # occlusion windows [lo, hi): the object's score is low for both its track query and the object queries
        self.occl = np.zeros((n_objects, 2), np.int64)
        for j in range(n_objects):
            lo = rng.randint(4, 9)
            self.occl[j] = (lo, lo + rng.randint(1, 4)) if j % 2 == 0 else (0, 0)
        # after an occlusion: even/4 objects come back through their (inactive) track query, the others only through a
        # fresh object-query detection
        self.back_by_query = np.array([j % 4 == 0 for j in range(n_objects)])
        self.flip = {(1, 6), (1, 7), (3, 9)}                 # (object, frame): best class is not 0
        self.dup = {(0, 2): 0.004, (2, 3): 0.03, (5, 8): 0.004, (7, 11): 0.002, (1, 12): 0.05}   # duplicate detection, shift
        self.flat = {(5, 5)}                                 # zero-width box
        self.noise = rng.standard_normal((N_FRAMES + 8, n_objects, hidden_dim)).astype(np.float32) * 0.04
        self.jitter = rng.uniform(-0.002, 0.002, (N_FRAMES + 8, n_objects, 4)).astype(np.float32)
        self.hi = rng.uniform(1.0, 3.0, (N_FRAMES + 8, n_objects)).astype(np.float32)
        self.lo = rng.uniform(-3.0, -1.0, (N_FRAMES + 8, n_objects)).astype(np.float32)

    def alive(self, j, t):
        return self.birth[j] <= t < self.death[j]

    def occluded(self, j, t):
        return self.occl[j, 0] <= t < self.occl[j, 1]

    def box(self, j, t):
        """cx, cy, w, h normalised"""
        c = self.start[j] + self.vel[j] * t
        b = np.concatenate([c, self.wh[j]]) + self.jitter[t, j]
        if (j, t) in self.flat:
            b[2] = 0.0
        return b.astype(np.float32)

    def public_dets(self, t):
        """pixel xyxy boxes of the visible objects except every third one"""
        out = []
        for j in range(self.n):
            if self.alive(j, t) and not self.occluded(j, t) and j % 3 != 2:
                cx, cy, w, h = self.box(j, t) + np.float32(0.003)
                out.append([(cx - w / 2) * IMG_W, (cy - h / 2) * IMG_H, (cx + w / 2) * IMG_W, (cy + h / 2) * IMG_H])
        return np.asarray(out, np.float32).reshape(-1, 4)


class ScriptedDetector(torch.nn.Module):
    def __init__(self, scene: Scene, num_queries=20, num_classes=20, overflow_boxes=True):
        super().__init__()
        self.scene = scene
        self.num_queries = num_queries
        self.num_classes = num_classes
        self.hidden_dim = scene.c
        self.overflow_boxes = overflow_boxes
        self.anchor = torch.nn.Parameter(torch.zeros(1))
        self.frame = 0
        self.prev_log = []
        rng = np.random.RandomState(99)
        self.slot_of = np.stack([rng.permutation(num_queries)[:scene.n] for _ in range(N_FRAMES + 8)])
        self.clutter = rng.uniform(0.1, 0.9, (N_FRAMES + 8, num_queries, 4)).astype(np.float32)
        self.clutter[..., 2:] *= 0.2

    def forward(self, img, targets=None, prev_features=None):
        sc, t, dev = self.scene, self.frame, self.anchor.device
        self.frame += 1
        self.prev_log.append(None if prev_features is None else prev_features[1])
        followed = []
        if targets is not None:
            assert targets[0]["track_query_boxes"].shape[1] == 4
            hs = targets[0]["track_query_hs_embeds"].detach().float().cpu().numpy()
            followed = list((hs @ sc.emb.T).argmax(1))
        k, nq = len(followed), self.num_queries
        logits = np.full((k + nq, self.num_classes), -5.0, np.float32)
        boxes = np.zeros((k + nq, 4), np.float32)
        embeds = np.zeros((k + nq, sc.c), np.float32)

        def answer(row, j, visible):
            boxes[row] = sc.box(j, t)
            embeds[row] = sc.emb[j] + sc.noise[t, j]
            logits[row, 0] = sc.hi[t, j] if visible else sc.lo[t, j]
            if (j, t) in sc.flip:
                logits[row, 3] = 4.0

        for row, j in enumerate(followed):
            visible = sc.alive(j, t) and not sc.occluded(j, t)
            if visible and t >= sc.occl[j, 1] > 0 and t < sc.occl[j, 1] + 3 and not sc.back_by_query[j]:
                # this object's old track query stays silent after the occlusion; it must be re-detected
                visible = False
            answer(row, j, visible)
        # object queries: clutter everywhere, then the scripted detections
        boxes[k:] = self.clutter[t]
        embeds[k:] = 0.01
        logits[k:, 0] = -3.0
        tracked_now = {j for row, j in enumerate(followed)
                       if logits[row, 0] > 0 and logits[row, 3] < 0}
        for j in range(sc.n):
            if not sc.alive(j, t) or sc.occluded(j, t):
                continue
            row = k + self.slot_of[t, j]
            if j not in tracked_now:
                answer(row, j, True)
            elif (j, t) in sc.dup:
                answer(row, j, True)
                boxes[row, :2] += sc.dup[(j, t)]
                embeds[row] += 0.02
        out = {
            "pred_logits": torch.from_numpy(logits)[None].to(dev),
            "pred_boxes": torch.from_numpy(boxes)[None].to(dev),
            "hs_embed": torch.from_numpy(embeds)[None].to(dev),
        }
        return out, None, ("features", t), None, None


def blobs(scene: Scene, n_frames=N_FRAMES):
    for t in range(n_frames):
        yield {
            "img": torch.zeros(1, 3, 32, 32),
            "orig_size": torch.tensor([[IMG_H, IMG_W]]),
            "dets": torch.from_numpy(scene.public_dets(t))[None],
        }


def summarise(tracker, detector=None):
    """Everything observable about a finished run, as arrays (works for the reference Tracker and for ours)."""
    rows = []
    for tid in sorted(tracker.results):
        for frame in sorted(tracker.results[tid]):
            r = tracker.results[tid][frame]
            rows.append([tid, frame, int(r["obj_ind"]), float(r["score"])] + [float(v) for v in r["bbox"]])
    out = {
        "rows": np.asarray(rows, np.float64).reshape(-1, 8),
        "num_reids": np.int64(tracker.num_reids),
        "track_num": np.int64(tracker.track_num),
        "frame_index": np.int64(tracker.frame_index),
        "active_ids": np.asarray([t.id for t in tracker.tracks], np.int64),
        "inactive_ids": np.asarray([t.id for t in tracker.inactive_tracks], np.int64),
        "inactive_counts": np.asarray([t.count_inactive for t in tracker.inactive_tracks], np.int64),
    }
    if detector is not None:
        out["prev_log"] = np.asarray([-1 if p is None else p for p in detector.prev_log], np.int64)
    return out


def run_case(tracker_cls, post, case, device="cpu", seed=0):
    scene = Scene(seed=seed)
    det = ScriptedDetector(scene, overflow_boxes=CASES[case]["overflow"]).to(device)
    tracker = tracker_cls(det, {"bbox": post}, tracker_cfg(case), False)
    tracker.reset()
    per_frame = []
    for blob in blobs(scene):
        tracker.step(blob)
        per_frame.append([t.id for t in tracker.tracks] + [-1] + [t.id for t in tracker.inactive_tracks])
    out = summarise(tracker, det)
    width = max(len(p) for p in per_frame)
    out["per_frame_ids"] = np.asarray([p + [-2] * (width - len(p)) for p in per_frame], np.int64)
    return out


# ---------------------------------------------------------------------------------- the real detector under the tracker
def model_frames(size, n_frames, seed=5):
    """A slowly changing random 'video': base image plus a growing perturbation."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(3, *size, generator=g)
    drift = torch.randn(3, *size, generator=g)
    return [base + 0.03 * t * drift for t in range(n_frames)]


class _ScoreLog:
    """Wraps a post-processor and records every score / label it hands to the tracker (to measure decision margins)."""

    def __init__(self, post):
        self.post, self.scores, self.labels = post, [], []

    def __call__(self, outputs, sizes):
        res = self.post(outputs, sizes)
        self.scores.append(res[0]["scores"].detach().cpu().numpy().copy())
        self.labels.append(res[0]["labels"].detach().cpu().numpy().copy())
        return res


def run_model_sequence(build, tracker_cls, post, cfg_overrides, size=(128, 160), n_frames=5, device="cpu",
                       multi_frame=False, log_scores=False):
    import model_fixtures as mf
    model, _ = build(True, multi_frame)
    mf.canonical_weights_(model, 0)
    with torch.no_grad():                      # random weights favour an arbitrary class: make it class 0 ("person")
        for name, p in model.named_parameters():
            if "class_embed" in name and name.endswith("bias"):
                p[0] += 6.0
    model.to(device)
    model.tracking()
    cfg = dict(BASE_CFG)
    cfg.update(cfg_overrides)
    post = _ScoreLog(post) if log_scores else post
    tracker = tracker_cls(model, {"bbox": post}, cfg, False)
    tracker.reset()
    with torch.no_grad():                         # src/track.py runs the tracker under no_grad
        for frame in model_frames(size, n_frames):
            tracker.step({"img": frame[None], "orig_size": torch.tensor([[size[0] * 4, size[1] * 4]]),
                          "dets": torch.zeros(1, 0, 4)})
    out = summarise(tracker)
    if log_scores:
        out["_scores"], out["_labels"] = post.scores, post.labels
    return out
