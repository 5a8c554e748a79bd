"""Online multi-object tracker driven by track queries.

Host-side mirror of src/trackformer/models/tracker.py:16-581 (`Tracker`, `Track`): same constructor, `reset`, `step(blob)`,
`get_results`, the same `tracker_cfg` keys (cfgs/track.yaml:27-50) and the same decisions frame by frame -- score
thresholds for keeping / re-identifying / starting tracks (tracker.py:337-373, 425-436), both NMS passes (388-406,
494-515), public-detection gating (122-164), ReID by embedding distance or greedy centre distance (166-264), the
`reid_sim_only` mode (547-548) and the result dictionary (533-545).

What is different is where the state lives.  The reference keeps one Python `Track` object per target whose fields are
device tensors, and reads them back one `.cpu()` at a time (several host synchronisations per track per frame).  Here
the per-target state is a struct of arrays:

  * output embeddings (the only state the detector consumes) stay in ONE device tensor, `_hs[slot]`, gathered into the
    next frame's track queries with one index_select and updated with one index_copy per frame;
  * boxes, scores, counters and ids live in host arrays (fp32 / int64, the very values the detector produced);
  * a frame costs ONE device->host copy: the packed rows {score, label, x0, y0, x1, y1} written by the fused
    post-processing kernel (csrc/track_post.cu) -- plus one more (the distance matrix) only when ReID has candidates.

All decisions are then taken on those host arrays in fp32, in the reference's order, so track ids, `obj_ind`,
`num_reids` and the frame sets of the results are identical and boxes / scores are the detector's values bit for bit.
"""
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

__all__ = ["Tracker", "Track", "nms_keep", "pairwise_iou"]


# ------------------------------------------------------------------------------------------------ fp32 box helpers
def _to_cxcywh(xyxy: np.ndarray) -> np.ndarray:
    """util/box_ops.py:18-22 in fp32."""
    x0, y0, x1, y1 = (xyxy[..., k] for k in range(4))
    return np.stack([(x0 + x1) / np.float32(2), (y0 + y1) / np.float32(2), x1 - x0, y1 - y0], axis=-1).astype(np.float32)


def _clip(xyxy: np.ndarray, height, width) -> np.ndarray:
    """torchvision.ops.boxes.clip_boxes_to_image for size (height, width)."""
    out = xyxy.copy()
    out[..., 0::2] = np.clip(out[..., 0::2], np.float32(0), np.float32(width))
    out[..., 1::2] = np.clip(out[..., 1::2], np.float32(0), np.float32(height))
    return out


def pairwise_iou(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """torchvision.ops.box_iou in fp32: [len(a), len(b)]."""
    a, b = a.astype(np.float32, copy=False), b.astype(np.float32, copy=False)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, np.float32(0), None)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def nms_keep(boxes: np.ndarray, scores: np.ndarray, threshold: float) -> np.ndarray:
    """Boolean keep mask of greedy NMS, torchvision.ops.nms semantics: visit boxes by descending score (stable, so equal
    scores -- e.g. the +inf given to established tracks, tracker.py:503 -- keep their list order) and suppress every later
    box whose IoU with a kept one EXCEEDS `threshold`; arithmetic in fp32 like the torchvision kernels."""
    n = len(boxes)
    keep = np.zeros(n, dtype=bool)
    if n == 0:
        return keep
    boxes = boxes.astype(np.float32, copy=False)
    order = np.argsort(-scores.astype(np.float32, copy=False), kind="stable")
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    dead = np.zeros(n, dtype=bool)
    thr = np.float32(threshold)
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep[i] = True
        rest = order[pos + 1:]
        rest = rest[~dead[rest]]
        if len(rest) == 0:
            continue
        w = np.clip(np.minimum(boxes[i, 2], boxes[rest, 2]) - np.maximum(boxes[i, 0], boxes[rest, 0]), np.float32(0), None)
        h = np.clip(np.minimum(boxes[i, 3], boxes[rest, 3]) - np.maximum(boxes[i, 1], boxes[rest, 1]), np.float32(0), None)
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            iou = inter / (area[i] + area[rest] - inter)
        dead[rest[iou > thr]] = True
    return keep


# ---------------------------------------------------------------------------------------------------- track views
class Track:
    """Read-only view of one target with the reference's field names (tracker.py:555-581).  `pos`, `score` and `obj_ind`
    are host values; `hs_embed` is a one-element list holding the device embedding."""

    def __init__(self, tracker, slot):
        s = tracker._state
        self.id = int(s.ids[slot])
        self.pos = torch.from_numpy(s.pos[slot].copy())
        self.last_pos = deque([torch.from_numpy(s.anchor[slot].copy())])
        self.score = torch.tensor(s.score[slot])
        self.count_inactive = int(s.count_inactive[slot])
        self.count_termination = int(s.count_termination[slot])
        self.obj_ind = torch.tensor([int(s.obj_ind[slot])])
        self.hs_embed = [tracker._embedding_of(slot)]
        self.gt_id = None
        self.mask = None
        self.attention_map = None

    def has_positive_area(self) -> bool:
        return bool(self.pos[2] > self.pos[0] and self.pos[3] > self.pos[1])


class _State:
    """Host half of the struct of arrays; grows by doubling."""

    def __init__(self, capacity=64):
        self.capacity = capacity
        self.pos = np.zeros((capacity, 4), np.float32)
        self.anchor = np.zeros((capacity, 4), np.float32)      # = last_pos[-1] of the reference's Track
        self.score = np.zeros(capacity, np.float32)
        self.obj_ind = np.zeros(capacity, np.int64)
        self.ids = np.full(capacity, -1, np.int64)
        self.count_inactive = np.zeros(capacity, np.int64)
        self.count_termination = np.zeros(capacity, np.int64)

    def grow(self):
        for name in ("pos", "anchor", "score", "obj_ind", "ids", "count_inactive", "count_termination"):
            old = getattr(self, name)
            new = np.zeros((2 * self.capacity,) + old.shape[1:], old.dtype)
            if name == "ids":
                new[:] = -1
            new[:self.capacity] = old
            setattr(self, name, new)
        self.capacity *= 2


class Tracker:
    """Drop-in for trackformer.models.tracker.Tracker (same arguments, tracker.py:19-20)."""

    def __init__(self, obj_detector, obj_detector_post, tracker_cfg, generate_attention_maps=False, logger=None,
                 verbose=False):
        self.obj_detector = obj_detector
        self.obj_detector_post = obj_detector_post
        self.detection_obj_score_thresh = tracker_cfg["detection_obj_score_thresh"]
        self.track_obj_score_thresh = tracker_cfg["track_obj_score_thresh"]
        self.detection_nms_thresh = tracker_cfg["detection_nms_thresh"]
        self.track_nms_thresh = tracker_cfg["track_nms_thresh"]
        self.public_detections = tracker_cfg["public_detections"]
        self.inactive_patience = float(tracker_cfg["inactive_patience"])
        self.reid_sim_threshold = tracker_cfg["reid_sim_threshold"]
        self.reid_sim_only = tracker_cfg["reid_sim_only"]
        self.reid_score_thresh = tracker_cfg["reid_score_thresh"]
        self.reid_greedy_matching = tracker_cfg["reid_greedy_matching"]
        self.prev_frame_dist = tracker_cfg["prev_frame_dist"]
        self.steps_termination = tracker_cfg["steps_termination"]
        if generate_attention_maps:
            # the reference asserts the same for the deformable model (tracker.py:38)
            raise ValueError("Generation of attention maps not possible for deformable DETR.")
        if "segm" in obj_detector_post:
            raise NotImplementedError("mask heads are outside the deformable tracking path")
        self.generate_attention_maps = False
        self._logger = logger if logger is not None else (lambda *log_strs: None)
        self._verbose = verbose
        self._pinned = None
        self.reset()

    # ------------------------------------------------------------------------------------------------ accessors
    @property
    def num_object_queries(self):
        return self.obj_detector.num_queries

    @property
    def device(self):
        return next(self.obj_detector.parameters()).device

    @property
    def tracks(self):
        return [Track(self, s) for s in self._active]

    @property
    def inactive_tracks(self):
        return [Track(self, s) for s in self._inactive]

    def get_results(self):
        return self.results

    def reset(self, hard=True):
        """tracker.py:71-80"""
        self._state = _State()
        self._hs = None                       # [capacity, hidden] device embeddings, allocated at the first detection
        self._active, self._inactive = [], []
        self._free = list(range(self._state.capacity - 1, -1, -1))
        self._prev_features = deque([None], maxlen=self.prev_frame_dist)
        if hard:
            self.track_num = 0
            self.results = {}
            self.frame_index = 0
            self.num_reids = 0

    # ------------------------------------------------------------------------------------------------- internals
    def _embedding_of(self, slot):
        return self._hs[slot] if self._hs is not None else None

    def _take_slot(self):
        if not self._free:
            old = self._state.capacity
            self._state.grow()
            self._free = list(range(2 * old - 1, old - 1, -1))
            if self._hs is not None:
                self._hs = torch.cat([self._hs, torch.zeros_like(self._hs)], 0)
        return self._free.pop()

    def _alive(self, slot) -> bool:
        """has_positive_area and still within the patience window (tracker.py:171-174, 270-273)"""
        p = self._state.pos[slot]
        return bool(p[2] > p[0] and p[3] > p[1]) and self._state.count_inactive[slot] <= self.inactive_patience

    def _prune_inactive(self, released):
        kept = [s for s in self._inactive if self._alive(s)]
        released.extend(s for s in self._inactive if s not in kept)
        self._inactive = kept

    def _deactivate(self, slots):
        """tracker.py:86-91: back to the position at the start of the step, then onto the inactive list"""
        gone = set(slots)
        self._active = [s for s in self._active if s not in gone]
        for s in slots:
            self._state.pos[s] = self._state.anchor[s]
        self._inactive += list(slots)

    def _packed_rows(self, outputs, orig_size):
        """One [Q, 6] host array {score, label, x0, y0, x1, y1} for the (single) image of this step."""
        post = self.obj_detector_post["bbox"]
        if hasattr(post, "packed"):
            rows = post.packed(outputs, orig_size)[0]
        else:                                   # any reference-shaped post-processor
            res = post(outputs, orig_size)[0]
            rows = torch.cat([res["scores"][:, None], res["labels"][:, None].float(), res["boxes"]], 1)
        if rows.is_cuda:
            if self._pinned is None or self._pinned.shape[0] < rows.shape[0]:
                self._pinned = torch.empty(max(1024, rows.shape[0]), 6, dtype=torch.float32).pin_memory()
            host = self._pinned[:rows.shape[0]]
            host.copy_(rows, non_blocking=True)
            torch.cuda.current_stream(rows.device).synchronize()
            return host.numpy().copy()
        return rows.detach().float().numpy().copy()

    # ------------------------------------------------------------------------------------- public detections gating
    def public_detections_mask(self, new_boxes, public_boxes):
        """tracker.py:122-164 on host arrays."""
        n = len(new_boxes)
        if not self.public_detections:
            return np.ones(n, dtype=bool)
        mask = np.zeros(n, dtype=bool)
        if not len(public_boxes) or not n:
            return mask
        public_boxes = np.asarray(public_boxes, dtype=np.float32).reshape(-1, 4)
        if self.public_detections == "center_distance":
            size = ((new_boxes[:, 2] - new_boxes[:, 0]) * (new_boxes[:, 3] - new_boxes[:, 1])).astype(np.float32)
            d = _to_cxcywh(new_boxes)[:, None, :2] - _to_cxcywh(public_boxes)[None, :, :2]
            d = (d ** 2).sum(axis=2)
            for j in range(len(public_boxes)):
                i = d[:, j].argmin()
                if d[i, j] < size[i]:
                    d[i, :] = 1e18
                    mask[i] = True
        elif self.public_detections == "min_iou_0_5":
            iou = pairwise_iou(new_boxes, public_boxes)
            for j in range(len(public_boxes)):
                i = iou[:, j].argmax()
                if iou[i, j] >= 0.5:
                    iou[i, :] = 0
                    mask[i] = True
        else:
            raise NotImplementedError
        return mask

    # ------------------------------------------------------------------------------------------------------ ReID
    def reid(self, new_boxes, new_scores, new_query_idx, hs_embeds, hs_writes, released):
        """tracker.py:166-264.  Returns the mask of detections NOT consumed by a re-identified track."""
        st = self._state
        self._prune_inactive(released)
        n = len(new_boxes)
        free_mask = np.ones(n, dtype=bool)
        if not self._inactive or not n:
            return free_mask
        if self.reid_greedy_matching:
            det = _to_cxcywh(new_boxes)
            old = _to_cxcywh(st.pos[self._inactive])
            dist = old[:, None, :2] - det[None, :, :2]
            dist = (dist ** 2).sum(axis=2)
            track_size, item_size = old[:, 2] * old[:, 3], det[:, 2] * det[:, 3]
            invalid = (dist > track_size[:, None]) + (dist > item_size[None, :])
            dist = dist + invalid * 1e18
            work = dist                         # the reference's greedy pass edits the matrix it later thresholds
            rows, cols = [], []
            for i in range(work.shape[0]):
                j = work[i].argmin()
                if work[i][j] < 1e16:
                    work[:, j] = 1e18
                    work[i, j] = 0.0
                    rows.append(i)
                    cols.append(j)
        else:
            # L2 distance of the last output embeddings, F.pairwise_distance semantics (eps inside the norm)
            old = self._hs.index_select(0, torch.as_tensor(self._inactive, device=self._hs.device))
            new = hs_embeds.index_select(0, torch.as_tensor(new_query_idx, device=hs_embeds.device))
            c = old.shape[1]
            dist = F.pairwise_distance(old[:, None, :].expand(-1, n, -1).reshape(-1, c),
                                       new[None, :, :].expand(len(self._inactive), -1, -1).reshape(-1, c))
            dist = dist.view(len(self._inactive), n).cpu().numpy()
            rows, cols = linear_sum_assignment(dist)
        revived = []
        for r, c in zip(rows, cols):
            if dist[r, c] <= self.reid_sim_threshold:
                slot = self._inactive[r]
                self._logger(f"REID: track.id={st.ids[slot]} - count_inactive={st.count_inactive[slot]} - "
                             f"to_inactive_frame={self.frame_index - st.count_inactive[slot]}")
                st.count_inactive[slot] = 0
                st.pos[slot] = new_boxes[c]
                st.anchor[slot] = new_boxes[c]
                st.score[slot] = new_scores[c]
                hs_writes[slot] = int(new_query_idx[c])
                free_mask[c] = False
                revived.append(slot)
                self._active.append(slot)
                self.num_reids += 1
        self._inactive = [s for s in self._inactive if s not in revived]
        return free_mask

    # ------------------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, blob):
        """One frame (tracker.py:266-548)."""
        st = self._state
        released = []                            # slots freed during this step; recycled only afterwards
        hs_writes = {}                           # slot -> query index whose embedding the slot takes this frame
        self._prune_inactive(released)
        self._logger(f"FRAME: {self.frame_index + 1}")
        if self._inactive:
            self._logger(f"INACTIVE TRACK IDS: {[int(st.ids[s]) for s in self._inactive]}")
        for s in self._active:
            st.anchor[s] = st.pos[s]

        dev = self.device
        img = blob["img"].to(dev)
        orig_size = blob["orig_size"]
        height, width = (int(v) for v in orig_size[0].tolist())
        orig_size = orig_size.to(dev)

        order = self._active + self._inactive
        n_prev = len(order)
        target = None
        if n_prev:
            q_boxes = _to_cxcywh(st.pos[order]) / np.array([width, height, width, height], dtype=np.float32)
            target = [{
                "track_query_boxes": torch.from_numpy(q_boxes).to(dev),
                "image_id": torch.tensor([1]).to(dev),
                "track_query_hs_embeds": self._hs.index_select(0, torch.as_tensor(order, device=dev)),
            }]

        outputs, _, features, _, _ = self.obj_detector(img, target, self._prev_features[0])
        hs_embeds = outputs["hs_embed"][0]
        rows = self._packed_rows(outputs, orig_size)
        scores, labels = rows[:, 0], rows[:, 1]
        boxes = rows[:, 2:6] if self.obj_detector.overflow_boxes else _clip(rows[:, 2:6], height, width)
        nq = self.num_object_queries
        n_total = rows.shape[0]
        if self._hs is None:
            self._hs = torch.zeros(st.capacity, hs_embeds.shape[1], dtype=hs_embeds.dtype, device=hs_embeds.device)

        # ---- established tracks and re-identification queries (tracker.py:329-406)
        if n_prev:
            t_scores, t_boxes, t_person = scores[:-nq], boxes[:-nq], labels[:-nq] == 0
            keep = np.logical_and(t_scores > self.track_obj_score_thresh, t_person)
            to_inactive = []
            for i, s in enumerate(self._active):
                if keep[i]:
                    st.score[s], st.pos[s] = t_scores[i], t_boxes[i]
                    hs_writes[s] = i
                    st.count_termination[s] = 0
                else:
                    st.count_termination[s] += 1
                    if st.count_termination[s] >= self.steps_termination:
                        to_inactive.append(s)
            keep = np.logical_and(t_scores > self.reid_score_thresh, t_person)
            from_inactive = []
            for i, s in enumerate(self._inactive, start=len(self._active)):
                if keep[i]:
                    st.score[s], st.pos[s] = t_scores[i], t_boxes[i]
                    hs_writes[s] = i
                    from_inactive.append(s)
            if to_inactive:
                self._logger(f"NEW INACTIVE TRACK IDS (track_obj_score_thresh={self.track_obj_score_thresh}): "
                             f"{[int(st.ids[s]) for s in to_inactive]}")
            self.num_reids += len(from_inactive)
            self._inactive = [s for s in self._inactive if s not in from_inactive]
            self._active += from_inactive
            self._deactivate(to_inactive)

            if self.track_nms_thresh and self._active:
                ok = nms_keep(st.pos[self._active], st.score[self._active], self.track_nms_thresh)
                dropped = [s for s, k in zip(self._active, ok) if not k]
                if dropped:
                    self._logger(f"REMOVE TRACK IDS (track_nms_thresh={self.track_nms_thresh}): "
                                 f"{[int(st.ids[s]) for s in dropped]}")
                self._active = [s for s, k in zip(self._active, ok) if k]
                released += dropped

        # ---- new detections (tracker.py:408-492)
        first_obj = n_total - nq
        d_keep = np.logical_and(scores[-nq:] > self.detection_obj_score_thresh, labels[-nq:] == 0)
        d_idx = np.nonzero(d_keep)[0]
        d_boxes, d_scores = boxes[-nq:][d_idx], scores[-nq:][d_idx]
        public = blob["dets"][0] if "dets" in blob else []
        if torch.is_tensor(public):
            public = public.detach().cpu().numpy()
        sel = self.public_detections_mask(d_boxes, public)
        d_idx, d_boxes, d_scores = d_idx[sel], d_boxes[sel], d_scores[sel]
        sel = self.reid(d_boxes, d_scores, d_idx + first_obj, hs_embeds, hs_writes, released)
        d_idx, d_boxes, d_scores = d_idx[sel], d_boxes[sel], d_scores[sel]

        new_slots = []
        for i in range(len(d_idx)):
            s = self._take_slot()
            st = self._state                     # _take_slot may have re-allocated the arrays
            st.ids[s] = self.track_num + i
            st.pos[s] = st.anchor[s] = d_boxes[i]
            st.score[s] = d_scores[i]
            st.obj_ind[s] = d_idx[i]
            st.count_inactive[s] = st.count_termination[s] = 0
            hs_writes[s] = int(d_idx[i] + first_obj)
            new_slots.append(s)
        self._active += new_slots
        self.track_num += len(new_slots)
        if new_slots:
            self._logger(f"INIT TRACK IDS (detection_obj_score_thresh={self.detection_obj_score_thresh}): "
                         f"{[int(st.ids[s]) for s in new_slots]}")

        if self.detection_nms_thresh and self._active:
            fresh = set(new_slots)
            nms_scores = np.array([st.score[s] if s in fresh else np.inf for s in self._active], dtype=np.float32)
            ok = nms_keep(st.pos[self._active], nms_scores, self.detection_nms_thresh)
            dropped = [s for s, k in zip(self._active, ok) if not k]
            if dropped:
                self._logger(f"REMOVE TRACK IDS (detection_nms_thresh={self.detection_nms_thresh}): "
                             f"{[int(st.ids[s]) for s in dropped]}")
            self._active = [s for s, k in zip(self._active, ok) if k]
            released += dropped

        # ---- results (tracker.py:533-548)
        for s in self._active:
            box = st.pos[s] if self.obj_detector.overflow_boxes else _clip(st.pos[s], height, width)
            self.results.setdefault(int(st.ids[s]), {})[self.frame_index] = {
                "bbox": box.copy(), "score": np.array(st.score[s]), "obj_ind": int(st.obj_ind[s])}
        for s in self._inactive:
            st.count_inactive[s] += 1
        self.frame_index += 1
        self._prev_features.append(features)
        if self.reid_sim_only:
            self._deactivate(list(self._active))

        # ---- one scatter of this frame's embeddings into the bank, then recycle the freed slots
        live = set(self._active) | set(self._inactive)
        writes = [(s, q) for s, q in hs_writes.items() if s in live]
        if writes:
            slots = torch.as_tensor([s for s, _ in writes], device=hs_embeds.device)
            qidx = torch.as_tensor([q for _, q in writes], device=hs_embeds.device)
            self._hs.index_copy_(0, slots, hs_embeds.index_select(0, qidx))
        for s in released:
            st.ids[s] = -1
        self._free.extend(released)
