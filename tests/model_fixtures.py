"""Shared, deterministic test material for model-level parity (test infrastructure).

The same functions are applied (a) to the REFERENCE model inside the build container by
tests/golden/make_golden_model.py and (b) to the trackformer_b200 model in the tests, so both sides see
bit-identical weights, images and targets without shipping 160 MB of parameters:

  * ``canonical_weights_(model, seed)`` overwrites every parameter from a per-key generator
    (key names are identical on both sides -- that is part of the drop-in contract);
  * ``make_images`` / ``make_targets`` build seeded inputs.
"""
from __future__ import annotations

import hashlib
import math
import re

import torch

# 1-D parameters whose constructor value is a deterministic constant in both implementations
# (norm scales/shifts, explicitly initialised biases).  They keep that value plus a small perturbation;
# every other 1-D parameter (default-initialised Linear biases) is drawn afresh.
_CONST_1D = re.compile(
    r"(norm\d*\.(weight|bias)$)|(input_proj\.\d+\.[01]\.(weight|bias)$)|(sampling_offsets\.bias$)|"
    r"(attention_weights\.bias$)|(value_proj\.bias$)|(output_proj\.bias$)|(class_embed\.\d+\.bias$)|"
    r"(bbox_embed\.\d+\.layers\.2\.bias$)|(reference_points\.bias$)|(in_proj_bias$)|(out_proj\.bias$)")
# weights the reference initialises to exactly zero: small noise instead, so that sampling offsets and
# attention logits actually depend on the queries in the parity runs
_SMALL_2D = re.compile(r"(sampling_offsets\.weight$)|(bbox_embed\.\d+\.layers\.2\.weight$)")


def _gen(seed: int, key: str) -> torch.Generator:
    h = int.from_bytes(hashlib.sha256(f"{seed}:{key}".encode()).digest()[:6], "little")
    return torch.Generator().manual_seed(h)


@torch.no_grad()
def canonical_weights_(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    for name, p in sorted(model.named_parameters()):
        g = _gen(seed, name)
        if p.dim() >= 2:
            fan_out = p.shape[0] * (p[0][0].numel() if p.dim() > 2 else 1)
            fan_in = p.shape[1] * (p[0][0].numel() if p.dim() > 2 else 1)
            bound = 0.02 if _SMALL_2D.search(name) else math.sqrt(6.0 / (fan_in + fan_out))
            new = (torch.rand(p.shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        else:
            noise = (torch.rand(p.shape, generator=g, dtype=torch.float32) * 2 - 1) * 0.05
            new = p.detach().cpu().float() + noise if _CONST_1D.search(name) else noise
        p.copy_(new.to(p.device, p.dtype))
    return model


def make_images(seed: int, sizes, device="cpu"):
    """List of CHW float images, one per (H, W) in ``sizes``."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(3, h, w, generator=g).to(device) for h, w in sizes]


def tensor_digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def make_targets(seed: int, batch: int, n_boxes: int, n_classes: int = 1, device="cpu"):
    """Ground-truth dicts: normalised cxcywh boxes, labels, identities."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(batch):
        cxcy = torch.rand(n_boxes, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n_boxes, 2, generator=g) * 0.25 + 0.05
        out.append({"boxes": torch.cat([cxcy, wh], 1).to(device),
                    "labels": torch.randint(0, n_classes, (n_boxes,), generator=g).to(device),
                    "track_ids": torch.arange(n_boxes).to(device),
                    "image_id": torch.tensor([b]).to(device)})
    return out


# gradients recorded by the golden generator (small tensors only)
GRAD_KEYS = (
    "transformer.level_embed",
    "transformer.reference_points.weight",
    "transformer.encoder.layers.0.self_attn.sampling_offsets.bias",
    "transformer.encoder.layers.5.self_attn.attention_weights.bias",
    "transformer.decoder.layers.0.cross_attn.sampling_offsets.bias",
    "transformer.decoder.layers.5.cross_attn.attention_weights.bias",
    "transformer.decoder.layers.2.norm1.weight",
    "input_proj.0.1.weight",
    "input_proj.3.0.bias",
    "class_embed.5.bias",
    "bbox_embed.0.layers.2.bias",
)


# ---------------------------------------------------------------------------------------------
# Case runners: identical code drives the reference (golden generation) and the product (tests).
# ``build(tracking, multi_frame, **overrides)`` must return (model, criterion) on ``device``.
# ---------------------------------------------------------------------------------------------
def _to_np(x):
    # a COPY: on the CPU `.numpy()` aliases the tensor, and the optimizer leg clips / updates gradients and weights in place
    return x.detach().cpu().numpy().copy()


def run_detection(build, sizes, device="cpu", seed=0, **overrides):
    model, _ = build(False, False, **overrides)
    canonical_weights_(model, seed)
    model.to(device).eval()
    imgs = make_images(seed + 1, sizes, device)
    with torch.no_grad():
        out, _, feats, memory, hs = model(imgs if len(imgs) > 1 else imgs[0][None])
    return {"pred_logits": _to_np(out["pred_logits"]), "pred_boxes": _to_np(out["pred_boxes"]),
            "hs_last_mean": _to_np(out["hs_embed"].mean(-1)),
            "aux4_boxes": _to_np(out["aux_outputs"][4]["pred_boxes"]),
            "memory0_mean": _to_np(memory[0].mean(1)),
            "image_digest": tensor_digest(imgs[0])}


def run_two_frame_tracking(build, size, n_track, device="cpu", seed=0, multi_frame=False, **overrides):
    """Online-tracking style: frame 1 plain, frame 2 with ``n_track`` track queries taken from frame 1
    (what models/tracker.py:286-306 feeds the detector)."""
    model, _ = build(True, multi_frame, **overrides)
    canonical_weights_(model, seed)
    model.to(device)
    model.tracking()
    f1, f2 = make_images(seed + 2, [size, size], device)
    with torch.no_grad():
        out1, _, feats1, _, _ = model(f1[None], None, None)
        tgt = [{"track_query_boxes": out1["pred_boxes"][0, :n_track],
                "track_query_hs_embeds": out1["hs_embed"][0, :n_track],
                "image_id": torch.tensor([1]).to(device)}]
        out2, _, _, memory2, hs2 = model(f2[None], tgt, feats1)
    return {"f1_logits": _to_np(out1["pred_logits"]), "f1_boxes": _to_np(out1["pred_boxes"]),
            "f2_logits": _to_np(out2["pred_logits"]), "f2_boxes": _to_np(out2["pred_boxes"]),
            "f2_hs_mean": _to_np(out2["hs_embed"].mean(-1)), "n_levels_memory": len(memory2)}


def reference_optimizer_groups(model, lr=2e-4, lr_backbone=2e-5, lr_linear_proj_mult=0.1, weight_decay=1e-4):
    """The three AdamW parameter groups of the reference's src/train.py:100-119 with the defaults of
    cfgs/train.yaml:1-10 (names containing 'backbone.0' -> lr_backbone; 'reference_points' /
    'sampling_offsets' -> lr * lr_linear_proj_mult)."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    bb, lin = ("backbone.0",), ("reference_points", "sampling_offsets")

    def hit(n, keys):
        return any(k in n for k in keys)
    return [{"params": [p for n, p in named if not hit(n, bb + lin + ("layers_track_attention",))], "lr": lr,
             "weight_decay": weight_decay},
            {"params": [p for n, p in named if hit(n, bb)], "lr": lr_backbone, "weight_decay": weight_decay},
            {"params": [p for n, p in named if hit(n, lin)], "lr": lr * lr_linear_proj_mult,
             "weight_decay": weight_decay}]


def run_train_step(build, sizes, n_boxes, device="cpu", seed=0, tracking=False, optimizer=False, **overrides):
    """One forward + criterion + backward with dropout disabled (module in train mode).  With ``optimizer``
    the reference's clip_grad_norm_(0.1) + AdamW step (engine.py:147-151, train.py:100-119) follows and the
    updated values of the GRAD_KEYS parameters are recorded as ``param/<key>``."""
    model, criterion = build(tracking, False, dropout=0.0, **overrides)
    canonical_weights_(model, seed)
    model.to(device).train()
    criterion.to(device).train()
    imgs = make_images(seed + 3, sizes, device)
    targets = make_targets(seed + 4, len(sizes), n_boxes, 1, device)
    if tracking:
        prev_imgs = make_images(seed + 5, sizes, device)
        for t, pim in zip(targets, prev_imgs):
            pt = {k: v.clone() for k, v in t.items()}
            pt["boxes"] = (pt["boxes"] + 0.01).clamp(0.05, 0.95)
            # the previous frame misses the last identity and has one the current frame lost
            pt["track_ids"] = torch.cat([t["track_ids"][:-1], torch.tensor([99]).to(device)])
            t["prev_target"] = pt
            t["prev_image"] = pim
    torch.manual_seed(1234)                       # CPU generator consumed by add_track_queries_to_targets
    from_list = imgs if len(imgs) > 1 else imgs[0][None]
    out, targets_out, _, _, _ = model(from_list, targets)
    loss_dict = criterion(out, targets_out)
    wd = criterion.weight_dict
    total = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
    model.zero_grad(set_to_none=True)
    total.backward()
    named = dict(model.named_parameters())
    res = {"loss_total": _to_np(total), "pred_logits": _to_np(out["pred_logits"]),
           "pred_boxes": _to_np(out["pred_boxes"])}
    for k in sorted(loss_dict):
        res["loss/" + k] = _to_np(loss_dict[k])
    for k in GRAD_KEYS:
        if k in named and named[k].grad is not None:
            res["grad/" + k] = _to_np(named[k].grad)
    sq = sum(float((p.grad.double() ** 2).sum()) for p in model.parameters() if p.grad is not None)
    res["grad_global_norm"] = torch.tensor(sq).sqrt().numpy()
    if optimizer:
        opt = torch.optim.AdamW(reference_optimizer_groups(model), lr=2e-4, weight_decay=1e-4)
        before = {k: named[k].detach().clone() for k in GRAD_KEYS if k in named}
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 0.1)
        opt.step()
        for k, b in before.items():
            res["param/" + k] = _to_np(named[k])
            res["param_before/" + k] = _to_np(b)
    if tracking:
        for i, t in enumerate(targets_out):
            for key in ("track_query_match_ids", "track_queries_mask", "track_queries_fal_pos_mask"):
                res[f"idx/{i}/{key}"] = _to_np(t[key].long())
            res[f"idx/{i}/n_track_queries"] = torch.tensor(len(t["track_query_boxes"])).numpy()
    return res


def run_bookkeeping(model, n_queries_prev, seeds=range(6), device="cpu"):
    """Drive add_track_queries_to_targets directly with synthetic previous-frame outputs and matchings;
    returns every index tensor it writes (bit-exact contract)."""
    res = {}
    for seed in seeds:
        g = torch.Generator().manual_seed(1000 + seed)
        batch = 2
        prev_out = {"pred_boxes": torch.rand(batch, n_queries_prev, 4, generator=g).to(device),
                    "hs_embed": torch.randn(batch, n_queries_prev, model.hidden_dim, generator=g).to(device)}
        targets, prev_indices = [], []
        for b in range(batch):
            n_prev, n_cur = 5 + seed + b, 6 + (seed % 3)
            prev_ids = torch.randperm(12, generator=g)[:n_prev]
            cur_ids = torch.randperm(12, generator=g)[:n_cur]
            targets.append({"track_ids": cur_ids.to(device), "prev_target": {"track_ids": prev_ids.to(device)}})
            out_idx = torch.randperm(n_queries_prev, generator=g)[:n_prev].sort()[0]
            prev_indices.append((out_idx, torch.randperm(n_prev, generator=g)))
        torch.manual_seed(77 + seed)
        model.add_track_queries_to_targets(targets, prev_indices, prev_out, add_false_pos=(seed % 2 == 0))
        for b, t in enumerate(targets):
            res[f"s{seed}/b{b}/match_ids"] = _to_np(t["track_query_match_ids"].long())
            res[f"s{seed}/b{b}/track_mask"] = _to_np(t["track_queries_mask"].long())
            res[f"s{seed}/b{b}/fal_pos_mask"] = _to_np(t["track_queries_fal_pos_mask"].long())
            res[f"s{seed}/b{b}/boxes"] = _to_np(t["track_query_boxes"])
            res[f"s{seed}/b{b}/hs_sum"] = _to_np(t["track_query_hs_embeds"].sum(-1))
    return res
