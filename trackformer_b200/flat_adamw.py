"""AdamW over flat parameter / gradient / moment buffers -- one kernel launch per learning-rate group.

Stands where the reference has ``clip_grad_norm_`` + ``torch.optim.AdamW`` (src/trackformer/engine.py:147-151,
src/train.py:100-119): the same update rule and the same parameter groups (default lr, ``lr_backbone`` for names
containing ``backbone.0``, ``lr * lr_linear_proj_mult`` for ``reference_points`` / ``sampling_offsets``), but over
the flat buffers TrainStep lays the model out in, with the clip coefficient folded into the update (csrc/flat_adamw.cu).
``param_groups`` keeps the torch layout (list of dicts with ``lr`` / ``weight_decay`` / ``params``) so learning-rate
schedulers that edit ``group['lr']`` keep working.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch

from . import ext

__all__ = ["FlatAdamW", "reference_param_groups"]


def reference_param_groups(model, lr=2e-4, lr_backbone=2e-5, lr_backbone_names=("backbone.0",),
                           lr_linear_proj_names=("reference_points", "sampling_offsets"), lr_linear_proj_mult=0.1,
                           weight_decay=1e-4) -> List[dict]:
    """The three groups of src/train.py:100-110 with the defaults of cfgs/train.yaml:1-10."""
    def hit(name, keys):
        return any(k in name for k in keys)
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    special = tuple(lr_backbone_names) + tuple(lr_linear_proj_names) + ("layers_track_attention",)
    return [
        {"params": [p for n, p in named if not hit(n, special)], "lr": lr, "weight_decay": weight_decay},
        {"params": [p for n, p in named if hit(n, lr_backbone_names)], "lr": lr_backbone, "weight_decay": weight_decay},
        {"params": [p for n, p in named if hit(n, lr_linear_proj_names)], "lr": lr * lr_linear_proj_mult,
         "weight_decay": weight_decay},
    ]


class FlatAdamW:
    """``ranges[i] = (begin, end)`` is the slice of the flat buffers holding ``groups[i]['params']`` (begin % 4 == 0)."""

    def __init__(self, groups: Sequence[dict], ranges: Sequence[tuple], flat_param: torch.Tensor,
                 flat_grad: torch.Tensor, betas=(0.9, 0.999), eps: float = 1e-8):
        assert len(groups) == len(ranges)
        self.param_groups = [dict(g, betas=tuple(betas), eps=eps) for g in groups]
        self.ranges = [tuple(int(v) for v in r) for r in ranges]
        self.flat_param, self.flat_grad = flat_param, flat_grad
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.steps = 0

    @torch.no_grad()
    def step(self, grad_norm: Optional[torch.Tensor] = None, max_norm: float = 0.0) -> None:
        """One update; with ``grad_norm`` (device scalar) the gradient is clipped to ``max_norm`` on the fly."""
        self.steps += 1
        op = ext.load().flat_adamw
        for group, (begin, end) in zip(self.param_groups, self.ranges):
            if end > begin:
                b1, b2 = group["betas"]
                op(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, begin, end, grad_norm,
                   float(max_norm), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                   float(group["weight_decay"]), self.steps)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.flat_grad.zero_()

    def state_dict(self) -> dict:
        return {"steps": self.steps, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, state: dict) -> None:
        self.steps = int(state["steps"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        for g, s in zip(self.param_groups, state["param_groups"]):
            g.update(s)
