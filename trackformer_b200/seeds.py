"""Per-step pool of dropout seeds for the mask-free (hash RNG) fused kernels.

Every fused dropout site needs one 64-bit seed per call; drawing it with its own ``torch.randint`` costs one tiny
launch per site (36 per training step of the R50 model).  ``begin_step`` draws a whole pool with ONE launch from the
device generator (graph-safe: a captured ``randint`` advances the generator's Philox offset on every replay, so every
replayed step sees fresh seeds); call sites then take consecutive one-element views.  Outside a ``begin_step`` scope
(plain module calls, tests) each site falls back to its own draw -- same distribution, one launch more.
"""
from __future__ import annotations

import torch

_POOL_SIZE = 128
_pool = None
_next = 0


def begin_step(device) -> None:
    """Draw the pool for one forward(+backward) pass; the views handed out stay valid until the next call."""
    global _pool, _next
    _pool = torch.randint(0, 2 ** 62, (_POOL_SIZE,), dtype=torch.int64, device=device)
    _next = 0


def end_step() -> None:
    global _pool
    _pool = None


def next_seed(device) -> torch.Tensor:
    """A one-element int64 tensor holding a fresh seed."""
    global _next
    if _pool is not None and _pool.device == torch.device(device) and _next < _POOL_SIZE:
        _next += 1
        return _pool[_next - 1:_next]
    return torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=device)
