"""GPU test of the CUDA-graph training step: replaying the captured forward/backward graphs must give the same loss
and the same flat gradient as the eager step, for changing inputs, and the optimizer path must move the weights."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def small_model(dev):
    from trackformer_b200.model_factory import build_model, default_args
    torch.manual_seed(0)
    model, criterion, _ = build_model(default_args(device=str(dev), enc_layers=2, dec_layers=3, num_queries=50,
                                                   dropout=0.0))
    return model.to(dev).train(), criterion.to(dev).train()


def targets_for(dev, seed, n):
    g = torch.Generator().manual_seed(seed)
    boxes = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.2 + 0.05], 1)
    return [{"boxes": boxes.to(dev), "labels": torch.zeros(n, dtype=torch.int64, device=dev)}]


def test_graph_replay_equals_eager(cuda_device):
    from trackformer_b200.train_step import TrainStep
    dev = cuda_device
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, criterion = small_model(dev)
    model_e = copy.deepcopy(model)
    g = torch.Generator().manual_seed(3)
    frames = [torch.randn(1, 3, 192, 256, generator=g).to(dev) for _ in range(3)]
    graphed = TrainStep(model, criterion, None, use_graphs=True, example_frames=frames[0])
    eager = TrainStep(model_e, criterion, None, use_graphs=False)
    for i, f in enumerate(frames):                      # new frame and new ground truth every step
        tg = targets_for(dev, 10 + i, 4 + i)
        loss_g = graphed(f, tg)
        loss_e = eager(f, tg)
        torch.testing.assert_close(loss_g, loss_e, rtol=1e-4, atol=1e-4)
        scale = float(eager.flat_grad.abs().max())
        torch.testing.assert_close(graphed.flat_grad, eager.flat_grad, rtol=1e-3, atol=2e-4 * scale)


def test_full_step_graph_equals_eager(cuda_device):
    """forward + device matching + loss + backward captured as ONE graph; ground truth of the captured box count is
    fed through static buffers, other counts fall back to the two-graph path."""
    from trackformer_b200.train_step import TrainStep
    dev = cuda_device
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, criterion = small_model(dev)
    model_e = copy.deepcopy(model)
    g = torch.Generator().manual_seed(4)
    frames = [torch.randn(1, 3, 192, 256, generator=g).to(dev) for _ in range(3)]
    full = TrainStep(model, criterion, None, use_graphs=True, example_frames=frames[0],
                     example_targets=targets_for(dev, 0, 5))
    assert full.g_full is not None
    eager = TrainStep(model_e, criterion, None, use_graphs=False)
    for i, (f, n) in enumerate(zip(frames, (5, 5, 8))):          # third step: different box count -> fallback path
        tg = targets_for(dev, 20 + i, n)
        loss_g = full(f, tg).clone()
        loss_e = eager(f, tg)
        torch.testing.assert_close(loss_g, loss_e, rtol=1e-4, atol=1e-4)
        scale = float(eager.flat_grad.abs().max())
        torch.testing.assert_close(full.flat_grad, eager.flat_grad, rtol=1e-3, atol=2e-4 * scale)


def test_optimizer_step_updates_weights_and_clips(cuda_device):
    from trackformer_b200.train_step import TrainStep
    dev = cuda_device
    model, criterion = small_model(dev)
    before = [p.detach().clone() for p in model.parameters() if p.requires_grad]
    frames = torch.randn(1, 3, 192, 256, device=dev)
    step = TrainStep(model, criterion, lambda ps: torch.optim.AdamW(ps, lr=1e-3, weight_decay=1e-4, fused=True),
                     max_norm=0.1, use_graphs=True, example_frames=frames)
    loss = step(frames, targets_for(dev, 1, 5))
    assert torch.isfinite(loss)
    assert float(torch.linalg.vector_norm(step.flat_grad)) <= 0.1 * 1.001      # clip_grad_norm_(0.1) semantics
    after = [p for p in model.parameters() if p.requires_grad]
    assert sum(int(not torch.equal(a, b)) for a, b in zip(after, before)) > len(before) // 2
