"""CPU coverage of the N>1 path (gloo, world size 2): the flat-buffer gradient all-reduce of
trackformer_b200.train_step.TrainStep and SetCriterion's num_boxes all-reduce.

The MSDeformAttn core has no CPU path, so -- in this test only -- the oracle's torch restatement stands in for
the CUDA function; what is under test is the host-side data-parallel logic: after one step both ranks must hold
identical gradients equal to the mean of the two per-rank gradients computed without communication."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build_small(seed=0):
    sys.path.insert(0, ROOT)
    import trackformer_b200.msda_module as mm
    from oracle.torch_ref import msda_core_torch
    from trackformer_b200.model_factory import build_model, default_args

    class _OracleFn:
        @staticmethod
        def apply(value, shapes, loc, attn, step):
            return msda_core_torch(value, shapes, loc, attn)
    mm.MSDeformAttnFunction = _OracleFn
    torch.manual_seed(seed)
    model, criterion, _ = build_model(default_args(device="cpu", enc_layers=1, dec_layers=2, num_queries=20,
                                                   dropout=0.0))
    return model.train(), criterion.train()


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    frames = torch.randn(1, 3, 96, 128, generator=g)
    n = 3 + rank                                                   # different box counts per rank -> num_boxes all-reduce matters
    boxes = torch.cat([torch.rand(n, 2, generator=g) * 0.5 + 0.25, torch.rand(n, 2, generator=g) * 0.2 + 0.05], 1)
    return frames, [{"boxes": boxes, "labels": torch.zeros(n, dtype=torch.int64)}]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trackformer_b200.train_step import TrainStep
        model, criterion = _build_small()
        step = TrainStep(model, criterion, None, use_graphs=False)
        assert step.world == world
        assert step.phased            # three-piece backward, one all-reduce per finished gradient slice
        frames, targets = _data(rank)
        if rank == 1:                 # ranks may hand their frames over differently: same collectives either way
            step.prefetch(frames)
            loss = step(None, targets)
        else:
            loss = step(frames, targets)
        torch.save({"grad": step.flat_grad.clone(), "loss": loss}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_flat_allreduce_matches_mean_of_local_grads(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert torch.equal(got[0]["grad"], got[1]["grad"])              # every rank ends with the same gradient

    # single-process recomputation: per-rank gradients with the world-averaged num_boxes, then their mean
    sys.path.insert(0, ROOT)
    from trackformer_b200.train_step import TrainStep
    local = []
    total_boxes = sum(len(_data(r)[1][0]["labels"]) for r in range(world))
    for r in range(world):
        model, criterion = _build_small()
        step = TrainStep(model, criterion, None, use_graphs=False)
        frames, targets = _data(r)
        step(frames, targets)
        local.append((step.flat_grad.clone(), len(targets[0]["labels"])))
    # local grads were normalised by their own box count; rescale to the shared normaliser before averaging
    shared = max(total_boxes / world, 1)
    mean = sum(g * (n / shared) for g, n in local) / world
    # class_error / cardinality do not contribute; focal + box losses are all divided by num_boxes
    torch.testing.assert_close(got[0]["grad"], mean, rtol=2e-4, atol=1e-6)
