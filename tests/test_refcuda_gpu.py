"""GPU-vs-GPU parity against the reference's OWN CUDA kernels (compiled for sm_100a from /root/reference by
oracle/Makefile `ref` in the build container; the shared object travels to the GPU box).  Skipped when it was not built."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(cuda_device):
    from oracle import refcuda
    if not refcuda.available():
        pytest.skip("oracle/_ref/libmsda_refcuda.so not built (needs /root/reference)")
    return refcuda


CASES = {
    "c1_enc": (1, 8, 32, [(60, 80), (30, 40), (15, 20), (8, 10)], 4, None),
    "c2_dec_n2": (2, 8, 32, [(100, 167), (50, 84), (25, 42), (13, 21)], 4, 300),
    "d36_l8": (1, 8, 36, [(17, 30), (9, 15), (5, 8), (3, 4)] * 2, 4, 311),
    "ref_test_shape": (2, 2, 4, [(8, 8), (4, 4), (2, 2)], 2, 3),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_reference_cuda_kernels(ref, cuda_device, name):
    from trackformer_b200 import ext
    msda = ext.load()
    N, M, D, hw, P, Lq = CASES[name]
    g = torch.Generator().manual_seed(len(name))
    shapes = torch.as_tensor(hw, dtype=torch.long)
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    Lq = S if Lq is None else Lq
    L = len(hw)
    dev = cuda_device
    value = torch.randn(N, S, M, D, generator=g).to(dev)
    loc = (torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.2 - 0.1).to(dev)
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P).to(dev)
    gout = torch.randn(N, Lq, M * D, generator=g).to(dev)
    shapes = shapes.to(dev)
    out = msda.ms_deform_attn_forward(value, shapes, loc, attn, 64)
    out_ref = ref.forward(value, shapes, loc, attn)
    torch.testing.assert_close(out, out_ref, rtol=1e-4, atol=1e-4)
    gv, gl, ga = msda.ms_deform_attn_backward(value, shapes, loc, attn, gout, 64)
    rv, rl, ra = ref.backward(value, shapes, loc, attn, gout)
    scale = max(1.0, float(rv.abs().max()))
    torch.testing.assert_close(gv, rv, rtol=1e-4, atol=1e-4 * scale)
    torch.testing.assert_close(ga, ra, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gl, rl, rtol=5e-4, atol=5e-3)
