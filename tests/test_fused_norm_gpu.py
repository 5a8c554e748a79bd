"""GPU tests of the fused residual + dropout + LayerNorm kernels against the PyTorch op chain the reference
executes (deformable_transformer.py:284-285, 291-292): forward values, all four gradients, dropout-mask
handling, row-count tails, the supported widths, and the Python-level fallback for unsupported widths."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_device):
    torch.backends.cuda.matmul.allow_tf32 = False
    return cuda_device


def reference_chain(x, branch, mask, gamma, beta, keep, eps):
    b = branch if mask is None else branch * mask.to(branch.dtype) / keep
    return F.layer_norm(x + b, (x.shape[-1],), gamma, beta, eps)


@pytest.mark.parametrize("rows,C", [(1, 256), (7, 128), (22223, 256), (301, 384), (64, 512), (4097, 256), (5000, 288), (33, 36)])
@pytest.mark.parametrize("with_mask", [False, True])
def test_forward_backward_match_torch(dev, rows, C, with_mask):
    from trackformer_b200 import ext
    m = ext.load()
    g = torch.Generator(device="cpu").manual_seed(rows * 7 + C)
    x = torch.randn(rows, C, generator=g).to(dev)
    br = torch.randn(rows, C, generator=g).to(dev) * 2
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    beta = torch.randn(C, generator=g).to(dev)
    dy = torch.randn(rows, C, generator=g).to(dev)
    keep = 0.9
    mask = (torch.rand(rows, C, generator=g) < keep).to(dev) if with_mask else None
    y, s, mean, rstd = m.add_dropout_layernorm_forward(x, br, mask, gamma, beta, keep, 1e-5)
    xr, brr, gr, ber = (t.clone().requires_grad_(True) for t in (x, br, gamma, beta))
    y_ref = reference_chain(xr, brr, mask, gr, ber, keep, 1e-5)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    dx, dbr, dgamma, dbeta = m.add_dropout_layernorm_backward(dy, s, mask, gamma, mean, rstd, keep)
    y_ref.backward(dy)
    torch.testing.assert_close(dx, xr.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dbr, brr.grad, rtol=1e-4, atol=1e-5)
    scale = max(1.0, float(gr.grad.abs().max()))
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-4, atol=1e-4 * scale)
    torch.testing.assert_close(dbeta, ber.grad, rtol=1e-4, atol=1e-4 * scale)
    # deterministic column reduction
    again = m.add_dropout_layernorm_backward(dy, s, mask, gamma, mean, rstd, keep)
    assert torch.equal(again[2], dgamma) and torch.equal(again[3], dbeta)


def test_module_level_wrapper(dev):
    from trackformer_b200.fused_norm import add_dropout_layernorm, supported
    norm = torch.nn.LayerNorm(256).to(dev)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_()
    x = torch.randn(2, 300, 256, device=dev, requires_grad=True)
    br = torch.randn(2, 300, 256, device=dev, requires_grad=True)
    drop = torch.nn.Dropout(0.1).eval()
    assert supported(x, norm)
    y = add_dropout_layernorm(x, br, drop, norm)
    y_ref = norm(x + br)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    gx, gb = torch.autograd.grad(y.square().sum(), (x, br))
    rx, rb = torch.autograd.grad(y_ref.square().sum(), (x, br))
    torch.testing.assert_close(gx, rx, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(gb, rb, rtol=1e-4, atol=1e-4)
    # training mode: inverted dropout statistics (mean preserved, ~10 % zeros in the branch gradient)
    drop.train()
    y_t = add_dropout_layernorm(x, br, drop, norm)
    gb_t, = torch.autograd.grad(y_t.sum() + (y_t * torch.randn_like(y_t)).sum(), (br,))
    zero_frac = float((gb_t == 0).float().mean())
    assert 0.07 < zero_frac < 0.13
    # hidden 288 (multi-frame TrackFormer) is inside the kernels' domain; widths that are not a multiple of 4 are not
    norm288 = torch.nn.LayerNorm(288).to(dev)
    x288 = torch.randn(3, 5, 288, device=dev)
    assert supported(x288, norm288)
    torch.testing.assert_close(add_dropout_layernorm(x288, x288, drop.eval(), norm288), norm288(2 * x288), rtol=1e-5, atol=1e-5)
    norm30 = torch.nn.LayerNorm(30).to(dev)
    x30 = torch.randn(3, 5, 30, device=dev)
    assert not supported(x30, norm30)
    torch.testing.assert_close(add_dropout_layernorm(x30, x30, drop.eval(), norm30), norm30(2 * x30))


def test_colsum_and_fused_linear(dev):
    from trackformer_b200 import ext
    from trackformer_b200.fused_linear import linear
    m = ext.load()
    for rows, c in ((22223, 256), (5000, 384), (4097, 1024), (3, 128), (6000, 288), (100, 1000), (300, 256), (2048, 4),
                    (2049, 8), (129, 36)):
        x = torch.randn(rows, c, device=dev)
        got = m.colsum(x)
        torch.testing.assert_close(got, x.double().sum(0).float(), rtol=1e-4, atol=1e-3)
        assert torch.equal(got, m.colsum(x))                               # deterministic
    x = torch.randn(1, 6000, 256, device=dev, requires_grad=True)
    w = torch.randn(384, 256, device=dev, requires_grad=True)
    b = torch.randn(384, device=dev, requires_grad=True)
    y = linear(x, w, b)
    y_ref = torch.nn.functional.linear(x, w, b)
    torch.testing.assert_close(y, y_ref)
    gy = torch.randn_like(y)
    g = torch.autograd.grad(y, (x, w, b), gy)
    r = torch.autograd.grad(y_ref, (x, w, b), gy)
    for a_, b_ in zip(g, r):
        torch.testing.assert_close(a_, b_, rtol=1e-4, atol=1e-3)


def test_relu_dropout(dev):
    from trackformer_b200.fused_linear import relu_dropout
    a = torch.randn(1, 22223, 1024, device=dev, requires_grad=True)
    drop = torch.nn.Dropout(0.1).eval()
    h = relu_dropout(a, drop)
    torch.testing.assert_close(h, torch.relu(a))
    g, = torch.autograd.grad(h, a, torch.ones_like(h))
    torch.testing.assert_close(g, (a > 0).float())
    drop.train()
    torch.manual_seed(0)
    h1 = relu_dropout(a, drop)
    h2 = relu_dropout(a, drop)
    pos = a > 0
    kept = (h1 > 0)
    assert not (kept & ~pos).any()                                         # never resurrects a negative input
    frac = float(kept[pos].float().mean())
    assert 0.895 < frac < 0.905                                            # keep probability 0.9
    torch.testing.assert_close(h1[kept], (a[kept] / 0.9))                  # inverted-dropout scaling
    assert not torch.equal(h1 > 0, h2 > 0)                                 # a fresh mask every call
    # per-column / per-row keep rates are flat (no visible structure in the hash)
    col = kept.float().sum(1) / pos.float().sum(1).clamp(min=1)
    assert float((col - 0.9).abs().max()) < 0.02
    g1, = torch.autograd.grad(h1, a, torch.ones_like(h1))
    torch.testing.assert_close(g1, kept.float() / 0.9)


@pytest.mark.parametrize("ref_dim,L,P,Lq", [(2, 4, 4, 22223), (4, 4, 4, 300), (2, 8, 4, 777), (4, 2, 4, 64), (2, 1, 4, 33)])
def test_sampling_prep_matches_torch_chain(dev, ref_dim, L, P, Lq):
    """One fused pass == the reference's split / softmax / normalise / add chain (ms_deform_attn.py:69-82)."""
    from trackformer_b200.msda_module import _SamplingPrep
    M, N = 8, 2 if Lq < 1000 else 1
    g = torch.Generator().manual_seed(L * 10 + P + ref_dim)
    proj = torch.randn(N, Lq, 3 * M * L * P, generator=g).to(dev).requires_grad_(True)
    ref = torch.rand(N, Lq, L, ref_dim, generator=g).to(dev)
    shapes = torch.tensor([(100 // (l + 1) + 3, 167 // (l + 1) + 2) for l in range(L)], dtype=torch.long, device=dev)
    loc, attn = _SamplingPrep.apply(proj, ref, shapes.float(), M, L, P)
    # reference chain
    p2 = proj.detach().clone().requires_grad_(True)
    n_off = M * L * P * 2
    off = p2[..., :n_off].reshape(N, Lq, M, L, P, 2)
    a_ref = torch.softmax(p2[..., n_off:].reshape(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    if ref_dim == 2:
        l_ref = ref[:, :, None, :, None, :] + off / shapes[None, None, None, :, None, :]
    else:
        l_ref = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
    torch.testing.assert_close(loc, l_ref, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(attn, a_ref, rtol=1e-5, atol=1e-7)
    gl, ga = torch.randn_like(loc), torch.randn_like(attn)
    (g1,) = torch.autograd.grad([loc, attn], [proj], [gl, ga])
    (g2,) = torch.autograd.grad([l_ref, a_ref], [p2], [gl, ga])
    torch.testing.assert_close(g1, g2, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("rows,C,p", [(22223, 256, 0.1), (4097, 288, 0.3), (33, 128, 0.5)])
def test_seeded_dropout_layernorm(dev, rows, C, p):
    """mask-free dropout inside the fused residual+LayerNorm: the kept set is reproducible from the seed, has the right
    rate, and forward / backward agree with the torch chain run on the SAME kept set"""
    from trackformer_b200 import ext
    m = ext.load()
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=g).to(dev)
    br = torch.randn(rows, C, generator=g).to(dev)
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    beta = torch.randn(C, generator=g).to(dev)
    dy = torch.randn(rows, C, generator=g).to(dev)
    seed = torch.tensor([123456789012345], dtype=torch.int64, device=dev)
    keep = 1.0 - p
    y, s, mean, rstd = m.add_dropout_layernorm_seeded_forward(x, br, seed, gamma, beta, keep, 1e-5)
    # recover the kept set from s = x + branch * kept / keep
    kept = ((s - x).abs() > 0) | (br == 0)
    rate = float(kept.float().mean())
    assert abs(rate - keep) < 4 * (keep * p / kept.numel()) ** 0.5 + 1e-3
    xr, brr = x.clone().requires_grad_(True), br.clone().requires_grad_(True)
    gr, btr = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    s_ref = xr + brr * kept / keep
    y_ref = torch.nn.functional.layer_norm(s_ref, (C,), gr, btr, 1e-5)
    torch.testing.assert_close(y, y_ref, rtol=1e-4, atol=1e-4)
    y_ref.backward(dy)
    dx, dbr, dgamma, dbeta = m.add_dropout_layernorm_seeded_backward(dy, s, seed, gamma, mean, rstd, keep)
    torch.testing.assert_close(dx, xr.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dbr, brr.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dgamma, gr.grad, rtol=1e-3, atol=1e-3 * float(gr.grad.abs().max()))
    torch.testing.assert_close(dbeta, btr.grad, rtol=1e-3, atol=1e-3 * float(btr.grad.abs().max()))
    # same seed -> same kept set; another seed -> another one
    y2 = m.add_dropout_layernorm_seeded_forward(x, br, seed, gamma, beta, keep, 1e-5)[0]
    assert torch.equal(y, y2)
    y3 = m.add_dropout_layernorm_seeded_forward(x, br, seed + 1, gamma, beta, keep, 1e-5)[0]
    assert not torch.equal(y, y3)
