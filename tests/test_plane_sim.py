"""CPU: the tiling plans and index arithmetic of the plane conv kernels (csrc/conv_plane.cu), replayed in numpy
(tests/plane_sim.py) against torch convolutions.  No GPU compute: only the host-side planner of the library
(mrssm_pl_describe / mrssm_pl_packed_shape) is called."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import plane_sim as S

# (n_img, Hl, Cl, Hs, Cs, ksz): small-channel versions of every spatial geometry on the path
DOWN = [(3, 64, 3, 31, 16, 4), (5, 31, 16, 14, 16, 4), (7, 14, 16, 6, 32, 4), (20, 6, 16, 2, 16, 4),
        (2, 64, 3, 30, 16, 6), (3, 30, 16, 13, 16, 6), (5, 13, 16, 5, 32, 5), (1, 128, 3, 63, 16, 4),
        # 64 small channels: the weight gradient stacks two row taps per MMA (mrep)
        (3, 30, 16, 13, 64, 6), (5, 31, 16, 14, 64, 4), (7, 14, 8, 6, 64, 4)]
UP = [(2, 64, 3, 30, 16, 6), (3, 30, 8, 13, 16, 6), (5, 13, 16, 5, 32, 5), (5, 31, 8, 14, 16, 4),
      (7, 14, 16, 6, 32, 4), (20, 6, 8, 2, 16, 4), (1, 128, 3, 62, 16, 6)]


def _lib():
    from mrssm_b200 import _lib as L
    return L


def _pad(c, m):
    return (c + m - 1) // m * m


@pytest.mark.parametrize("g", DOWN)
def test_down_plan_matches_conv2d(g):
    n, Hl, Cl, Hs, Cs, k = g
    L = _lib()
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, Cl, Hl, Hl, generator=gen)
    w = torch.randn(Cs, Cl, k, k, generator=gen)
    ref = F.conv2d(x, w, stride=2)                                  # [n, Cs, Hs, Hs]
    assert ref.shape[-1] == Hs
    Clp, Csp = _pad(Cl, 8), _pad(Cs, 16)
    src = np.zeros((n, Hl, Hl, Clp), np.float32)
    src[..., :Cl] = x.permute(0, 2, 3, 1).numpy()
    out, P = S.sim_fwd(L, 0, src, w.numpy(), Hl, Hl, Hs, Hs, Csp)
    assert not np.isnan(out).any(), P["text"]
    np.testing.assert_allclose(out[..., :Cs], ref.permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-4)
    assert np.abs(out[..., Cs:]).max(initial=0.0) == 0.0


@pytest.mark.parametrize("g", UP)
def test_up_plan_matches_conv_transpose2d(g):
    n, Hl, Cl, Hs, Cs, k = g
    L = _lib()
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(n, Cs, Hs, Hs, generator=gen)
    w = torch.randn(Cs, Cl, k, k, generator=gen)
    full = F.conv_transpose2d(x, w, stride=2)                        # [n, Cl, 2(Hs-1)+k, ..]
    ref = torch.zeros(n, Cl, Hl, Hl)
    hh = min(Hl, full.shape[-1])
    ref[:, :, :hh, :hh] = full[:, :, :hh, :hh]                       # floor geometries: the extra row/col gets zero
    Csp, Clp = _pad(Cs, 16), _pad(Cl, 8)
    src = np.zeros((n, Hs, Hs, Csp), np.float32)
    src[..., :Cs] = x.permute(0, 2, 3, 1).numpy()
    out, P = S.sim_fwd(L, 1, src, w.numpy(), Hl, Hl, Hs, Hs, Clp)
    assert not np.isnan(out).any(), P["text"]
    np.testing.assert_allclose(out[..., :Cl], ref.permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("g", DOWN)
def test_wgrad_plan_matches_autograd(g):
    n, Hl, Cl, Hs, Cs, k = g
    L = _lib()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, Cl, Hl, Hl, generator=gen)
    w = torch.zeros(Cs, Cl, k, k, requires_grad=True)
    gy = torch.randn(n, Cs, Hs, Hs, generator=gen)
    F.conv2d(x, w, stride=2).backward(gy)
    Clp, Csp = _pad(Cl, 8), _pad(Cs, 8)
    large = np.zeros((n, Hl, Hl, Clp), np.float32)
    large[..., :Cl] = x.permute(0, 2, 3, 1).numpy()
    small = np.zeros((n, Hs, Hs, Csp), np.float32)
    small[..., :Cs] = gy.permute(0, 2, 3, 1).numpy()
    dW, P = S.sim_wgrad(L, small, large, k)
    np.testing.assert_allclose(dW[:Cs, :Cl], w.grad.numpy(), rtol=1e-3, atol=1e-3)


def test_full_size_plans_fit_shared_memory():
    """Every layer of the 64x64 and 128x128 stacks at bench batch sizes gets a plan within 227 KB."""
    L = _lib()
    n = 50176
    enc = [(64, 8, 31, 32, 4), (31, 32, 14, 64, 4), (14, 64, 6, 128, 4), (6, 128, 2, 256, 4)]
    dec = [(13, 64, 5, 128, 5), (30, 32, 13, 64, 6), (64, 8, 30, 32, 6)]
    big = [(128, 8, 63, 16, 4), (63, 16, 30, 32, 4), (30, 32, 14, 64, 4), (14, 64, 6, 128, 4), (6, 128, 2, 256, 4),
           (14, 128, 6, 256, 4), (30, 64, 14, 128, 4), (62, 32, 30, 64, 4), (128, 8, 62, 32, 6)]
    for (Hl, Cl, Hs, Cs, k) in enc + dec + big:
        geom = (n, Hl, Hl, Cl, Hs, Hs, Cs, k)
        for op, npad in ((0, Cs), (1, Cl), (2, 0)):
            if op == 1 and Cs % 16:
                continue
            P = S.plan(L, geom, op, npad)
            assert P["smem"] <= 227 * 1024, P["text"]


S2D = [(3, 64, 3, 31, 16, 4), (2, 64, 3, 30, 16, 6), (2, 63, 3, 30, 16, 4), (1, 128, 3, 63, 16, 4), (4, 14, 4, 6, 16, 4)]


@pytest.mark.parametrize("g", S2D)
def test_down_and_wgrad_over_space_to_depth_source(g):
    """<= 4-channel images enter as their space-to-depth form (16 channels): the first conv and its weight gradient."""
    n, Hl, Cl, Hs, Cs, k = g
    L = _lib()
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(n, Cl, Hl, Hl, generator=gen)
    w = torch.randn(Cs, Cl, k, k, generator=gen, requires_grad=True)
    y = F.conv2d(x, w, stride=2)
    assert y.shape[-1] == Hs
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    src = S.s2d16(x.permute(0, 2, 3, 1).numpy(), Cl)
    Csp = _pad(Cs, 16)
    out, P = S.sim_fwd(L, 0, src, w.detach().numpy(), Hl, Hl, Hs, Hs, Csp, s2d_cq=Cl)
    assert not np.isnan(out).any(), P["text"]
    np.testing.assert_allclose(out[..., :Cs], y.detach().permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-4)
    small = np.zeros((n, Hs, Hs, Csp), np.float32)
    small[..., :Cs] = gy.permute(0, 2, 3, 1).numpy()
    dW, P = S.sim_wgrad(L, small, src, k, s2d_cq=Cl, Hl=Hl)
    np.testing.assert_allclose(dW[:Cs, :Cl], w.grad.numpy(), rtol=1e-3, atol=1e-3)
