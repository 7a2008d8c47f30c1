"""-m gpu: tcgen05 conv-family kernels against the exact-fp32 CUDA-core kernels (same C ABI) on
bf16-representable inputs.  Tolerance: bf16 output rounding (rtol 1e-2) — inputs are identical."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (n_img, Hl, Cl, Hs, Cs, ksz): every conv / convT / dense geometry of the 64x64 and 128x128 stacks
GEOMS = [
    (3, 64, 3, 31, 32, 4), (3, 31, 32, 14, 64, 4), (3, 14, 64, 6, 128, 4), (5, 6, 128, 2, 256, 4),   # encoder 64
    (3, 5, 128, 1, 1024, 5), (3, 13, 64, 5, 128, 5), (2, 30, 32, 13, 64, 6), (2, 64, 3, 30, 32, 6),   # decoder 64
    (300, 1, 1024, 1, 200, 1), (130, 1, 128, 1, 128, 1), (70, 1, 200, 1, 1024, 1),                   # dense
    (2, 128, 3, 63, 16, 4), (2, 14, 256, 6, 128, 4), (1, 128, 3, 62, 32, 6),                           # 128x128 stacks
]


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _setup(g):
    from mrssm_b200 import _lib as L, ops
    n, Hl, Cl, Hs, Cs, k = g
    gen = torch.Generator(device=DEV).manual_seed(hash(g) % 1000)
    large = _bf16_round(torch.randn(n, Hl, Hl, Cl, device=DEV, generator=gen))
    small = _bf16_round(torch.randn(n, Hs, Hs, Cs, device=DEV, generator=gen))
    w = _bf16_round(torch.randn(Cs, Cl, k, k, device=DEV, generator=gen) / (Cl * k * k) ** 0.5)
    geom = (n, Hl, Hl, Cl, Hs, Hs, Cs, k)
    Clp, Csp = ops.pad8(Cl), ops.pad8(Cs)
    lb = ops.tc_to_bf16(L.nhwc(large, Hl, Hl, Cl), n, Hl, Hl, Cl, DEV)
    sb = ops.tc_to_bf16(L.nhwc(small, Hs, Hs, Cs), n, Hs, Hs, Cs, DEV)
    return L, ops, geom, (large, small, w), (lb, sb), (Clp, Csp)


@pytest.mark.parametrize("g", GEOMS)
def test_tc_down_matches_simt(g):
    L, ops, geom, (large, small, w), (lb, sb), (Clp, Csp) = _setup(g)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    bias = torch.randn(Cs, device=DEV)
    ref = torch.empty_like(small)
    ops._conv("mrssm_conv_down", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(ref, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
              L.ptr(bias), ops.RELU)
    wp = ops.tc_pack_weight(w, 0, Csp, Clp)
    out = torch.zeros(n, Hs, Hs, ops.pad16(Cs), device=DEV, dtype=torch.bfloat16)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    ops.tc_conv_down(gp, L.nhwc(lb, Hl, Hl, Clp), L.nhwc(out, Hs, Hs, out.shape[-1]), wp, bias, Cs, act=ops.RELU)
    torch.testing.assert_close(out[..., :Cs].float(), ref, rtol=1e-2, atol=1e-2)
    assert float(out[..., Cs:].float().abs().max() if out.shape[-1] > Cs else 0.0) == 0.0
    # fp32 NCHW output variant
    out32 = torch.empty(n, Cs, Hs, Hs, device=DEV)
    ops.tc_conv_down(gp, L.nhwc(lb, Hl, Hl, Clp), L.nchw(out32, Hs, Hs, Cs), wp, bias, Cs, act=ops.RELU, out_f32=1)
    torch.testing.assert_close(out32.permute(0, 2, 3, 1), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("g", GEOMS)
def test_tc_up_matches_simt(g):
    L, ops, geom, (large, small, w), (lb, sb), (Clp, Csp) = _setup(g)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    if Hl != 2 * (Hs - 1) + k:
        pytest.skip("floor geometry only occurs as a Conv2d (down/dgrad) — covered by the mask test")
    bias = torch.randn(Cl, device=DEV)
    ref = torch.empty_like(large)
    ops._conv("mrssm_conv_up", geom, L.nhwc(ref, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
              L.ptr(bias), ops.RELU)
    wp = ops.tc_pack_weight(w, 1, Csp, Clp)
    out = torch.zeros(n, Hl, Hl, ops.pad16(Cl), device=DEV, dtype=torch.bfloat16)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    ops.tc_conv_up(gp, L.nhwc(out, Hl, Hl, out.shape[-1]), L.nhwc(sb, Hs, Hs, Csp), wp, bias, Cl, act=ops.RELU)
    torch.testing.assert_close(out[..., :Cl].float(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("g", GEOMS)
def test_tc_dgrad_with_mask_matches_simt(g):
    """Conv2d dgrad = up with the ReLU mask of the layer input, including floor geometries (64->31->14)."""
    L, ops, geom, (large, small, w), (lb, sb), (Clp, Csp) = _setup(g)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    ref = torch.empty_like(large)
    ops._conv("mrssm_conv_up", geom, L.nhwc(ref, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
              None, 0, L.ptr(large), ops.RELU)
    wp = ops.tc_pack_weight(w, 1, Csp, Clp)
    out = torch.zeros(n, Hl, Hl, ops.pad16(Cl), device=DEV, dtype=torch.bfloat16)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    ops.tc_conv_up(gp, L.nhwc(out, Hl, Hl, out.shape[-1]), L.nhwc(sb, Hs, Hs, Csp), wp, None, Cl,
                   mask=L.nhwc(lb, Hl, Hl, Clp), mask_mode=ops.RELU)
    torch.testing.assert_close(out[..., :Cl].float(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("g", GEOMS)
def test_tc_wgrad_matches_simt(g):
    L, ops, geom, (large, small, w), (lb, sb), (Clp, Csp) = _setup(g)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    ref = torch.zeros_like(w)
    ops._conv("mrssm_conv_wgrad", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(ref), Cl * k * k, k * k)
    out = torch.zeros_like(w)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    ops.tc_conv_wgrad(gp, L.nhwc(lb, Hl, Hl, Clp), L.nhwc(sb, Hs, Hs, Csp), L.ptr(out), Cl * k * k, k * k, Cs, Cl)
    scale = float(ref.abs().max())
    assert float((out - ref).abs().max()) <= 2e-3 * scale + 1e-4


def test_tc_up_1x1_as_dense():
    """ConvTranspose2d on a 1x1 input (decoder layer 0) lowered to a dense tcgen05 GEMM."""
    from mrssm_b200 import _lib as L, ops
    n, Cs, Cl, k = 200, 1024, 128, 5
    gen = torch.Generator(device=DEV).manual_seed(1)
    x = _bf16_round(torch.randn(n, Cs, device=DEV, generator=gen))
    w = _bf16_round(torch.randn(Cs, Cl, k, k, device=DEV, generator=gen) / Cs ** 0.5)
    bias = torch.randn(Cl, device=DEV)
    ref = torch.empty(n, k, k, Cl, device=DEV)
    ops._conv("mrssm_conv_up", (n, k, k, Cl, 1, 1, Cs, k), L.nhwc(ref, k, k, Cl), L.nhwc(x, 1, 1, Cs), L.ptr(w), Cl * k * k, k * k,
              L.ptr(bias), ops.RELU)
    wp = ops.tc_pack_weight(w, 2, Cs, Cl)
    xb = ops.tc_to_bf16(L.nhwc(x, 1, 1, Cs), n, 1, 1, Cs, DEV)
    out = torch.zeros(n, k * k * Cl, device=DEV, dtype=torch.bfloat16)
    ops.tc_conv_down((n, 1, 1, Cs, 1, 1, k * k * Cl, 1), L.nhwc(xb, 1, 1, Cs), L.nhwc(out, 1, 1, k * k * Cl), wp, bias,
                     k * k * Cl, act=ops.RELU, bias_mod=Cl)
    torch.testing.assert_close(out.float().reshape(n, k, k, Cl), ref, rtol=1e-2, atol=1e-2)


# ---- plane kernels (csrc/conv_plane.cu): TMA-staged planes + shifted descriptors, every spatial geometry -------------
# (down needs 4*Cl/8 planes of >= 128 pixels resident: Cl <= 128)
PL_GEOMS = [g for g in GEOMS if g[5] >= 2 and g[3] > 1 and g[2] <= 128] + [(30, 6, 128, 2, 256, 4), (5, 13, 64, 5, 128, 5),
                                                                         (37, 14, 64, 6, 128, 4), (2, 14, 128, 6, 256, 4)]
# (layout of `large`, layout of `small`): the production pairing and the all-NHWC one (element-wise TMA maps)
PL_LAYOUTS = [("parity", "planar"), ("nhwc", "nhwc")]


def export_view(t, layout, n, H, W, Cp):
    """bf16 view buffer -> fp32 [n,H,W,Cp] (test-side inverse of the layouts in mrssm_b200/_lib.py)."""
    if layout == "nhwc":
        return t.view(n, H, W, Cp).float()
    if layout == "planar":
        return t.view(n, Cp // 8, H, W, 8).permute(0, 2, 3, 1, 4).reshape(n, H, W, Cp).float()
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    v = t.view(n, 2, 2, Cp // 8, H2, W2, 8)
    full = torch.zeros(n, 2 * H2, 2 * W2, Cp, device=t.device)
    for py in range(2):
        for px in range(2):
            full[:, py::2, px::2, :] = v[:, py, px].permute(0, 2, 3, 1, 4).reshape(n, H2, W2, Cp).float()
    return full[:, :H, :W, :]


def _pl_setup(g, ll, sl):
    from mrssm_b200 import _lib as L, ops
    n, Hl, Cl, Hs, Cs, k = g
    gen = torch.Generator(device=DEV).manual_seed(hash(g) % 1000)
    large = _bf16_round(torch.randn(n, Hl, Hl, Cl, device=DEV, generator=gen))
    small = _bf16_round(torch.randn(n, Hs, Hs, Cs, device=DEV, generator=gen))
    w = _bf16_round(torch.randn(Cs, Cl, k, k, device=DEV, generator=gen) / (Cl * k * k) ** 0.5)
    Clp, Csp = ops.pad8(Cl), ops.pad16(Cs)
    lb = ops.pl_import(L.nhwc(large, Hl, Hl, Cl), n, Hl, Hl, Cl, Clp, ll, DEV)
    sb = ops.pl_import(L.nhwc(small, Hs, Hs, Cs), n, Hs, Hs, Cs, Csp, sl, DEV)
    torch.testing.assert_close(export_view(lb[0], ll, n, Hl, Hl, Clp)[..., :Cl], large)      # import/export round trip
    torch.testing.assert_close(export_view(sb[0], sl, n, Hs, Hs, Csp)[..., :Cs], small)
    return L, ops, (n, Hl, Hl, Cl, Hs, Hs, Cs, k), (n, Hl, Hl, Clp, Hs, Hs, Csp, k), (large, small, w), (lb, sb)


def _filled(ops, L, n, H, W, Cp, layout):
    t, v = ops.new_act(n, H, W, Cp, layout, DEV)
    t.fill_(7.0)
    return t, v


@pytest.mark.parametrize("ll,sl", PL_LAYOUTS)
@pytest.mark.parametrize("g", PL_GEOMS)
def test_plane_down_matches_simt(g, ll, sl):
    L, ops, geom, gp, (large, small, w), (lb, sb) = _pl_setup(g, ll, sl)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    Clp, Csp = gp[3], gp[6]
    bias = torch.randn(Cs, device=DEV)
    ref = torch.empty_like(small)
    ops._conv("mrssm_conv_down", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(ref, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
              L.ptr(bias), ops.RELU)
    wp = ops.pl_pack_weight(w, ops.DOWN, Csp, Clp)
    for ol in ("parity", "planar", "nhwc"):          # every output layout the epilogue can address
        out = _filled(ops, L, n, Hs, Hs, Csp, ol)
        ops.pl_conv_down(gp, lb[1], out[1], wp, bias, Cs, Csp, act=ops.RELU)
        o = export_view(out[0], ol, n, Hs, Hs, Csp)
        torch.testing.assert_close(o[..., :Cs], ref, rtol=1e-2, atol=1e-2)
        assert float(o[..., Cs:].abs().max() if Csp > Cs else 0.0) == 0.0
    out32 = torch.empty(n, Cs, Hs, Hs, device=DEV)
    ops.pl_conv_down(gp, lb[1], L.nchw(out32, Hs, Hs, Cs), wp, bias, Cs, Csp, act=ops.RELU)
    torch.testing.assert_close(out32.permute(0, 2, 3, 1), ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("ll,sl", PL_LAYOUTS)
@pytest.mark.parametrize("g", PL_GEOMS)
def test_plane_up_and_dgrad_match_simt(g, ll, sl):
    L, ops, geom, gp, (large, small, w), (lb, sb) = _pl_setup(g, ll, sl)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    Clp, Csp = gp[3], gp[6]
    wp = ops.pl_pack_weight(w, ops.UP, Csp, Clp)
    if Hl == 2 * (Hs - 1) + k:      # ConvTranspose2d forward (+bias, ReLU), bf16 and fp32 NCHW outputs
        bias = torch.randn(Cl, device=DEV)
        ref = torch.empty_like(large)
        ops._conv("mrssm_conv_up", geom, L.nhwc(ref, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
                  L.ptr(bias), ops.RELU)
        for ol in ("planar", "parity", "nhwc"):
            out = _filled(ops, L, n, Hl, Hl, Clp, ol)
            ops.pl_conv_up(gp, out[1], sb[1], wp, bias, Cl, Clp, act=ops.RELU)
            torch.testing.assert_close(export_view(out[0], ol, n, Hl, Hl, Clp)[..., :Cl], ref, rtol=1e-2, atol=1e-2)
        out32 = torch.empty(n, Cl, Hl, Hl, device=DEV)
        ops.pl_conv_up(gp, L.nchw(out32, Hl, Hl, Cl), sb[1], wp, bias, Cl, Clp, act=ops.RELU)
        torch.testing.assert_close(out32.permute(0, 2, 3, 1), ref, rtol=2e-3, atol=2e-3)
    # Conv2d dgrad with the ReLU mask of the layer input (floor geometries included); mask in the layout of `large`
    ref = torch.empty_like(large)
    ops._conv("mrssm_conv_up", geom, L.nhwc(ref, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
              None, 0, L.ptr(large), ops.RELU)
    out = _filled(ops, L, n, Hl, Hl, Clp, "planar")
    ops.pl_conv_up(gp, out[1], sb[1], wp, None, Cl, Clp, mask=lb[1], mask_mode=ops.RELU)
    torch.testing.assert_close(export_view(out[0], "planar", n, Hl, Hl, Clp)[..., :Cl], ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("ll,sl", PL_LAYOUTS)
@pytest.mark.parametrize("g", PL_GEOMS)
def test_plane_wgrad_and_colsum_match_simt(g, ll, sl):
    L, ops, geom, gp, (large, small, w), (lb, sb) = _pl_setup(g, ll, sl)
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    Clp, Csp = gp[3], gp[6]
    ref = torch.zeros_like(w)
    ops._conv("mrssm_conv_wgrad", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(ref), Cl * k * k, k * k)
    out = torch.zeros_like(w)
    ops.pl_conv_wgrad(gp, lb[1], sb[1], L.ptr(out), Cl * k * k, k * k, Cs, Cl)
    scale = float(ref.abs().max())
    assert float((out - ref).abs().max()) <= 2e-3 * scale + 1e-4
    # the same kernel with the bias gradient fused: per-channel sums of `small` (Conv2d) / of `large` (ConvTranspose2d, exact
    # geometries: every pixel of `large` is inside the tile) from the operand tile in shared memory
    for frm, x, Cv in ((1, small, Cs), (2, large, Cl)):
        if frm == 2 and Hl != 2 * (Hs - 1) + k:
            continue
        out2, db = torch.zeros_like(w), torch.full((Cv,), 0.5, device=DEV)
        ops.pl_conv_wgrad(gp, lb[1], sb[1], L.ptr(out2), Cl * k * k, k * k, Cs, Cl, dbias=db, dbias_from=frm)
        assert float((out2 - ref).abs().max()) <= 2e-3 * scale + 1e-4
        torch.testing.assert_close(db, 0.5 + x.sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-2)
    for (t, v, H, Cp, Cv, x) in ((lb[0], lb[1], Hl, Clp, Cl, large), (sb[0], sb[1], Hs, Csp, Cs, small)):
        acc = torch.zeros(Cv, device=DEV)
        ops.pl_colsum(v, n, H, H, Cp, Cv, acc)
        torch.testing.assert_close(acc, x.sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("g", [(3, 64, 3, 31, 32, 4), (2, 64, 3, 30, 32, 6), (2, 63, 3, 30, 16, 4), (1, 128, 3, 62, 32, 6), (4, 14, 4, 6, 16, 4)])
def test_plane_space_to_depth_source_matches_simt(g):
    """<= 4-channel images travel as their space-to-depth form: import, down (Conv2d fwd / ConvT dgrad), wgrad, bias sums."""
    from mrssm_b200 import _lib as L, ops
    n, Hl, Cl, Hs, Cs, k = g
    gen = torch.Generator(device=DEV).manual_seed(5)
    large = _bf16_round(torch.randn(n, Hl, Hl, Cl, device=DEV, generator=gen))
    small = _bf16_round(torch.randn(n, Hs, Hs, Cs, device=DEV, generator=gen))
    w = _bf16_round(torch.randn(Cs, Cl, k, k, device=DEV, generator=gen) / (Cl * k * k) ** 0.5)
    geom = (n, Hl, Hl, Cl, Hs, Hs, Cs, k)
    Csp = ops.pad16(Cs)
    H2 = (Hl + 1) // 2
    lt, lv = ops.pl_import_s2d(L.nhwc(large, Hl, Hl, Cl), n, Hl, Hl, Cl, DEV)
    s2d = export_view(lt, "planar", n, H2, H2, 16)
    for par in range(4):
        sub = large[:, par >> 1::2, par & 1::2, :]
        torch.testing.assert_close(s2d[:, :sub.shape[1], :sub.shape[2], par * Cl:(par + 1) * Cl], sub)
    assert float(s2d[..., 4 * Cl:].abs().max() if 4 * Cl < 16 else 0.0) == 0.0
    gp = (n, Hl, Hl, 16, Hs, Hs, Csp, k)
    bias = torch.randn(Cs, device=DEV)
    ref = torch.empty_like(small)
    ops._conv("mrssm_conv_down", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(ref, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
              L.ptr(bias), ops.RELU)
    wp = ops.pl_pack_weight(w, ops.DOWN_S2D, Csp, 16, Cl)
    out = _filled(ops, L, n, Hs, Hs, Csp, "parity")
    ops.pl_conv_down(gp, lv, out[1], wp, bias, Cs, Csp, act=ops.RELU, s2d_cq=Cl)
    torch.testing.assert_close(export_view(out[0], "parity", n, Hs, Hs, Csp)[..., :Cs], ref, rtol=1e-2, atol=1e-2)
    sb = ops.pl_import(L.nhwc(small, Hs, Hs, Cs), n, Hs, Hs, Cs, Csp, "planar", DEV)
    refw = torch.zeros_like(w)
    ops._conv("mrssm_conv_wgrad", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(refw), Cl * k * k, k * k)
    outw = torch.zeros_like(w)
    ops.pl_conv_wgrad(gp, lv, sb[1], L.ptr(outw), Cl * k * k, k * k, Cs, Cl, s2d_cq=Cl)
    assert float((outw - refw).abs().max()) <= 2e-3 * float(refw.abs().max()) + 1e-4
    for frm, x, Cv in ((1, small, Cs), (2, large, Cl)):       # fused bias gradient; from the space-to-depth tile the parities fold
        if frm == 2 and Hl != 2 * (Hs - 1) + k:
            continue
        outw2, db = torch.zeros_like(w), torch.zeros(Cv, device=DEV)
        sc = torch.full((1,), 0.5, device=DEV)
        ops.pl_conv_wgrad(gp, lv, sb[1], L.ptr(outw2), Cl * k * k, k * k, Cs, Cl, s2d_cq=Cl, dbias=db, dbias_from=frm, scale=(sc, 3.0))
        assert float((outw2 - 1.5 * refw).abs().max()) <= 3e-3 * float(refw.abs().max()) + 1e-4
        torch.testing.assert_close(db, 1.5 * x.sum(dim=(0, 1, 2)), rtol=1e-3, atol=2e-2)
    acc = torch.zeros(Cl, device=DEV)
    ops.pl_colsum(lv, n, H2, H2, 16, Cl, acc, fold=Cl)
    torch.testing.assert_close(acc, large.sum(dim=(0, 1, 2)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("g", [(3, 64, 3, 30, 32, 6), (2, 128, 3, 62, 32, 6), (5, 14, 4, 6, 16, 4)])
def test_plane_up_fused_mse_matches_separate_ops(g):
    """Last ConvTranspose2d with the reconstruction loss fused in the epilogue: loss value, optional reconstruction and the
    bf16 space-to-depth residual against the separate kernels."""
    from mrssm_b200 import _lib as L, ops
    n, Hl, Cl, Hs, Cs, k = g
    gen = torch.Generator(device=DEV).manual_seed(11)
    small = _bf16_round(torch.randn(n, Hs, Hs, Cs, device=DEV, generator=gen))
    w = _bf16_round(torch.randn(Cs, Cl, k, k, device=DEV, generator=gen) / (Cs * k * k / 4) ** 0.5)
    bias = torch.randn(Cl, device=DEV)
    target = torch.randn(n, Cl, Hl, Hl, device=DEV, generator=gen)
    Csp, Clp = ops.pad16(Cs), ops.pad8(Cl)
    sb = ops.pl_import(L.nhwc(small, Hs, Hs, Cs), n, Hs, Hs, Cs, Csp, "planar", DEV)
    wp = ops.pl_pack_weight(w, ops.UP, Csp, Clp)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    recon_ref = torch.empty(n, Cl, Hl, Hl, device=DEV)
    ops.pl_conv_up(gp, L.nchw(recon_ref, Hl, Hl, Cl), sb[1], wp, bias, Cl, Clp)
    H2 = (Hl + 1) // 2
    for with_recon in (False, True):
        resid = ops.new_act(n, H2, H2, 16, "planar", DEV)
        resid[0].fill_(7.0)
        total = torch.zeros(1, device=DEV)
        recon = torch.zeros(n, Cl, Hl, Hl, device=DEV) if with_recon else None
        ops.pl_conv_up_mse(gp, resid[1], sb[1], wp, bias, Cl, Clp, target, L.nchw(target, Hl, Hl, Cl), total, 1.0 / n,
                           recon_t4=L.nchw(recon, Hl, Hl, Cl) if with_recon else None)
        ref_loss = ((recon_ref - target) ** 2).sum() / n
        torch.testing.assert_close(total[0], ref_loss, rtol=1e-4, atol=1e-4)
        if with_recon:
            torch.testing.assert_close(recon, recon_ref, rtol=0, atol=0)
        r = export_view(resid[0], "planar", n, H2, H2, 16)
        d = (recon_ref - target).permute(0, 2, 3, 1)
        for par in range(4):
            sub = d[:, par >> 1::2, par & 1::2, :]
            torch.testing.assert_close(r[:, :sub.shape[1], :sub.shape[2], par * Cl:(par + 1) * Cl], sub, rtol=1e-2, atol=1e-2)
    # target handed over as the bf16 space-to-depth import of the same image (what a training step does: the decoder's target
    # is the encoder's input): the loss is taken against the image rounded to bf16
    ts, tsv = ops.pl_import_s2d(L.nchw(target, Hl, Hl, Cl), n, Hl, Hl, Cl, DEV)
    resid = ops.new_act(n, H2, H2, 16, "planar", DEV)
    total = torch.zeros(1, device=DEV)
    ops.pl_conv_up_mse(gp, resid[1], sb[1], wp, bias, Cl, Clp, target, L.nchw(target, Hl, Hl, Cl), total, 1.0 / n, target_s2d=tsv)
    tb = _bf16_round(target)
    torch.testing.assert_close(total[0], ((recon_ref - tb) ** 2).sum() / n, rtol=1e-4, atol=1e-4)
    r = export_view(resid[0], "planar", n, H2, H2, 16)
    d = (recon_ref - tb).permute(0, 2, 3, 1)
    for par in range(4):
        sub = d[:, par >> 1::2, par & 1::2, :]
        torch.testing.assert_close(r[:, :sub.shape[1], :sub.shape[2], par * Cl:(par + 1) * Cl], sub, rtol=1e-2, atol=1e-2)


def decode_relu_bits(bits, n, H, W, Cp):
    """[n*H*W*Cp/8] sign bytes (include/mrssm_b200.h: relu_bits_out) -> bool [n,H,W,Cp]: within a byte, bit (3 - j) is channel
    2j and bit (7 - j) channel 2j + 1 of the 8-channel chunk."""
    b = bits[: n * H * W * (Cp // 8)].view(n, H, W, Cp // 8).to(torch.int32)
    out = torch.zeros(n, H, W, Cp // 8, 8, dtype=torch.bool, device=bits.device)
    for j in range(4):
        out[..., 2 * j] = ((b >> (3 - j)) & 1).bool()
        out[..., 2 * j + 1] = ((b >> (7 - j)) & 1).bool()
    return out.reshape(n, H, W, Cp)


@pytest.mark.parametrize("g", PL_GEOMS)
def test_plane_relu_sign_bits_round_trip(g):
    """Forward epilogues emit 1 bit per element (output > 0); the dgrad epilogues that take those bits instead of the bf16
    activation must give bit-identical gradients.  Both directions: Conv2d (down fwd -> up dgrad) and ConvTranspose2d
    (up fwd -> down dgrad)."""
    L, ops, geom, gp, (large, small, w), (lb, sb) = _pl_setup(g, "parity", "planar")
    n, Hl, _, Cl, Hs, _, Cs, k = geom
    Clp, Csp = gp[3], gp[6]
    # Conv2d forward: y = relu(conv(large) + b) with sign bits; then "dgrad of the layer above" masked by y
    bias = torch.randn(Cs, device=DEV)
    out = _filled(ops, L, n, Hs, Hs, Csp, "parity")
    bits = ops.new_relu_bits(n, Hs, Hs, Csp, DEV)
    bits.fill_(0xAA)
    ops.pl_conv_down(gp, lb[1], out[1], ops.pl_pack_weight(w, ops.DOWN, Csp, Clp), bias, Cs, Csp, act=ops.RELU, bits_out=bits)
    y = export_view(out[0], "parity", n, Hs, Hs, Csp)
    assert torch.equal(decode_relu_bits(bits, n, Hs, Hs, Csp), y > 0)
    # a masked op writing an [n,Hs,Hs,Csp] tensor: ConvTranspose2d dgrad == `down` with the mask of its own output's ReLU
    wpd = ops.pl_pack_weight(w, ops.DOWN, Csp, Clp)
    ref = _filled(ops, L, n, Hs, Hs, Csp, "nhwc")
    ops.pl_conv_down(gp, lb[1], ref[1], wpd, None, Cs, Csp, mask=out[1], mask_mode=ops.RELU)
    got = _filled(ops, L, n, Hs, Hs, Csp, "nhwc")
    ops.pl_conv_down(gp, lb[1], got[1], wpd, None, Cs, Csp, bits_in=bits)
    assert torch.equal(ref[0], got[0])
    torch.testing.assert_close(export_view(got[0], "nhwc", n, Hs, Hs, Csp)[..., :Cs],
                               _ref_masked_down(large, w, y[..., :Cs]), rtol=1e-2, atol=1e-2)
    if Hl != 2 * (Hs - 1) + k or Clp < 16:       # (3-channel outputs take the generic epilogue: no sign bits)
        return
    # ConvTranspose2d forward with sign bits, then Conv2d dgrad (`up`) masked by them
    bias_l = torch.randn(Cl, device=DEV)
    wpu = ops.pl_pack_weight(w, ops.UP, Csp, Clp)
    outl = _filled(ops, L, n, Hl, Hl, Clp, "planar")
    bits_l = ops.new_relu_bits(n, Hl, Hl, Clp, DEV)
    ops.pl_conv_up(gp, outl[1], sb[1], wpu, bias_l, Cl, Clp, act=ops.RELU, bits_out=bits_l)
    yl = export_view(outl[0], "planar", n, Hl, Hl, Clp)
    assert torch.equal(decode_relu_bits(bits_l, n, Hl, Hl, Clp), yl > 0)
    ref = _filled(ops, L, n, Hl, Hl, Clp, "planar")
    ops.pl_conv_up(gp, ref[1], sb[1], wpu, None, Cl, Clp, mask=outl[1], mask_mode=ops.RELU)
    got = _filled(ops, L, n, Hl, Hl, Clp, "planar")
    ops.pl_conv_up(gp, got[1], sb[1], wpu, None, Cl, Clp, bits_in=bits_l)
    assert torch.equal(ref[0], got[0])


def _ref_masked_down(large, w, y):
    import torch.nn.functional as F
    d = F.conv2d(large.permute(0, 3, 1, 2), w, None, stride=2).permute(0, 2, 3, 1)
    return d * (y > 0)


@pytest.mark.parametrize("shape", [((3,), (128, 128, 128), True, 2), ((200, 30), (128, 128, 3), False, 2), ((1024,), (512,), True, 1)])
def test_mlp_tensor_core_matches_exact_kernels(shape):
    """MlpTCFn (bf16 tcgen05 GEMMs per Linear) against MlpFn (exact fp32 CUDA-core kernels): SymbolicEncoder, DenseDecoder and
    the fc branch of the conv encoder — forward, input gradients, every weight / bias gradient."""
    from mrssm_b200 import ops
    in_dims, widths, final_act, act = shape
    M = 3000
    gen = torch.Generator(device=DEV).manual_seed(3)
    parts = [torch.randn(M, k, device=DEV, generator=gen).requires_grad_(len(in_dims) > 1) for k in in_dims]
    params, k = [], sum(in_dims)
    for n in widths:
        params += [torch.nn.Parameter(torch.randn(n, k, device=DEV, generator=gen) / k ** 0.5), torch.nn.Parameter(torch.randn(n, device=DEV, generator=gen) * 0.1)]
        k = n
    gout = torch.randn(M, widths[-1], device=DEV, generator=gen)
    res = {}
    for name, fn in (("exact", ops.MlpFn), ("tc", ops.MlpTCFn)):
        for p in params + parts:
            p.grad = None
        y = fn.apply(act, final_act, len(parts), *parts, *params)
        y.backward(gout)
        res[name] = (y.detach().clone(), [p.grad.clone() for p in params], [p.grad.clone() for p in parts if p.requires_grad])
    (y0, gw0, gx0), (y1, gw1, gx1) = res["exact"], res["tc"]
    torch.testing.assert_close(y1, y0, rtol=2e-2, atol=2e-2)
    # a ReLU whose pre-activation is ~0 may flip with bf16 operands: Frobenius, not max.  With a ReLU on the LAST layer the flipped
    # units (0.07 % of them at K = 1024) carry the full upstream gradient: an fp32 emulation of the bf16 operand rounding alone gives
    # 3.7e-2 for that shape (2.3e-3 with the masks held fixed), so it is bounded at 5e-2.
    tol = 5e-2 if (final_act and len(widths) == 1) else 2e-2
    for a, b in zip(gw1 + gx1, gw0 + gx0):
        assert float((a - b).norm()) <= tol * float(b.norm()) + 1e-6, (float((a - b).norm()), float(b.norm()))
