"""-m gpu: the exact fp32 NCHW kernels behind the shipped YAML's sound modality and BatchNorm image stacks (csrc/generic_nchw.cu,
through the C ABI and the autograd Functions of mrssm_b200/ops.py) against torch.nn.functional on the same inputs — the layers of
SoundEncoder_v2 / SoundDecoder_v2 (reference encoder.py:661-721, observation_model.py:420-472) and of the BatchNorm variants of
ImageEncoder / ImageDecoder (encoder.py:324-337, observation_model.py:75-86).  Tolerance: rtol 1e-4 on values and gradients
(fp32 accumulation-order differences only), atol scaled to the tensor."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    """torch's CUDA convolutions default to TF32 (10-bit mantissa products): the reference side must be real fp32."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _close(a, b, rtol=1e-4):
    scale = max(1.0, float(b.abs().max()))
    torch.testing.assert_close(a, b, rtol=rtol, atol=2e-5 * scale)


# (N, Cin, H, W, Cout, kernel, stride, padding): every Conv2d of the sound encoder / decoder + the BatchNorm image encoder's
CONVS = [
    (3, 1, 128, 20, 128, (3, 9), (1, 1), (1, 4)),       # down_sample_1
    (2, 64, 128, 20, 256, (4, 8), (2, 2), (1, 3)),      # down_sample_2
    (2, 128, 64, 10, 512, (4, 8), (2, 2), (1, 3)),      # down_sample_3
    (2, 256, 32, 5, 512, (3, 4), (1, 1), (1, 1)),       # down_sample_4
    (2, 64, 128, 20, 1, (7, 7), (1, 1), (3, 3)),        # SoundDecoder_v2.out
    (5, 3, 64, 64, 32, (4, 4), (2, 2), (0, 0)),         # image encoder, BatchNorm variant: conv.0
    (5, 128, 6, 6, 256, (4, 4), (2, 2), (0, 0)),        # conv.9
    (3, 7, 9, 11, 5, (2, 3), (3, 2), (2, 0)),           # odd everything
]


@pytest.mark.parametrize("g", CONVS)
def test_conv2d_matches_torch(g):
    from mrssm_b200 import ops
    N, Cin, H, W, Cout, k, s, p = g
    gen = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(N, Cin, H, W, device=DEV, generator=gen, requires_grad=True)
    w = (torch.randn(Cout, Cin, *k, device=DEV, generator=gen) / (Cin * k[0] * k[1]) ** 0.5).requires_grad_(True)
    ref = F.conv2d(x, w, None, stride=s, padding=p)
    go = torch.randn(ref.shape, device=DEV, generator=gen)
    gx_ref, gw_ref = torch.autograd.grad(ref, (x, w), go)
    w.grad = None
    y = ops.conv2d_nobias(x, w, s, p)
    _close(y, ref.detach())
    (gx,) = torch.autograd.grad(y, (x,), go)
    _close(gx, gx_ref)
    _close(w.grad, gw_ref)
    # the weight gradient ACCUMULATES (it lands in the optimiser's flat buffer): a second backward doubles it
    y2 = ops.conv2d_nobias(x, w, s, p)
    torch.autograd.grad(y2, (x,), go)
    _close(w.grad, 2 * gw_ref)


# (N, Cin, Hs, Ws, Cout, kernel, stride, padding): every ConvTranspose2d of the sound decoder + the BatchNorm image decoder's
CONVTS = [
    (2, 256, 32, 4, 512, (3, 4), (1, 1), (1, 1)),       # up_sample_0
    (2, 256, 32, 5, 256, (4, 4), (2, 2), (1, 1)),       # up_sample_1
    (2, 128, 64, 10, 128, (4, 4), (2, 2), (1, 1)),      # up_sample_2
    (6, 1024, 1, 1, 128, (5, 5), (2, 2), (0, 0)),       # image decoder conv.0 on the 1x1 map
    (3, 64, 13, 13, 32, (6, 6), (2, 2), (0, 0)),        # conv.6
    (3, 32, 30, 30, 3, (6, 6), (2, 2), (0, 0)),         # conv.9 (with bias, added separately)
]


@pytest.mark.parametrize("g", CONVTS)
def test_conv_transpose2d_matches_torch(g):
    from mrssm_b200 import ops
    N, Cin, Hs, Ws, Cout, k, s, p = g
    gen = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(N, Cin, Hs, Ws, device=DEV, generator=gen, requires_grad=True)
    w = (torch.randn(Cin, Cout, *k, device=DEV, generator=gen) / (Cin * k[0] * k[1]) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device=DEV, generator=gen).requires_grad_(True)
    ref = F.conv_transpose2d(x, w, b, stride=s, padding=p)
    go = torch.randn(ref.shape, device=DEV, generator=gen)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(ref, (x, w, b), go)
    w.grad = b.grad = None
    y = ops.add_channel_bias(ops.conv_transpose2d_nobias(x, w, s, p), b)
    _close(y, ref.detach())
    (gx,) = torch.autograd.grad(y, (x,), go)
    _close(gx, gx_ref)
    _close(w.grad, gw_ref)
    _close(b.grad, gb_ref)


def test_conv1d_k1_matches_torch():
    from mrssm_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(3)
    for N, Cin, Lr, Cout in ((3, 8192, 4, 128), (5, 1152, 1, 4096)):       # down_conversion, a slice of up_conversion
        x = torch.randn(N, Cin, Lr, device=DEV, generator=gen, requires_grad=True)
        w = (torch.randn(Cout, Cin, 1, device=DEV, generator=gen) / Cin ** 0.5).requires_grad_(True)
        ref = F.conv1d(x, w)
        go = torch.randn(ref.shape, device=DEV, generator=gen)
        gx_ref, gw_ref = torch.autograd.grad(ref, (x, w), go)
        w.grad = None
        y = ops.conv1d_k1(x, w)
        _close(y, ref.detach())
        (gx,) = torch.autograd.grad(y, (x,), go)
        _close(gx, gx_ref)
        _close(w.grad, gw_ref)


@pytest.mark.parametrize("shape,relu", [((7, 32, 31, 31), True), ((4, 256, 2, 2), True), ((5, 64, 13, 13), False)])
def test_batch_norm_matches_torch(shape, relu):
    from mrssm_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(4)
    Cn = shape[1]
    bn_ref, bn = nn.BatchNorm2d(Cn).to(DEV), nn.BatchNorm2d(Cn).to(DEV)
    with torch.no_grad():
        bn_ref.weight.copy_(torch.rand(Cn, device=DEV, generator=gen) + 0.5)
        bn_ref.bias.copy_(torch.randn(Cn, device=DEV, generator=gen))
    bn.load_state_dict(bn_ref.state_dict())
    for step in range(2):                                             # two steps: the running statistics move twice
        x = (torch.randn(shape, device=DEV, generator=gen) * 2 + 0.7).requires_grad_(True)
        ref = bn_ref(x)
        ref = F.relu(ref) if relu else ref
        go = torch.randn(shape, device=DEV, generator=gen)
        gx_ref, gg_ref, gb_ref = torch.autograd.grad(ref, (x, bn_ref.weight, bn_ref.bias), go)
        bn.weight.grad = bn.bias.grad = None
        y = ops.batch_norm(x, bn, relu=relu)
        _close(y, ref.detach())
        (gx,) = torch.autograd.grad(y, (x,), go)
        _close(gx, gx_ref)
        _close(bn.weight.grad, gg_ref)
        _close(bn.bias.grad, gb_ref)
        _close(bn.running_mean, bn_ref.running_mean)
        _close(bn.running_var, bn_ref.running_var)
        assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked) == step + 1
    bn.eval(), bn_ref.eval()
    x = torch.randn(shape, device=DEV, generator=gen).requires_grad_(True)
    ref = F.relu(bn_ref(x)) if relu else bn_ref(x)
    go = torch.randn(shape, device=DEV, generator=gen)
    (gx_ref,) = torch.autograd.grad(ref, (x,), go)
    y = ops.batch_norm(x, bn, relu=relu)
    _close(y, ref.detach())
    (gx,) = torch.autograd.grad(y, (x,), go)
    _close(gx, gx_ref)


@pytest.mark.parametrize("shape,tracked", [((3, 256, 64, 10), True), ((2, 512, 32, 4), True), ((4, 128, 4), False)])
def test_instance_norm_matches_torch(shape, tracked):
    from mrssm_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(5)
    Cn = shape[1]
    mk = (lambda: nn.InstanceNorm2d(Cn, affine=True, track_running_stats=True)) if tracked else (lambda: nn.InstanceNorm1d(Cn, affine=True))
    m_ref, m = mk().to(DEV), mk().to(DEV)
    with torch.no_grad():
        m_ref.weight.copy_(torch.rand(Cn, device=DEV, generator=gen) + 0.5)
        m_ref.bias.copy_(torch.randn(Cn, device=DEV, generator=gen))
    m.load_state_dict(m_ref.state_dict())
    for mode in (["train", "train", "eval"] if tracked else ["train", "eval"]):
        (m.train(), m_ref.train()) if mode == "train" else (m.eval(), m_ref.eval())
        x = (torch.randn(shape, device=DEV, generator=gen) * 1.5 - 0.3).requires_grad_(True)
        ref = m_ref(x)
        go = torch.randn(shape, device=DEV, generator=gen)
        gx_ref, gg_ref, gb_ref = torch.autograd.grad(ref, (x, m_ref.weight, m_ref.bias), go)
        m.weight.grad = m.bias.grad = None
        y = ops.instance_norm(x, m)
        _close(y, ref.detach())
        (gx,) = torch.autograd.grad(y, (x,), go)
        _close(gx, gx_ref)
        _close(m.weight.grad, gg_ref)
        _close(m.bias.grad, gb_ref)
        if tracked:
            _close(m.running_mean, m_ref.running_mean)
            _close(m.running_var, m_ref.running_var)
            assert int(m.num_batches_tracked) == int(m_ref.num_batches_tracked)


def test_glu_matches_torch():
    from mrssm_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(6)
    for shape in ((3, 128, 128, 20), (5, 128, 4), (2, 6, 3, 1)):
        x = torch.randn(shape, device=DEV, generator=gen, requires_grad=True)
        ref = F.glu(x, dim=1)
        go = torch.randn(ref.shape, device=DEV, generator=gen)
        (gx_ref,) = torch.autograd.grad(ref, (x,), go)
        y = ops.GluFn.apply(x)
        _close(y, ref.detach())
        (gx,) = torch.autograd.grad(y, (x,), go)
        _close(gx, gx_ref)


def test_sound_modules_match_torch_modules_built_from_the_same_layers():
    """SoundEncoder_v2 / SoundDecoder_v2 mirrors (their nn layers are parameter containers) against calling those very nn layers."""
    from utils.models.encoder import SoundEncoder_v2
    from utils.models.observation_model import SoundDecoder_v2
    torch.manual_seed(0)
    enc = SoundEncoder_v2(embbed_size=256).to(DEV)
    dec = SoundDecoder_v2(belief_size=40, state_size=24).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(3, 128, 20, device=DEV, generator=gen)
    ref = x.unsqueeze(1)
    for st in (enc.down_sample_1, enc.down_sample_2, enc.down_sample_3, enc.down_sample_4):
        ref = st(ref)
    ref = enc.down_conversion(ref.contiguous().view(-1, enc.conversion_channels, 4)).contiguous().view(-1, 256)
    sd = {k: v.clone() for k, v in enc.state_dict().items()}
    ref_stats = {k: v.clone() for k, v in enc.state_dict().items() if "running" in k}
    enc.load_state_dict(sd)                                                          # (the torch pass above moved the statistics)
    for k, v in enc.state_dict().items():
        if "running_mean" in k:
            v.zero_()
        elif "running_var" in k:
            v.fill_(1.0)
    out = enc(x)
    _close(out, ref.detach(), rtol=2e-4)
    for k, v in enc.state_dict().items():
        if "running" in k:
            _close(v, ref_stats[k])
    h, s = torch.randn(2, 3, 40, device=DEV, generator=gen), torch.randn(2, 3, 24, device=DEV, generator=gen)
    y = dec(h, s)["loc"]                                                             # called as every caller does: (beliefs, states)
    z = torch.cat([s.reshape(6, -1, 1), h.reshape(6, -1, 1)], dim=1)
    z = dec.up_conversion(z).view(-1, 256, 32, 4)
    for st in (dec.up_sample_0, dec.up_sample_1, dec.up_sample_2):
        z = st(z)
    z = dec.out(z).squeeze(1).reshape(2, 3, 128, 20)
    assert y.shape == (2, 3, 128, 20)
    _close(y, z.detach(), rtol=2e-4)


def _fro(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("g,transposed", [(c, False) for c in CONVS[1:4]] + [((4, 32, 31, 31, 64, (4, 4), (2, 2), (0, 0)), False),
                                                                            ((70, 8192, 4, 1, 128, (1, 1), (1, 1), (0, 0)), False)]
                         + [(c, True) for c in CONVTS[0:3]] + [((70, 1024, 1, 1, 128, (5, 5), (2, 2), (0, 0)), True)])
def test_tensor_core_route_matches_exact_kernels(g, transposed):
    """GConvTCFn (bf16 NHWC staging, explicit im2col rows, tcgen05 GEMMs; bf16 mode's route for the 128 .. 512-channel sound layers
    and the inner layers of the BatchNorm image stacks) against GConvFn (exact fp32) on the same inputs: output, input gradient and
    weight gradient within 1e-2 relative Frobenius error (bf16 operands, fp32 accumulation; measured 2e-3 .. 4e-3)."""
    from mrssm_b200 import ops
    N, C0, H0, W0, C1, k, s, p = g
    gen = torch.Generator(device=DEV).manual_seed(8)
    x = torch.randn(N, C0, H0, W0, device=DEV, generator=gen, requires_grad=True)
    wshape = (C0, C1, *k) if transposed else (C1, C0, *k)
    w = (torch.randn(wshape, device=DEV, generator=gen) / (C0 * k[0] * k[1]) ** 0.5).requires_grad_(True)
    res = []
    go = None
    for fn in (ops.GConvFn, ops.GConvTCFn):
        w.grad = None
        y = fn.apply(x, w, s, p, transposed)
        if go is None:
            go = torch.randn(y.shape, device=DEV, generator=gen)
        (gx,) = torch.autograd.grad(y, (x,), go)
        res.append((y.detach(), gx, w.grad.clone()))
    for name, a, b in zip(("y", "gx", "gw"), res[1], res[0]):
        assert a.shape == b.shape and torch.isfinite(a).all(), name
        assert _fro(a, b) <= 1e-2, (name, _fro(a, b))


def test_tensor_core_route_chunks_over_images():
    from mrssm_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(9, 64, 32, 10, device=DEV, generator=gen, requires_grad=True)
    w = (torch.randn(128, 64, 4, 8, device=DEV, generator=gen) / 45.0).requires_grad_(True)
    outs = []
    for col_bytes in (2 << 30, 3 * 16 * 5 * 2048 * 2):                 # everything in one chunk; 3 images per chunk (3 chunks)
        old, ops._GC_COL_BYTES = ops._GC_COL_BYTES, col_bytes
        try:
            w.grad = None
            y = ops.GConvTCFn.apply(x, w, (2, 2), (1, 3), False)
            (gx,) = torch.autograd.grad(y, (x,), torch.ones_like(y))
            outs.append((y.detach(), gx, w.grad.clone()))
        finally:
            ops._GC_COL_BYTES = old
    for a, b in zip(*outs):
        assert _fro(a, b) <= 1e-5                                        # (atomics order in the weight gradient only)


def _guarded(numel, dtype=torch.float32, pad=4096):
    """A buffer of `numel` elements between two sentinel-filled guard zones: (whole, view, check())."""
    whole = torch.full((numel + 2 * pad,), 12345.0, device=DEV, dtype=dtype)
    view = whole[pad:pad + numel]

    def check():
        assert bool((whole[:pad] == 12345.0).all()) and bool((whole[pad + numel:] == 12345.0).all()), "write outside the output buffer"
    return whole, view, check


@pytest.mark.parametrize("g", [(3, 7, 9, 11, 5, (2, 3), (3, 2), (2, 0)), (2, 64, 20, 10, 1, (7, 7), (1, 1), (3, 3)),
                               (5, 3, 30, 30, 32, (6, 6), (2, 2), (0, 0)), (2, 40, 13, 9, 70, (3, 4), (1, 2), (1, 1))])
def test_general_convolution_kernels_stay_inside_their_outputs(g):
    """Through the C ABI with guard zones around every output (compute-sanitizer is not available on the pool): the three roles of
    the general convolution — 64x64, 256x16, 16x256 and direct small-channel tiles, ragged sizes — write only what they own."""
    import ctypes as C
    from mrssm_b200 import _lib as L
    N, Cin, H, W, Cout, k, s, p = g
    Ho, Wo = (H + 2 * p[0] - k[0]) // s[0] + 1, (W + 2 * p[1] - k[1]) // s[1] + 1
    gen = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(N, Cin, H, W, device=DEV, generator=gen)
    w = torch.randn(Cout, Cin, *k, device=DEV, generator=gen)
    dy = torch.randn(N, Cout, Ho, Wo, device=DEV, generator=gen)
    _, y, chk_y = _guarded(N * Cout * Ho * Wo)
    _, dx, chk_dx = _guarded(x.numel())
    _, dw, chk_dw = _guarded(w.numel())
    dw.zero_()

    def args(**kw):
        return L.GConvArgs(N, Cin, H, W, Cout, k[0], k[1], s[0], s[1], p[0], p[1], Ho, Wo, kw.get("x"), kw.get("w"), kw.get("y"), kw.get("dx"),
                           kw.get("dw"))
    L.call("mrssm_gconv_fwd", C.byref(args(x=L.ptr(x), w=L.ptr(w), y=y.data_ptr())))
    L.call("mrssm_gconv_dgrad", C.byref(args(w=L.ptr(w), y=L.ptr(dy), dx=dx.data_ptr())))
    L.call("mrssm_gconv_wgrad", C.byref(args(x=L.ptr(x), y=L.ptr(dy), dw=dw.data_ptr())))
    torch.cuda.synchronize()
    chk_y(), chk_dx(), chk_dw()
    ref = F.conv2d(x, w, None, stride=s, padding=p)
    _close(y.view_as(ref), ref)


def test_tensor_core_staging_kernels_stay_inside_their_outputs():
    import ctypes as C
    from mrssm_b200 import _lib as L
    N, Cin, H, W, Cout, k, s, p = 3, 24, 11, 7, 16, (3, 4), (2, 1), (1, 2)
    Ho, Wo = (H + 2 * p[0] - k[0]) // s[0] + 1, (W + 2 * p[1] - k[1]) // s[1] + 1
    K = Cin * k[0] * k[1]
    gen = torch.Generator(device=DEV).manual_seed(12)
    x = torch.randn(N, Cin, H, W, device=DEV, generator=gen)
    a = L.GConvArgs(N, Cin, H, W, Cout, k[0], k[1], s[0], s[1], p[0], p[1], Ho, Wo, None, None, None, None, None)
    _, xh, chk_xh = _guarded(N * H * W * Cin, torch.bfloat16)
    L.call("mrssm_nchw_to_nhwc_bf16", L.ptr(x), N, Cin, H * W, xh.data_ptr())
    _, col, chk_col = _guarded(N * Ho * Wo * K, torch.bfloat16)
    L.call("mrssm_im2col_nhwc", C.byref(a), xh.data_ptr(), col.data_ptr())
    _, dxh, chk_dxh = _guarded(N * H * W * Cin)
    L.call("mrssm_col2im_nhwc", C.byref(a), col.data_ptr(), dxh.data_ptr())
    _, back, chk_back = _guarded(N * Cin * H * W)
    L.call("mrssm_nhwc_to_nchw_f32", dxh.data_ptr(), N, Cin, H * W, back.data_ptr())
    torch.cuda.synchronize()
    chk_xh(), chk_col(), chk_dxh(), chk_back()
    # im2col followed by col2im multiplies every input pixel by the number of windows that cover it
    cover = F.conv_transpose2d(torch.ones(1, 1, Ho, Wo, device=DEV), torch.ones(1, 1, *k, device=DEV), stride=s,
                               padding=p, output_padding=(H - ((Ho - 1) * s[0] - 2 * p[0] + k[0]), W - ((Wo - 1) * s[1] - 2 * p[1] + k[1])))
    _close(back.view(N, Cin, H, W), x.bfloat16().float() * cover, rtol=1e-5)
