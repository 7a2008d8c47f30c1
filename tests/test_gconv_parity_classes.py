"""CPU: the stride-parity decomposition of the general convolution's input gradient (csrc/generic_nchw.cu, MODE 1) restated in
numpy / torch and checked against torch's own conv_transpose2d — the index arithmetic the kernel's grid.z classes implement:
class (rh, rw) owns the input positions with (h + PH) % SH == rh, (w + PW) % SW == rw and only the taps kh = rh + i*SH,
kw = rw + j*SW reach them, with output coordinates ho = (h + PH - rh) / SH - i, wo likewise."""
import pytest
import torch
import torch.nn.functional as F


def dgrad_by_parity_classes(dy, w, H, W, stride, padding):
    N, Cout, Ho, Wo = dy.shape
    _, Cin, KH, KW = w.shape
    SH, SW = stride
    PH, PW = padding
    dx = torch.full((N, Cin, H, W), float("nan"), dtype=dy.dtype)
    taps_used = 0
    for rh in range(SH):
        for rw in range(SW):
            h0, w0 = (rh - PH) % SH, (rw - PW) % SW
            hs, ws = list(range(h0, H, SH)), list(range(w0, W, SW))
            nkh = (KH - rh + SH - 1) // SH if rh < KH else 0
            nkw = (KW - rw + SW - 1) // SW if rw < KW else 0
            if not hs or not ws:
                continue
            acc = torch.zeros(N, Cin, len(hs), len(ws), dtype=dy.dtype)
            for i in range(nkh):
                for j in range(nkw):
                    kh, kw = rh + i * SH, rw + j * SW
                    taps_used += 1
                    for a, h in enumerate(hs):
                        ho = (h + PH - rh) // SH - i
                        assert (h + PH - kh) % SH == 0                      # the class's taps always divide
                        if not 0 <= ho < Ho:
                            continue
                        for b, x in enumerate(ws):
                            wo = (x + PW - rw) // SW - j
                            if 0 <= wo < Wo:
                                acc[:, :, a, b] += torch.einsum("no,oc->nc", dy[:, :, ho, wo], w[:, :, kh, kw])
            dx[:, :, h0::SH, w0::SW] = acc
    return dx, taps_used


@pytest.mark.parametrize("g", [(2, 3, 9, 11, 4, (2, 3), (3, 2), (2, 0)), (1, 2, 12, 10, 3, (4, 8), (2, 2), (1, 3)),
                               (2, 2, 7, 6, 2, (3, 4), (1, 1), (1, 1)), (1, 3, 8, 8, 2, (6, 6), (2, 2), (0, 0)), (1, 1, 5, 7, 2, (1, 1), (2, 3), (0, 0))])
def test_parity_classes_reproduce_conv_transpose(g):
    N, Cin, H, W, Cout, k, s, p = g
    Ho, Wo = (H + 2 * p[0] - k[0]) // s[0] + 1, (W + 2 * p[1] - k[1]) // s[1] + 1
    gen = torch.Generator().manual_seed(5)
    dy = torch.randn(N, Cout, Ho, Wo, generator=gen, dtype=torch.float64)
    w = torch.randn(Cout, Cin, *k, generator=gen, dtype=torch.float64)
    out_pad = (H - ((Ho - 1) * s[0] - 2 * p[0] + k[0]), W - ((Wo - 1) * s[1] - 2 * p[1] + k[1]))
    ref = F.conv_transpose2d(dy, w, None, stride=s, padding=p, output_padding=out_pad)
    dx, taps_used = dgrad_by_parity_classes(dy, w, H, W, s, p)
    assert not torch.isnan(dx).any()                                       # the classes tile every input position exactly once
    torch.testing.assert_close(dx, ref, rtol=1e-12, atol=1e-12)
    assert taps_used == k[0] * k[1]                                        # and every tap belongs to exactly one class
