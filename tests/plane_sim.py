"""CPU model of the index arithmetic of csrc/conv_plane.cu (test infrastructure, not a product path).

It replays, in numpy, what the sm_100a kernels do with their shared-memory planes: the TMA box loads (parity
split, zero out-of-bounds fill), the K-step -> (plane, pixel shift) table, the packed weight layout and the
epilogue's position -> pixel decode, driven by the REAL tiling plan returned by the library's host-side
planner (mrssm_pl_describe).  Garbage shared memory is modelled as NaN, so a valid output that touched
memory the kernel never filled shows up as NaN.  Used by tests/test_plane_sim.py against torch convolutions.
"""
import ctypes as C
import re

import numpy as np


def plan(L, geom, op, n_out_pad=0, s2d_cq=0):
    """geom = (n, Hl, Wl, Clp, Hs, Ws, Csp, k) with PADDED channel counts; op 0 down, 1 up, 2 wgrad."""
    n, Hl, Wl, Clp, Hs, Ws, Csp, k = geom
    a = L.PlConvArgs(*geom, 0, 0, 0, n_out_pad, n_out_pad, geom[6], geom[3], s2d_cq,
                     L.tv(16, L.PARITY, Hl, Wl, Clp), L.tv(16, L.PLANAR, Hs, Ws, Csp), L.NO_TV, L.NO_T4, 16, None, 16, 0, 0)
    buf = C.create_string_buffer(1024)
    lib = L.load()
    rc = lib.mrssm_pl_describe(C.byref(a), op, buf, 1024)
    if rc != 0:
        raise RuntimeError(lib.mrssm_last_error().decode())
    s = buf.value.decode()
    d = {k: (float(v) if "." in v else int(v)) for k, v in re.findall(r"(\w+)=([\d.]+)", s)}
    d["text"] = s
    return d


def packed_shape(L, op, Csp, Clp, k):
    n, kk = C.c_int32(), C.c_int32()
    lib = L.load()
    assert lib.mrssm_pl_packed_shape(op, Csp, Clp, k, C.byref(n), C.byref(kk)) == 0, lib.mrssm_last_error()
    return n.value, kk.value


def pack_weight(w, op, Csp, Clp, N_total, K_total, cq=0):
    """numpy twin of pack_plane_kernel.  w: [Cs, Cl, k, k]."""
    Cs, Cl, k, _ = w.shape
    nt = (k + 1) // 2
    out = np.zeros((N_total, K_total), np.float32)
    for n in range(N_total):
        for kk in range(K_total):
            ks, e = kk >> 4, kk & 15
            if op == 2:
                J = Clp // 16
                j, t = ks % J, ks // J
                b, aa = t % nt, t // nt
                ch = (2 * j + (e >> 3)) * 8 + (e & 7)
                par, c = ch // cq, ch % cq
                kh, kw = 2 * aa + (par >> 1), 2 * b + (par & 1)
                if aa < nt and par < 4 and n < Cs and c < Cl and kh < k and kw < k:
                    out[n, kk] = w[n, c, kh, kw]
            elif op == 0:
                J = Clp // 8
                j, t = ks % J, ks // J
                b, kh = t % nt, t // nt
                pl = 2 * j + (e >> 3)
                px, chunk = pl // J, pl % J
                cl, kw = chunk * 8 + (e & 7), 2 * b + px
                if n < Cs and cl < Cl and kh < k and kw < k:
                    out[n, kk] = w[n, cl, kh, kw]
            else:
                J = Csp // 16
                j, t = ks % J, ks // J
                b, aa = t % nt, t // nt
                cls, cl = n // Clp, n % Clp
                cs = (2 * j + (e >> 3)) * 8 + (e & 7)
                kh, kw = (cls >> 1) + 2 * aa, (cls & 1) + 2 * b
                if aa < nt and cls < 4 and cl < Cl and cs < Cs and kh < k and kw < k:
                    out[n, kk] = w[cs, cl, kh, kw]
    return out


def _box(src, c0, x0, y0, i0, bx, by, bi, sx=1, sy=1, ox=0, oy=0):
    """TMA box {8, bx, by, bi} at (c0, x0, y0, i0) of src[n, H, W, C] sub-sampled (sy, oy)/(sx, ox); zeros out of bounds.
    Returns [bi*by*bx, 8] in shared-memory order."""
    n, H, W, Cc = src.shape
    sub = src[:, oy::sy, ox::sx, :]
    out = np.zeros((bi, by, bx, 8), np.float32)
    for i in range(bi):
        for y in range(by):
            for x in range(bx):
                ii, yy, xx = i0 + i, y0 + y, x0 + x
                if 0 <= ii < sub.shape[0] and 0 <= yy < sub.shape[1] and 0 <= xx < sub.shape[2]:
                    v = sub[ii, yy, xx, c0:c0 + 8]
                    out[i, y, x, :len(v)] = v
    return out.reshape(-1, 8)


def s2d16(x, cq):
    """[n,H,W,cq] -> space-to-depth [n,ceil(H/2),ceil(W/2),16] with channel = (py*2+px)*cq + c."""
    n, H, W, _ = x.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    out = np.zeros((n, H2, W2, 16), np.float32)
    for py in range(2):
        for px in range(2):
            sub = x[:, py::2, px::2, :cq]
            out[:, :sub.shape[1], :sub.shape[2], (py * 2 + px) * cq:(py * 2 + px + 1) * cq] = sub
    return out


def sim_fwd(L, op, src, w, Hl, Wl, Hs, Ws, n_out_pad, s2d_cq=0):
    """src: padded-channel NHWC source (down: large, up: small; the space-to-depth tensor when s2d_cq); w [Cs, Cl, k, k]
    master.  Returns the output tensor [n, Ho, Wo, n_out_pad] (NaN where the kernel would not write)."""
    n = src.shape[0]
    k = w.shape[2]
    nt = (k + 1) // 2
    if op == 0:
        Clp, Csp = src.shape[3], n_out_pad
    else:
        Csp, Clp = src.shape[3], n_out_pad
    geom = (n, Hl, Wl, Clp, Hs, Ws, Csp, k)
    P = plan(L, geom, op, n_out_pad, s2d_cq)
    pop = 2 if s2d_cq else op
    N_total, K_total = packed_shape(L, pop, Csp, Clp, k)
    wp = pack_weight(w, pop, Csp, Clp, N_total, K_total, s2d_cq)
    BI, BX, BY, TH, nb = P["BI"], P["BX"], P["BY"], P["TH"], P["bands"]
    PSpos = P["PS"] // 16
    planes, n_ksteps = P["planes"], P["ksteps"]
    if op == 0 and s2d_cq:
        J = Clp // 16
        Hv, Wv, Ho, Wo = Hs, Ws, Hs, Ws
    elif op == 0:
        J = Clp // 8
        Hv, Wv, Ho, Wo = Hs, Ws, Hs, Ws
    else:
        J = Csp // 16
        Hv, Wv, Ho, Wo = (Hl + 1) // 2, (Wl + 1) // 2, Hl, Wl
    out = np.full((n, Ho, Wo, n_out_pad), np.nan, np.float32)
    ngroups = (n + BI - 1) // BI
    for ig in range(ngroups):
        for band in range(nb):
            # the activation stage as one run of 16-byte positions: plane q starts at q * PS (grouped boxes put the planes back
            # to back, so what an MMA block reads past its plane's end is the next plane; behind the last plane: never written)
            MB = P["MB_total"]
            rows = MB * 128
            flat = np.full((P["stage"] // 16 + rows + 4 * BX + 8, 8), np.nan, np.float32)
            for q in range(planes):
                if op == 0 and s2d_cq:
                    box = _box(src, q * 8, 0, band * TH, ig * BI, BX, BY, BI)
                elif op == 0:
                    ppm = Clp // 8
                    mi, c0 = q // ppm, (q % ppm) * 8
                    box = _box(src, c0, 0, band * TH, ig * BI, BX, BY, BI, 2, 2, mi & 1, mi >> 1)
                else:
                    box = _box(src, q * 8, -(nt - 1), band * TH - (nt - 1), ig * BI, BX, BY, BI)
                flat[q * PSpos:q * PSpos + BI * BY * BX] = box
            A = [flat[q * PSpos:] for q in range(planes)]
            acc = np.zeros((rows, N_total), np.float32)
            for ks in range(n_ksteps):
                j, t = ks % J, ks // J
                b, r = t % nt, t // nt
                if op == 0 and s2d_cq:
                    plane, shift = 2 * j, r * BX + b
                elif op == 0:
                    plane, shift = (r & 1) * 2 * J + 2 * j, (r >> 1) * BX + b
                else:
                    plane, shift = 2 * j, (nt - 1 - r) * BX + (nt - 1 - b)
                a16 = np.concatenate([A[plane][shift:shift + rows], A[plane + 1][shift:shift + rows]], axis=1)
                acc += a16 @ wp[:, ks * 16:ks * 16 + 16].T
            IP = BY * BX
            for p in range(rows):
                i, r = p // IP, p % IP
                yr, x = r // BX, r % BX
                img, y = ig * BI + i, band * TH + yr
                if not (i < BI and img < n and yr < TH and y < Hv and x < Wv):
                    continue
                if op == 0:
                    out[img, y, x, :] = acc[p]
                else:
                    for cls in range(4):
                        yy, xx = 2 * y + (cls >> 1), 2 * x + (cls & 1)
                        if yy < Ho and xx < Wo:
                            out[img, yy, xx, :] = acc[p, cls * Clp:(cls + 1) * Clp]
    return out, P


class _Planes:
    """Chunk planes inside a flat stage: take(p0, count, shift, K) -> [K, count, 8] rows shift.. of planes p0.."""

    def __init__(self, flat, ps):
        self.flat, self.ps = flat, ps

    def take(self, p0, count, shift, K):
        return np.stack([self.flat[(p0 + q) * self.ps + shift:(p0 + q) * self.ps + shift + K] for q in range(count)], 1)


def sim_wgrad(L, small, large, k, s2d_cq=0, Hl=None):
    """small [n, Hs, Ws, Csp], large [n, Hl, Wl, Clp] (padded channels; the space-to-depth tensor of an Hl x Hl image when
    s2d_cq).  Returns dW [Csp, Clp (or s2d_cq), k, k]."""
    n, Hs, Ws, Csp = small.shape
    if s2d_cq:
        Clp, Wl = large.shape[3], Hl
    else:
        _, Hl, Wl, Clp = large.shape
    nt = (k + 1) // 2
    geom = (n, Hl, Wl, Clp, Hs, Ws, Csp, k)
    P = plan(L, geom, 2, 0, s2d_cq)
    BI, BX, BY, TH, nb = P["BI"], P["BX"], P["BY"], P["TH"], P["bands"]
    banded = nb > 1
    SBY = TH if banded else BY
    PS_s, PS_l = P["PS_s"] // 16, P["PS_l"] // 16
    nks = P["tile"]          # "ksteps/tile=%d" parses as key 'tile'
    cpl = Clp // 8
    dW = np.zeros((Csp, s2d_cq if s2d_cq else Clp, k, k), np.float32)
    ngroups = (n + BI - 1) // BI
    nL = cpl if s2d_cq else 4 * cpl
    for ig in range(ngroups):
        for band in range(nb):
            # the operand stage as runs of 16-byte positions: plane q at q * PS; what no box writes is zero (the kernel zeroes the
            # stages once).  Grouped boxes put the planes back to back, so reads past a plane's end see the next plane.
            K = nks * 16
            nSp = Csp // 8
            sflat = np.zeros((nSp * PS_s + K + 64, 8), np.float32)
            lflat = np.zeros((nL * PS_l + K + 4 * BX + 64, 8), np.float32)
            for q in range(nSp):
                sflat[q * PS_s:q * PS_s + BI * SBY * BX] = _box(small, q * 8, 0, band * TH, ig * BI, BX, SBY, BI)
            for q in range(nL):
                if s2d_cq:
                    box = _box(large, q * 8, 0, band * TH, ig * BI, BX, BY, BI)
                else:
                    mi, c0 = q // cpl, (q % cpl) * 8
                    box = _box(large, c0, 0, band * TH, ig * BI, BX, BY, BI, 2, 2, mi & 1, mi >> 1)
                lflat[q * PS_l:q * PS_l + BI * BY * BX] = box
            Sm = np.stack([sflat[q * PS_s:q * PS_s + K] for q in range(nSp)], 1).reshape(K, Csp)            # [pixel][cs]
            Lg = _Planes(lflat, PS_l)
            if s2d_cq:
                for g in range(nt * nt):
                    a, b = g // nt, g % nt
                    shift = a * BX + b
                    Bm = Lg.take(0, nL, shift, K).reshape(K, Clp)                          # [pixel][(parity, c)]
                    D = Sm.T @ Bm
                    for c in range(Clp):
                        par, cl = c // s2d_cq, c % s2d_cq
                        kh, kw = 2 * a + (par >> 1), 2 * b + (par & 1)
                        if par < 4 and kh < k and kw < k:
                            dW[:, cl, kh, kw] += D[:, c]
                continue
            if P.get("mrep"):
                # row-tap stacking: a second copy of the small tile one row lower (box origin y - 1) fills MMA rows 64..127;
                # an MMA at base tap kh = 4 j + parity then also accumulates tap kh + 2 there
                assert nSp == 8 and not banded
                s1 = np.zeros_like(sflat)
                for q in range(nSp):
                    s1[q * PS_s:q * PS_s + BI * SBY * BX] = _box(small, q * 8, 0, band * TH - 1, ig * BI, BX, SBY, BI)
                Sm1 = np.stack([s1[q * PS_s:q * PS_s + K] for q in range(nSp)], 1).reshape(K, Csp)
                A = np.concatenate([Sm, Sm1], axis=1)                       # [pixel][128 MMA rows]
                for g in range(2 * ((k + 3) // 4) * nt):
                    gk, b = g // nt, g % nt
                    kh = 4 * (gk >> 1) + (gk & 1)
                    plane0, shift = (kh & 1) * 2 * cpl, (kh >> 1) * BX + b
                    Bm = Lg.take(plane0, 2 * cpl, shift, K).reshape(K, 2 * Clp)
                    D = A.T @ Bm
                    for half in range(2):
                        khh = kh + 2 * half
                        for px in range(2):
                            kw = 2 * b + px
                            if khh < k and kw < k:
                                dW[:, :, khh, kw] += D[half * Csp:(half + 1) * Csp, px * Clp:(px + 1) * Clp]
                continue
            for g in range(k * nt):
                kh, b = g // nt, g % nt
                plane0, shift = (kh & 1) * 2 * cpl, (kh >> 1) * BX + b
                Bm = Lg.take(plane0, 2 * cpl, shift, K).reshape(K, 2 * Clp)   # [pixel][(px, cl)]
                D = Sm.T @ Bm
                for px in range(2):
                    kw = 2 * b + px
                    if kw < k:
                        dW[:, :, kh, kw] += D[:, px * Clp:(px + 1) * Clp]
    return dW, P
