"""CPU: replay the tcgen05 rollout's host-built MMA program (csrc/rollout_tc.cu, mrssm_rollout_tc_plan) in numpy — pack the
weights exactly as rollout_tc_pack_kernel does, execute every MMA of one time step on random operand buffers, and compare
the TMEM columns at every commit point with the plain matrix products of transition_model.py:232-262 / encoder.py:126-190.
Checks the operand chunk addressing, the weight row segments (gate order r,z,n; mean / std split of fc2), the TMEM
column map, tile sizes and the wait / commit wiring."""
import ctypes

import numpy as np
import pytest

ROWS, CH_BYTES, SLOT_BYTES, ACC, F2 = 64, 1024, 14336, 112, 224
TILE = np.dtype([("src_off", "<u4"), ("bytes", "<u4"), ("op_begin", "<u2"), ("op_end", "<u2"), ("wait_ev", "i1"), ("commit", "i1"),
                 ("run_begin", "u1"), ("run_end", "u1")])
RUN = np.dtype([("a_off16", "<u2"), ("a_buf", "u1"), ("acc_first", "u1"), ("b_off16", "<u2"), ("d_col", "<u2"), ("idesc", "<u4"), ("count", "<u4")])
OP = np.dtype([("a_off16", "<u2"), ("a_buf", "u1"), ("acc", "u1"), ("b_off16", "<u2"), ("d_col", "<u2"), ("idesc", "<u4")])
PACK = np.dtype([("src_id", "<i4"), ("k0", "<i4"), ("kvalid", "<i4"), ("N", "<i4"), ("seg_n", "<i4", 2), ("seg_src", "<i4", 2),
                 ("seg_cnt", "<i4", 2), ("dst_off", "<u4"), ("pad", "<i4")])
EV_XIN, EV_X, EV_HA, EV_HB, EV_U0 = 0, 1, 2, 3, 4
CM_X, CM_GA, CM_GB, CM_F1, CM_F2 = 0, 1, 2, 3, 11


def _plan(D, S, H, A, E):
    from mrssm_b200 import _lib as L
    L.load()
    pb, kb = ctypes.c_int64(), ctypes.c_int64()
    L.call_host("mrssm_rollout_tc_plan_bytes", D, S, H, A, E, ctypes.byref(pb), ctypes.byref(kb))
    buf = ctypes.create_string_buffer(pb.value)
    L.call_host("mrssm_rollout_tc_plan", D, S, H, A, E, ctypes.cast(buf, ctypes.c_void_p), pb.value)
    raw = np.frombuffer(buf.raw, dtype=np.uint8)
    hdr = raw[:80].view("<i4")
    n_tiles, n_ops, n_pack = int(hdr[6]), int(hdr[7]), int(hdr[8])
    hu = raw[:80].view("<u4")
    packed_bytes, tile_off, op_off, pack_off, total, run_off, n_runs = [int(v) for v in hu[13:20]]
    assert total == pb.value and packed_bytes == kb.value
    tiles = raw[tile_off:tile_off + n_tiles * TILE.itemsize].view(TILE)
    ops = raw[op_off:op_off + n_ops * OP.itemsize].view(OP)
    packs = raw[pack_off:pack_off + n_pack * PACK.itemsize].view(PACK)
    runs = raw[run_off:run_off + n_runs * RUN.itemsize].view(RUN)
    # the runs the kernel issues expand to exactly the MMA list
    for tl in tiles:
        exp = []
        for r in runs[int(tl["run_begin"]):int(tl["run_end"])]:
            N = ((int(r["idesc"]) >> 17) & 63) << 3
            assert 1 <= int(r["count"]) <= 7
            for i in range(int(r["count"])):
                exp.append((int(r["a_off16"]) + 128 * i, int(r["a_buf"]), int(r["acc_first"]) if i == 0 else 1,
                            int(r["b_off16"]) + 2 * N * i, int(r["d_col"]), int(r["idesc"])))
        got = [tuple(int(o[k]) for k in ("a_off16", "a_buf", "acc", "b_off16", "d_col", "idesc"))
               for o in ops[int(tl["op_begin"]):int(tl["op_end"])]]
        assert exp == got
    assert n_runs < 192
    return dict(cA=int(hdr[9]), nD8=int(hdr[10]), cAH=int(hdr[11]), nH8=int(hdr[12])), tiles, ops, packs, packed_bytes


def _pack(packs, srcs, packed_bytes):
    """numpy twin of rollout_tc_pack_kernel: block [2 chunks][N][8] per MMA."""
    out = np.zeros(packed_bytes // 2, dtype=np.float32)      # one float per bf16 slot
    for pk in packs:
        w = srcs[int(pk["src_id"])]
        N = int(pk["N"])
        blk = np.zeros((2, N, 8), dtype=np.float32)
        for s in range(2):
            n0, r0, cnt = int(pk["seg_n"][s]), int(pk["seg_src"][s]), int(pk["seg_cnt"][s])
            for k in range(int(pk["kvalid"])):
                blk[k // 8, n0:n0 + cnt, k % 8] = w[r0:r0 + cnt, int(pk["k0"]) + k]
        out[int(pk["dst_off"]) // 2: int(pk["dst_off"]) // 2 + 16 * N] = blk.reshape(-1)
    return out


@pytest.mark.parametrize("dims", [(200, 30, 200, 3, 3), (200, 30, 200, 3, 0), (64, 8, 48, 2, 2), (208, 32, 104, 6, 3), (200, 30, 200, 3, 1)])
def test_program_computes_the_step_gemms(dims):
    D, S, H, A, E = dims
    NH = 1 + E
    hd, tiles, ops, packs, packed_bytes = _plan(D, S, H, A, E)
    rng = np.random.default_rng(0)
    w_sa = rng.standard_normal((D, S + A)).astype(np.float32)
    w_ih = rng.standard_normal((3 * D, D)).astype(np.float32)
    w_hh = rng.standard_normal((3 * D, D)).astype(np.float32)
    ld1 = [D, D, D + 24, D + 8][:NH]
    w1 = [rng.standard_normal((H, ld)).astype(np.float32) for ld in ld1]
    w2 = [rng.standard_normal((2 * S, H)).astype(np.float32) for _ in range(NH)]
    srcs = {0: w_sa, 1: w_ih, 2: w_hh}
    for h in range(NH):
        srcs[3 + h], srcs[7 + h] = w1[h], w2[h]
    image = _pack(packs, srcs, packed_bytes)

    # operand buffers as [chunk][row][8] planes; padding chunks zero like the kernel keeps them
    def planes(mat, nch):
        buf = np.zeros((nch, ROWS, 8), dtype=np.float32)
        for c in range(mat.shape[1] // 8 + (1 if mat.shape[1] % 8 else 0)):
            w = min(8, mat.shape[1] - 8 * c)
            buf[c, :, :w] = mat[:, 8 * c:8 * c + w]
        return buf
    xin, x, hp, hn, u = (rng.standard_normal((ROWS, n)).astype(np.float32) for n in (S + A, D, D, D, H))
    bufs = {0: planes(xin, 6), 1: planes(x, 26), 2: planes(hp, 26), 3: planes(hn, 26)}
    tmem = np.full((ROWS, 512), np.nan, dtype=np.float32)
    cA, nD8, cAH, nH8 = hd["cA"], hd["nD8"], hd["cAH"], hd["nH8"]
    seen_commits, waits = [], []
    off = 0
    for tl in tiles:
        assert int(tl["src_off"]) == off and 0 < int(tl["bytes"]) <= SLOT_BYTES and int(tl["bytes"]) % 16 == 0
        off += int(tl["bytes"])
        if tl["wait_ev"] >= 0:
            waits.append(int(tl["wait_ev"]))
            if int(tl["wait_ev"]) >= EV_U0 + 1 and (int(tl["wait_ev"]) - EV_U0) % 2 == 1:
                bufs[1] = planes(u, 26)          # fc2 reads U from the buffer X lived in
        for o in ops[int(tl["op_begin"]):int(tl["op_end"])]:
            N = ((int(o["idesc"]) >> 17) & 63) << 3
            assert (int(o["idesc"]) >> 24) & 31 == ROWS >> 4
            a_ch = int(o["a_off16"]) * 16 // CH_BYTES
            Amat = np.concatenate([bufs[int(o["a_buf"])][a_ch], bufs[int(o["a_buf"])][a_ch + 1]], axis=1)        # [64][16]
            b0 = (int(tl["src_off"]) + int(o["b_off16"]) * 16) // 2
            assert int(o["b_off16"]) * 16 + 32 * N <= int(tl["bytes"])
            blk = image[b0:b0 + 16 * N].reshape(2, N, 8)
            Bmat = np.concatenate([blk[0], blk[1]], axis=1)                                                       # [N][16]
            prod = Amat @ Bmat.T
            c0 = int(o["d_col"])
            assert c0 + N <= 512
            tmem[:, c0:c0 + N] = prod + (tmem[:, c0:c0 + N] if o["acc"] else 0.0)
        if tl["commit"] >= 0:
            cm = int(tl["commit"])
            seen_commits.append(cm)
            tol = dict(rtol=1e-4, atol=1e-3)
            if cm == CM_X:
                np.testing.assert_allclose(tmem[:, :D], xin @ w_sa.T, **tol)
            elif cm in (CM_GA, CM_GB):
                n0, cnt = (0, 8 * cA) if cm == CM_GA else (8 * cA, 8 * (nD8 - cA))
                sl = slice(n0, n0 + cnt)
                np.testing.assert_allclose(tmem[:, 0:cnt], x @ w_ih[0:D][sl].T + hp @ w_hh[0:D][sl].T, **tol)
                np.testing.assert_allclose(tmem[:, ACC:ACC + cnt], x @ w_ih[D:2 * D][sl].T + hp @ w_hh[D:2 * D][sl].T, **tol)
                np.testing.assert_allclose(tmem[:, 2 * ACC:2 * ACC + cnt], x @ w_ih[2 * D:][sl].T, **tol)
                np.testing.assert_allclose(tmem[:, 3 * ACC:3 * ACC + cnt], hp @ w_hh[2 * D:][sl].T, **tol)
            elif CM_F1 <= cm < CM_F2:
                h, hf = (cm - CM_F1) // 2, (cm - CM_F1) % 2
                n0, cnt = (0, 8 * cAH) if hf == 0 else (8 * cAH, 8 * (nH8 - cAH))
                np.testing.assert_allclose(tmem[:, hf * ACC:hf * ACC + cnt], hn @ w1[h][n0:n0 + cnt, :D].T, **tol)
            else:
                h = cm - CM_F2
                np.testing.assert_allclose(tmem[:, F2 + 64 * h:F2 + 64 * h + S], u @ w2[h][:S].T, **tol)
                np.testing.assert_allclose(tmem[:, F2 + 64 * h + 32:F2 + 64 * h + 32 + S], u @ w2[h][S:].T, **tol)
    assert off == packed_bytes
    expect = [CM_X, CM_GA, CM_GB, CM_F1, CM_F1 + 1]
    for h in range(1, NH):
        expect += [CM_F1 + 2 * h, CM_F2 + h - 1, CM_F1 + 2 * h + 1]
    expect.append(CM_F2 + NH - 1)
    assert seen_commits == expect
    assert waits[:4] == [EV_XIN, EV_X, EV_HA, EV_HB]


# ---- BPTT program ---------------------------------------------------------------------------------------------------
EVB_DO, EVB_DU0, EVB_GRU, EVB_DX = 0, 1, 9, 10
CMB_GB0, CMB_GCH0, CMB_GE, CMB_GF = 0, 8, 12, 13
BCH_DU, BCOL_SET, BCOL_GH, BCOL_GX, BCOL_GHH, BCOL_XIN = 32, 112, 224, 0, 208, 448
MAXC = 26


def _plan_bwd(D, S, H, A, E):
    from mrssm_b200 import _lib as L
    L.load()
    pb, kb = ctypes.c_int64(), ctypes.c_int64()
    L.call_host("mrssm_rollout_tc_bwd_plan_bytes", D, S, H, A, E, ctypes.byref(pb), ctypes.byref(kb))
    buf = ctypes.create_string_buffer(pb.value)
    L.call_host("mrssm_rollout_tc_bwd_plan", D, S, H, A, E, ctypes.cast(buf, ctypes.c_void_p), pb.value)
    raw = np.frombuffer(buf.raw, dtype=np.uint8)
    hdr, hu = raw[:80].view("<i4"), raw[:80].view("<u4")
    n_tiles, n_ops, n_pack = int(hdr[6]), int(hdr[7]), int(hdr[8])
    packed_bytes, tile_off, op_off, pack_off = [int(v) for v in hu[13:17]]
    tiles = raw[tile_off:tile_off + n_tiles * TILE.itemsize].view(TILE)
    ops = raw[op_off:op_off + n_ops * OP.itemsize].view(OP)
    packs = raw[pack_off:pack_off + n_pack * PACK.itemsize].view(PACK)
    return dict(cAH=int(hdr[11]), nH8=int(hdr[12])), tiles, ops, packs, packed_bytes


def _pack_any(packs, srcs, packed_bytes):
    """numpy twin of rollout_tc_pack_kernel including the transposed (dgrad-type) blocks."""
    out = np.zeros(packed_bytes // 2, dtype=np.float32)
    for pk in packs:
        sid, N = int(pk["src_id"]), int(pk["N"])
        w = srcs[sid & 0xff]
        blk = np.zeros((2, N, 8), dtype=np.float32)
        if sid & 0x100:           # Bblock[n][k'] = W[row(k0 + k')][n_off + n], n < n_cnt
            n_cnt, n_off = int(pk["kvalid"]), int(pk["pad"])
            for k in range(16):
                kk = int(pk["k0"]) + k
                for s in range(2):
                    k0s, r0, cnt = int(pk["seg_n"][s]), int(pk["seg_src"][s]), int(pk["seg_cnt"][s])
                    if k0s <= kk < k0s + cnt:
                        blk[k // 8, :n_cnt, k % 8] = w[r0 + kk - k0s, n_off:n_off + n_cnt]
        else:
            for s in range(2):
                n0, r0, cnt = int(pk["seg_n"][s]), int(pk["seg_src"][s]), int(pk["seg_cnt"][s])
                for k in range(int(pk["kvalid"])):
                    blk[k // 8, n0:n0 + cnt, k % 8] = w[r0:r0 + cnt, int(pk["k0"]) + k]
        out[int(pk["dst_off"]) // 2: int(pk["dst_off"]) // 2 + 16 * N] = blk.reshape(-1)
    return out


@pytest.mark.parametrize("dims", [(200, 30, 200, 3, 3), (200, 30, 200, 3, 0), (64, 8, 48, 2, 2), (208, 32, 104, 6, 1)])
def test_bptt_program_computes_the_step_gemms(dims):
    """Same replay for the backward program: go W2, du W1[:, :D] summed over heads, [dr dz dn] W_ih, [dr dz dn*r] W_hh,
    dxpre W_sa, with the operand region re-used exactly as the kernel re-uses it."""
    D, S, H, A, E = dims
    NH = 1 + E
    hd, tiles, ops, packs, packed_bytes = _plan_bwd(D, S, H, A, E)
    rng = np.random.default_rng(1)
    w_sa = rng.standard_normal((D, S + A)).astype(np.float32)
    w_ih = rng.standard_normal((3 * D, D)).astype(np.float32)
    w_hh = rng.standard_normal((3 * D, D)).astype(np.float32)
    ld1 = [D, D, D + 24, D + 8][:NH]
    w1 = [rng.standard_normal((H, ld)).astype(np.float32) for ld in ld1]
    w2 = [rng.standard_normal((2 * S, H)).astype(np.float32) for _ in range(NH)]
    srcs = {0: w_sa, 1: w_ih, 2: w_hh}
    for h in range(NH):
        srcs[3 + h], srcs[7 + h] = w1[h], w2[h]
    image = _pack_any(packs, srcs, packed_bytes)

    go = [rng.standard_normal((ROWS, 2 * S)).astype(np.float32) for _ in range(NH)]
    du = [rng.standard_normal((ROWS, H)).astype(np.float32) for _ in range(NH)]
    dr, dz, dn, dnr, dx = (rng.standard_normal((ROWS, D)).astype(np.float32) for _ in range(5))
    region = np.zeros((4 * MAXC, ROWS, 8), dtype=np.float32)

    def put(chunk0, mat):
        nch = (mat.shape[1] + 7) // 8
        region[chunk0:chunk0 + nch] = 0.0
        for c in range(nch):
            w = min(8, mat.shape[1] - 8 * c)
            region[chunk0 + c, :, :w] = mat[:, 8 * c:8 * c + w]

    def put_go(h):                  # means at K 0..S-1, raw stds at K 32..32+S-1
        m = np.zeros((ROWS, 64), dtype=np.float32)
        m[:, :S], m[:, 32:32 + S] = go[h][:, :S], go[h][:, S:]
        put(8 * h, m)

    for h in range(NH):
        put_go(h)
    tmem = np.full((ROWS, 512), np.nan, dtype=np.float32)
    cAH, nH8 = hd["cAH"], hd["nH8"]
    gh_ref = np.zeros((ROWS, D), dtype=np.float32)
    seen, off = [], 0
    for tl in tiles:
        assert int(tl["src_off"]) == off and 0 < int(tl["bytes"]) <= SLOT_BYTES
        off += int(tl["bytes"])
        ev = int(tl["wait_ev"])
        if ev >= EVB_DU0 and ev < EVB_GRU and (ev - EVB_DU0) % 2 == 1:
            put(BCH_DU, du[(ev - EVB_DU0) // 2])                       # the du buffer is rewritten per head
        elif ev == EVB_GRU:
            put(0, dr), put(MAXC, dz), put(2 * MAXC, dn), put(3 * MAXC, dnr)
        elif ev == EVB_DX:
            put(0, dx)
        for o in ops[int(tl["op_begin"]):int(tl["op_end"])]:
            N = ((int(o["idesc"]) >> 17) & 63) << 3
            a_ch = int(o["a_off16"]) * 16 // CH_BYTES
            Amat = np.concatenate([region[a_ch], region[a_ch + 1]], axis=1)
            b0 = (int(tl["src_off"]) + int(o["b_off16"]) * 16) // 2
            blk = image[b0:b0 + 16 * N].reshape(2, N, 8)
            prod = Amat @ np.concatenate([blk[0], blk[1]], axis=1).T
            c0 = int(o["d_col"])
            assert c0 + N <= 512
            tmem[:, c0:c0 + N] = prod + (tmem[:, c0:c0 + N] if o["acc"] else 0.0)
        cm = int(tl["commit"])
        if cm < 0:
            continue
        seen.append(cm)
        tol = dict(rtol=1e-4, atol=2e-3)
        if CMB_GB0 <= cm < CMB_GCH0:
            h, hf = (cm - CMB_GB0) // 2, (cm - CMB_GB0) % 2
            n0, cnt = (0, 8 * cAH) if hf == 0 else (8 * cAH, 8 * (nH8 - cAH))
            np.testing.assert_allclose(tmem[:, hf * BCOL_SET:hf * BCOL_SET + cnt], go[h] @ w2[h][:, n0:n0 + cnt], **tol)
        elif CMB_GCH0 <= cm < CMB_GE:
            h = cm - CMB_GCH0
            gh_ref = gh_ref + du[h] @ w1[h][:, :D]
            np.testing.assert_allclose(tmem[:, BCOL_GH:BCOL_GH + D], gh_ref, **tol)
        elif cm == CMB_GE:
            np.testing.assert_allclose(tmem[:, BCOL_GX:BCOL_GX + D], np.concatenate([dr, dz, dn], 1) @ w_ih, **tol)
            np.testing.assert_allclose(tmem[:, BCOL_GHH:BCOL_GHH + D], np.concatenate([dr, dz, dnr], 1) @ w_hh, **tol)
        else:
            assert cm == CMB_GF
            np.testing.assert_allclose(tmem[:, BCOL_XIN:BCOL_XIN + S + A], dx @ w_sa, **tol)
    assert off == packed_bytes
    expect = [CMB_GB0, CMB_GB0 + 1]
    for h in range(NH):
        if h + 1 < NH:
            expect.append(CMB_GB0 + 2 * (h + 1))
        expect.append(CMB_GCH0 + h)
        if h + 1 < NH:
            expect.append(CMB_GB0 + 2 * (h + 1) + 1)
    expect += [CMB_GE, CMB_GF]
    assert seen == expect
