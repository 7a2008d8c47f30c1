"""-m gpu: replay sampling on the device (`ExperienceReplay_Multimodal.sample` -> mrssm_replay_gather_u8 / mrssm_gather_rows
through the C ABI).

1. against the fixtures of the unmodified reference buffer (tests/golden/replay_*.pt): same seeded numpy / torch RNG
   streams, batches BIT-IDENTICAL (uint8 gather, power-of-two quantiser, the reference's rounding sequence);
2. the kernel against oracle/replay_oracle.py on shapes the fixtures do not reach (unaligned crops, 1 channel, other bit
   depths, widths that are not a multiple of 4 in the store), bit-exact;
3. in-kernel noise (no tensors supplied): same quantisation level as the oracle, dequantisation noise inside [0, 2^-bits),
   different per call;
4. the buffer feeds `optimize` (the D.sample contract of base/algo.py:236-240)."""
import os

import numpy as np
import pytest
import torch

from oracle import replay_oracle as RO
from tests import replay_util as R
from tests.replay_util import load_fixture_buffer as _load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cpu_rng_noise(kind, name, shape):
    """The reference draws its noise from torch's global (CPU) generator: randn for the Gaussian noise, rand for the
    dequantisation, in call order."""
    return torch.randn(*shape) if kind == "gauss" else torch.rand(shape)


@pytest.mark.parametrize("name", ["default", "augment"])
def test_sample_matches_reference_fixture_bit_for_bit(name, golden_dir, tmp_path):
    rec, cfg, D = _load(name, golden_dir, tmp_path, device=DEV)
    assert all(v.is_cuda for v in D.observations.values()) and D.observations[R.IMAGE].dtype == torch.uint8
    D.noise_source = _cpu_rng_noise
    for s in rec["samples"]:
        np.random.seed(s["seed"])
        torch.manual_seed(s["seed"])
        obs, actions, rewards, nonterminals = D.sample(R.N, R.L)
        for k, d in s["obs"].items():
            assert obs[k].is_cuda and obs[k].dtype == torch.float32
            R.assert_digest(obs[k], d)
        assert torch.equal(actions.cpu(), s["actions"]) and torch.equal(rewards.cpu(), s["rewards"])
        assert torch.equal(nonterminals.cpu(), s["nonterminals"])
        assert float(np.random.rand()) == s["np_next"] and float(torch.rand(())) == s["torch_next"]


CASES = [
    # C, Hs, Ws, H, W, dh, dw, bits, delta, gauss_scale
    (3, 64, 64, 64, 64, 0, 0, 5, False, 0.0),        # the hot case: whole frames, aligned 4-byte loads
    (3, 70, 70, 64, 64, 3, 5, 5, True, 0.03),        # odd crop origin -> byte loads
    (1, 66, 67, 64, 64, 2, 3, 8, False, 0.1),        # store width not a multiple of 4, 8-bit (no quantisation)
    (4, 40, 48, 32, 32, 8, 12, 1, True, 0.0),        # other sizes, 1-bit, colour shift only
    (3, 64, 64, 64, 64, 0, 0, 0, False, 0.0),        # raw (binary-mask path): no normalisation
]


@pytest.mark.parametrize("case", CASES)
def test_gather_kernel_matches_oracle(case):
    import ctypes as C
    from mrssm_b200 import _lib as L
    Cc, Hs, Ws, H, W, dh, dw, bits, use_delta, gscale = case
    g = torch.Generator().manual_seed(sum(case[:7]))
    size, n, Lc = 23, 5, 3
    store = torch.randint(0, 256, (size, Cc, Hs, Ws), dtype=torch.uint8, generator=g)
    vec = torch.randint(0, size, (n * Lc,), generator=g)
    delta = (torch.randn(Cc, generator=g) * 20) if use_delta else None
    shape = (n * Lc, Cc, H, W)
    gauss = torch.randn(shape, generator=g) if gscale > 0 else None
    uniform = torch.rand(shape, generator=g) if bits else None
    assert H == W
    ref = RO.gather_image(store, vec.numpy(), n, Lc, crop=(dh, dw), side=H, delta=delta, gauss=gauss, gauss_scale=gscale,
                          uniform=uniform, bit_depth=bits, normalise=bits > 0)
    dev = torch.device(DEV)
    tens = [None if t is None else t.to(dev).contiguous() for t in (store, vec, delta, gauss, uniform)]
    out = torch.empty(shape, device=dev, dtype=torch.float32)
    a = L.ReplayGatherArgs()
    a.frames, a.idx, a.rows = L.ptr_any(tens[0]), L.ptr_any(tens[1]), n * Lc
    a.C, a.Hs, a.Ws, a.H, a.W, a.dh, a.dw, a.bit_depth = Cc, Hs, Ws, H, W, dh, dw, bits
    a.delta, a.gauss, a.gauss_scale, a.uniform, a.seed, a.out = L.ptr(tens[2]), L.ptr(tens[3]), gscale, L.ptr(tens[4]), 1, L.ptr(out)
    L.call("mrssm_replay_gather_u8", C.byref(a))
    assert torch.equal(out.cpu().reshape(ref.shape), ref)


def test_gather_rows_and_bad_arguments():
    import ctypes as C
    from mrssm_b200 import _lib as L
    dev = torch.device(DEV)
    src = torch.randn(17, 7, device=dev)
    idx = torch.randint(0, 17, (40,), device=dev)
    out = torch.empty(40, 7, device=dev)
    L.call("mrssm_gather_rows", L.ptr(src), L.ptr_any(idx), 40, 7, L.ptr(out))
    assert torch.equal(out, src[idx])
    a = L.ReplayGatherArgs()
    frames = torch.zeros(2, 3, 64, 64, dtype=torch.uint8, device=dev)
    o = torch.empty(1, 3, 64, 64, device=dev)
    a.frames, a.idx, a.rows, a.out = L.ptr_any(frames), L.ptr_any(idx), 1, L.ptr(o)
    a.C, a.Hs, a.Ws, a.H, a.W, a.dh, a.dw, a.bit_depth = 3, 64, 64, 64, 64, 1, 0, 5         # crop leaves the frame
    with pytest.raises(RuntimeError, match="crop outside"):
        L.call("mrssm_replay_gather_u8", C.byref(a))


def test_in_kernel_noise(golden_dir, tmp_path):
    rec, cfg, D = _load("augment", golden_dir, tmp_path, device=DEV)
    batches = []
    for rep in range(2):
        np.random.seed(3)                       # same chunks and augmentation choices both times
        batches.append(D.sample(R.N, R.L)[0])
    np.random.seed(3)
    idxs = np.asarray([D._sample_idx(R.L) for _ in range(R.N)])
    vec_idxs, plan = D._plan_batch(idxs)
    p = plan[R.IMAGE]
    a, b = batches[0][R.IMAGE], batches[1][R.IMAGE]
    assert not torch.equal(a, b)                                    # fresh noise per call
    lvl = torch.floor((a + 0.5) * 32)                               # quantisation level 0..31
    frac = (a + 0.5) * 32 - lvl
    assert float(lvl.min()) >= 0 and float(lvl.max()) <= 32     # 32 only when level 31 + noise rounds up to 0.5
    assert 0.45 < float(frac.mean()) < 0.55                         # dequantisation noise ~ U[0,1) / 32
    # with the Gaussian noise off the level is deterministic: compare to the oracle with the colour shift only
    D.noise_scales = [0.0]
    np.random.seed(3)
    c = D.sample(R.N, R.L)[0][R.IMAGE]
    np.random.seed(3)
    idxs = np.asarray([D._sample_idx(R.L) for _ in range(R.N)])
    vec_idxs, plan = D._plan_batch(idxs)
    p = plan[R.IMAGE]
    ref = RO.gather_image(D.observations[R.IMAGE].cpu(), vec_idxs, R.N, R.L, crop=p["crop"], side=p["side"], delta=p["delta"],
                          uniform=torch.zeros(R.L, R.N, 3, 64, 64), bit_depth=5)
    noise = (c.cpu() + 0.5) * 32 - (ref + 0.5) * 32             # ref carries zero noise: the level itself
    assert float(noise.min()) >= 0.0 and float(noise.max()) <= 1.0      # 1.0 only when fp32 rounds level + u up (u -> 1)
    # binary masks come back raw (0 / 255), cropped
    m = batches[0][R.BIN]
    assert m.shape == (R.L, R.N, 1, 64, 64) and set(m.unique().tolist()) <= {0.0, 255.0}


def test_buffer_feeds_optimize(golden_dir, tmp_path):
    from tests import parity_util as U
    rec, cfg, D = _load("default", golden_dir, tmp_path, device=DEV)
    oc = U.oracle_cfg("MoPoE")
    model, _ = U.build_product(oc, R.N, R.L, DEV)
    np.random.seed(1)
    for _ in range(2):
        model.optimize(D)
    assert torch.isfinite(model.model_loss).item()
    assert float(model.model_optimizer.grad_norm) > 0
