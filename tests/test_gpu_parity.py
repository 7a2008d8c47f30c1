"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle and the committed reference fixtures."""
import os

import pytest
import torch

from oracle import mrssm_oracle as O
from tests import parity_util as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3      # BASELINE.json north_star: rtol 1e-3 in fp32 mode


@pytest.mark.parametrize("fusion,kw", [
    ("MoPoE", {}), ("PoE", {}), ("NN", {}), ("single", {}),
    ("MoPoE", dict(grad_clip_norm=0.5, kl_balancing_alpha=None)),
    ("PoE", dict(kl_balancing_alpha=None, global_kl_beta=0.0, free_nats=0.5)),
    ("MoPoE", dict(predict_reward=True)),
])
def test_train_step_matches_oracle(fusion, kw):
    out = U.run_train_parity(fusion, B=4, T=6, steps=2, device=DEV, rtol=RTOL, **kw)
    assert out["worst_grad_err"] < 2 * RTOL


@pytest.mark.parametrize("name", ["mopoe", "poe", "nn", "single", "mopoe_clip", "poe_noalpha", "mopoe_reward",
                                  "mopoe_over", "poe_over", "single_over",
                                  # LogProb loss (base/algo.py:164-167), LR ramp (:208-212), fc branch of the conv encoder
                                  # (encoder.py:261-262), 128x128 stacks (encoder.py:415-509, observation_model.py:162-229)
                                  "mopoe_logprob", "single_logprob", "mopoe_lrramp", "mopoe_emb512", "mopoe_img128",
                                  # the shipped YAML's layers: BatchNorm image stacks (encoder.py:324-337, observation_model.py:75-86)
                                  # and the sound modality (encoder.py:661-721, observation_model.py:420-472), running statistics
                                  # after every step and an eval-mode pass included
                                  "mopoe_bn", "single_bn", "mopoe_sound", "mopoe_sound_bn",
                                  # the remaining image stacks: 84x84 and 256x256 (encoder.py:362-413,511-615; observation_model.py:
                                  # 108-160,231-345), BatchNorm on a stack other than 64x64
                                  "mopoe_img84", "mopoe_img256", "single_img84_bn"])
def test_train_step_matches_reference_fixture(name, golden_dir):
    """Directly against tests/golden/train_*.pt (outputs of the unmodified reference)."""
    rec = torch.load(os.path.join(golden_dir, f"train_{name}.pt"), weights_only=False)
    oc = O.OracleConfig(**rec["meta"]["cfg"])
    B, T = rec["meta"]["B"], rec["meta"]["T"]
    model, P = U.build_product(oc, B, T, DEV)
    named = U.named_params(model, oc)
    for step in rec["steps"]:
        batch, noise = O.synthetic_batch(oc, B, T, seed=step["data_seed"])
        st = U.product_step(model, oc, batch, noise, DEV)
        U.assert_states_close(st, step["states"], RTOL, 2e-5)
        info = {k: float(v) for k, v in model.loss_info.items()}
        for k, v in step["loss_info"].items():
            assert info[k] == pytest.approx(v, rel=RTOL, abs=1e-5), k
        assert float(model.model_optimizer.grad_norm) == pytest.approx(step["grad_norm"], rel=RTOL)
        gmax = max(s["norm"] for s in step["grads"].values())
        for k, s in step["grads"].items():
            g = named[k].grad.detach().cpu().reshape(-1)
            assert float(g.double().norm()) == pytest.approx(s["norm"], rel=2 * RTOL, abs=1e-6 * gmax), k
            torch.testing.assert_close(g[s["idx"]], s["val"], rtol=5 * RTOL, atol=1e-5 * max(1.0, s["norm"]))
        for k in step["grad_none"]:
            assert float(named[k].grad.abs().max()) == 0.0
        for k, s in step["params_after"].items():
            p = named[k].detach().cpu().reshape(-1)
            torch.testing.assert_close(p[s["idx"]], s["val"], rtol=RTOL, atol=1e-5)
        bufs = U.named_buffers(model, oc) if "buffers_after" in step else {}
        for k, b in step.get("buffers_after", {}).items():            # running statistics / batch counters of the norm layers
            if b.dtype == torch.long:
                assert int(bufs[k]) == int(b), k
            else:
                torch.testing.assert_close(bufs[k].detach().cpu(), b, rtol=RTOL, atol=1e-6, msg=lambda m: f"buffer {k}: {m}")
    if "eval" in rec:                                                  # eval mode: the running statistics normalise
        ev = rec["eval"]
        batch, _ = O.synthetic_batch(oc, B, T, seed=ev["data_seed"])
        dev = torch.device(DEV)
        tgt = {n: batch["obs"][n][1:].to(dev) for n in oc.names_enc}
        model.eval()
        try:
            with torch.no_grad():
                st = model.estimate_state(tgt, batch["actions"][:-1].to(dev), None, batch["nonterminals"][:-1].to(dev), det=True)
                out = model.observation_model(h_t=st["beliefs"], s_t=st["posterior_states"])
        finally:
            model.train()
        U.assert_states_close(st, ev["states"], RTOL, 2e-5)
        for n, s in ev["recon"].items():
            r = (out[n] if oc.multimodal else out)["loc"].detach().cpu().reshape(-1)
            torch.testing.assert_close(r[s["idx"]], s["val"], rtol=RTOL, atol=1e-5)
            assert float(r.double().norm()) == pytest.approx(s["norm"], rel=RTOL), n


@pytest.mark.parametrize("fusion", ["MoPoE", "single"])
def test_bf16_tensor_core_mode_within_stated_tolerance(fusion):
    """bf16 mode (tcgen05 conv stacks, fp32 masters/accumulators) against the fp32 oracle.
    Stated tolerance: latents/KL/losses within 2e-2 relative, total gradient within 5e-2 (Frobenius)."""
    rep = U.run_train_parity_bf16(fusion, B=4, T=6, steps=2, device=DEV)
    assert rep["state_err"] < 2e-2, rep
    assert rep["loss_rel"] < 2e-2, rep
    assert rep["gnorm_rel"] < 3e-2, rep
    assert rep["grad_rel_fro"] < 5e-2, rep


def test_bf16_mode_shipped_layers_within_stated_tolerance():
    """The shipped YAML's layer set (image + sound, BatchNorm) in bf16 mode — the 128 .. 512-channel convolutions on the tensor-core
    route (ops.GConvTCFn), normalisation / GLU in fp32 — against the fp32 oracle, same stated tolerances as the other bf16 cases."""
    shapes = {"image_horizon": [3, 64, 64], "sound": [128, 20]}
    rep = U.run_train_parity_bf16("MoPoE", B=2, T=5, steps=2, device=DEV, names_enc=("image_horizon", "sound"),
                                  names_rec=("image_horizon", "sound"), observation_shapes=shapes, normalization="BatchNorm", lr=1e-5)
    assert rep["state_err"] < 2e-2 and rep["loss_rel"] < 2e-2 and rep["gnorm_rel"] < 3e-2 and rep["grad_rel_fro"] < 5e-2, rep


@pytest.mark.parametrize("side", [84, 256])
def test_bf16_mode_other_image_sizes_within_stated_tolerance(side):
    """84x84 / 256x256 stacks in bf16 mode (general NCHW kernels, tensor-core route for their wide layers) against the fp32 oracle."""
    name = f"image_horizon_{side}"
    rep = U.run_train_parity_bf16("MoPoE", B=3, T=5, steps=2, device=DEV, names_enc=(name, "pose_quat_v2"), names_rec=(name, "pose_quat_v2"),
                                  observation_shapes={name: [3, side, side], "pose_quat_v2": [3]})
    assert rep["state_err"] < 2e-2 and rep["loss_rel"] < 2e-2 and rep["gnorm_rel"] < 3e-2 and rep["grad_rel_fro"] < 5e-2, rep


def test_normalize_image_u8_matches_reference_formula():
    """mrssm_normalize_image_u8 against image_processing.py:5-11 restated in torch, with the noise supplied."""
    from mrssm_b200 import _lib as L
    g = torch.Generator(device=DEV).manual_seed(3)
    u8 = torch.randint(0, 256, (5, 3, 64, 64), generator=g, device=DEV, dtype=torch.uint8)
    noise = torch.rand(u8.shape, generator=g, device=DEV)
    out = torch.empty(u8.shape, device=DEV)
    for bits in (5, 8):
        L.call("mrssm_normalize_image_u8", L.ptr_any(u8), u8.numel(), bits, L.ptr(noise), 0, L.ptr(out))
        ref = u8.float().div(2 ** (8 - bits)).floor().div(2 ** bits).sub(0.5).add(noise.div(2 ** bits))
        torch.testing.assert_close(out, ref, rtol=0, atol=1e-7)
    # hashed noise: values stay inside the dequantisation bin and are not constant
    L.call("mrssm_normalize_image_u8", L.ptr_any(u8), u8.numel(), 5, None, 123, L.ptr(out))
    lo = u8.float().div(8).floor().div(32).sub(0.5)
    assert bool(((out >= lo) & (out < lo + 1 / 32 + 1e-6)).all())
    assert float((out - lo).std()) > 0.005


def test_pinned_chunk_source_feeds_optimize():
    """The e2e input path: pinned uint8/fp32 host chunks -> async H2D + on-device normalisation -> model.optimize."""
    from mrssm_b200.data import PinnedChunkSource
    from oracle import mrssm_oracle as O
    oc = O.OracleConfig(fusion="MoPoE")
    model, _ = U.build_product(oc, 3, 5, DEV)
    g = torch.Generator().manual_seed(0)
    chunks = []
    for _ in range(2):
        obs = {"image_horizon": torch.randint(0, 256, (5, 3, 3, 64, 64), generator=g, dtype=torch.uint8),
               "pose_quat_v2": torch.randn(5, 3, 3, generator=g)}
        chunks.append((obs, torch.randn(5, 3, 3, generator=g), torch.zeros(5, 3), torch.ones(5, 3, 1)))
    D = PinnedChunkSource(chunks, DEV, bit_depth=5, seed=1)
    first = D.sample(3, 5)
    img = first[0]["image_horizon"]
    lo = chunks[0][0]["image_horizon"].to(DEV).float().div(8).floor().div(32).sub(0.5)
    assert bool(((img >= lo) & (img < lo + 1 / 32 + 1e-6)).all())
    torch.testing.assert_close(first[0]["pose_quat_v2"], chunks[0][0]["pose_quat_v2"].to(DEV))
    D2 = PinnedChunkSource(chunks, DEV, bit_depth=5, seed=1)
    losses = []
    for _ in range(3):
        model.optimize(D2)
        losses.append(float(model.model_loss))
    assert all(l == l and abs(l) < 1e6 for l in losses), losses


def test_pinned_chunk_source_prefetched_s2d_is_the_same_step():
    """bf16 mode: the input pipeline converts frames 1.. to the bf16 space-to-depth form on its copy stream and the conv encoder /
    fused reconstruction loss pick that copy up instead of converting on the critical path — same kernel, same fp32 source, so the
    step is the same with and without it."""
    from mrssm_b200 import ops
    from mrssm_b200.data import PinnedChunkSource
    oc = O.OracleConfig(fusion="MoPoE")
    g = torch.Generator().manual_seed(0)
    chunks = []
    for _ in range(2):
        obs = {"image_horizon": torch.randint(0, 256, (5, 3, 3, 64, 64), generator=g, dtype=torch.uint8),
               "pose_quat_v2": torch.randn(5, 3, 3, generator=g)}
        chunks.append((obs, torch.randn(5, 3, 3, generator=g), torch.zeros(5, 3), torch.ones(5, 3, 1)))
    losses, hits = [], []
    for prefetch_s2d in (True, False):
        model, _ = U.build_product(oc, 3, 5, DEV, bf16=True)
        D = PinnedChunkSource(chunks, DEV, bit_depth=5, seed=1)
        if not prefetch_s2d:
            D._slot(0)[5].clear(), D._slot(1)[5].clear()           # no s2d buffers -> nothing is made or registered
        orig, seen = ops.recall_s2d, []
        ops.recall_s2d = lambda t: (seen.append(orig(t) is not None), orig(t))[1]
        try:
            torch.manual_seed(3)
            run = []
            for _ in range(3):
                model.optimize(D)
                run.append(float(model.model_loss))
        finally:
            ops.recall_s2d = orig
        losses.append(run)
        hits.append(seen)
    assert all(hits[0]) and len(hits[0]) >= 6                       # encoder and loss both found the prefetched copy, every step
    assert hits[1][0] is False                                      # without it the encoder converts itself (and then the loss finds that)
    for a, b in zip(*losses):                                       # (fp32 atomics in the loss reduction: equal up to summation order)
        assert a == pytest.approx(b, rel=1e-5), losses
