"""-m gpu: the tcgen05 rollout (csrc/rollout_tc.cu) against the exact fp32 CUDA-core rollout (same C ABI, same inputs and
noise).  The tensor-core kernel rounds the GEMM operands (weights, x, h, u, state) to bf16 every step and keeps the
recurrent state, accumulation and all gate / fusion math in fp32.  Tolerance (stated): per-element error of every
output <= 3e-2 absolute on O(1) quantities, and the mean absolute error <= 4e-3, after T recurrent steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _params(gen, D, S, H, A, emb_sizes):
    """fc_embed, GRU, prior head, then one head per expert (emb size 0 = no embedding); PyTorch layouts."""
    def lin(o, i):
        k = 1.0 / i ** 0.5
        return [(torch.rand(o, i, device=DEV, generator=gen) * 2 - 1) * k, (torch.rand(o, device=DEV, generator=gen) * 2 - 1) * k]
    w_sa, b_sa = lin(D, S + A)
    kk = 1.0 / D ** 0.5
    gru = [(torch.rand(3 * D, D, device=DEV, generator=gen) * 2 - 1) * kk for _ in range(2)] + \
          [(torch.rand(3 * D, device=DEV, generator=gen) * 2 - 1) * kk for _ in range(2)]
    ps = [w_sa, b_sa] + gru
    for e in [0] + list(emb_sizes):
        ps += lin(H, D + e) + lin(2 * S, H)
    return [p.requires_grad_(False) for p in ps]


def _run(ops, tc, spec, observe, det, ins, embs, params):
    ops.set_bf16_mode(True)
    ops.set_rollout_tc(tc)
    try:
        with torch.no_grad():
            return ops.RolloutFn.apply(spec, observe, det, *ins, *embs, *params)
    finally:
        ops.set_rollout_tc(True)
        ops.set_bf16_mode(False)


CASES = [
    # (D, S, H, A, fusion, emb sizes per expert (0 = none), T, B, det)
    (200, 30, 200, 3, "MoPoE", (0, 1024, 128), 7, 70, False),     # the benchmarked model (config 1-3)
    (200, 30, 200, 3, "PoE", (0, 1024, 128), 5, 64, False),
    (200, 30, 200, 3, "single", (1024,), 6, 130, False),          # single-modal RSSM
    (200, 30, 200, 3, None, (), 9, 33, False),                    # open-loop imagination (config 4)
    (200, 30, 200, 3, "MoPoE", (0, 1024, 128), 4, 16, True),      # det=True (estimate_state)
    (64, 8, 48, 2, "MoPoE", (0, 72), 5, 9, False),                # other sizes: chunk halves 4+4 / 3+3
    (208, 32, 104, 6, "PoE", (0, 64, 64), 3, 65, False),          # maximum D, S; odd chunk counts
]


@pytest.mark.parametrize("case", CASES)
def test_rollout_tc_matches_fp32(case):
    from mrssm_b200 import _lib as L, ops
    D, S, H, A, fusion, emb_sizes, T, B, det = case
    assert L.load().mrssm_rollout_tc_eligible(D, S, H, A, len(emb_sizes)) == 1
    gen = torch.Generator(device=DEV).manual_seed(7 + D + B)
    observe = fusion is not None
    E = len(emb_sizes)
    table = ops.FusionTable(E, S, fusion if observe else "single")
    spec = ops.RolloutSpec(D, S, H, A, ops.RELU, 0.1, table, [e > 0 for e in emb_sizes])
    params = _params(gen, D, S, H, A, emb_sizes)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    nonterm = (torch.rand(T, B, device=DEV, generator=gen) > 0.15).float()
    ins = [rn(B, S), rn(T, B, A), rn(B, D) * 0.5, nonterm, None if det else rn(T, B, S), None if (det or not observe) else rn(T, B, S)]
    embs = [rn(T, B, e) for e in emb_sizes if e > 0]
    ref = _run(ops, False, spec, observe, det, ins, embs, params)
    out = _run(ops, True, spec, observe, det, ins, embs, params)
    assert len(ref) == len(out) == (7 + 2 * E if observe else 4)
    names = ["beliefs", "prior_states", "prior_means", "prior_stds", "post_states", "post_means", "post_stds"] + \
            [f"exp_mean{e}" for e in range(E)] + [f"exp_std{e}" for e in range(E)]
    for n, r, o in zip(names, ref, out):
        assert torch.isfinite(o).all(), n
        err = (o - r).abs()
        assert float(err.max()) <= 3e-2 * max(1.0, float(r.abs().max())), (n, float(err.max()), float(r.abs().max()))
        assert float(err.mean()) <= 4e-3 * max(1.0, float(r.abs().mean())), (n, float(err.mean()))


def test_rollout_tc_stash_matches_fp32():
    """With gradients required the kernel also writes the BPTT stash (x, r, z, n, W_hn h, head hiddens)."""
    from mrssm_b200 import _lib as L, ops
    D, S, H, A, T, B = 200, 30, 200, 3, 5, 40
    gen = torch.Generator(device=DEV).manual_seed(3)
    emb_sizes = (0, 1024, 128)
    table = ops.FusionTable(3, S, "MoPoE")
    spec = ops.RolloutSpec(D, S, H, A, ops.RELU, 0.1, table, [e > 0 for e in emb_sizes])
    params = _params(gen, D, S, H, A, emb_sizes)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    ins = [rn(B, S), rn(T, B, A), (rn(B, D) * 0.5).requires_grad_(True), torch.ones(T, B, device=DEV), rn(T, B, S), rn(T, B, S)]
    embs = [rn(T, B, e) for e in emb_sizes if e > 0]
    stashes = []
    for tc in (False, True):
        ops.set_bf16_mode(True)
        ops.set_rollout_tc(tc)
        try:
            outs = ops.RolloutFn.apply(spec, True, False, *ins, *embs, *params)
            saved = outs[0].grad_fn.saved_tensors
            stashes.append([t.clone() for t in saved[-(5 + 4):]])
        finally:
            ops.set_rollout_tc(True)
            ops.set_bf16_mode(False)
    for i, (r, o) in enumerate(zip(*stashes)):
        err = (o - r).abs()
        assert float(err.max()) <= 3e-2 * max(1.0, float(r.abs().max())), (i, float(err.max()))


BWD_CASES = [
    (200, 30, 200, 3, "MoPoE", (0, 1024, 128), 6, 40, False),
    (200, 30, 200, 3, "PoE", (0, 1024, 128), 4, 70, False),
    (200, 30, 200, 3, "single", (1024,), 5, 33, False),
    (200, 30, 200, 3, None, (), 7, 20, False),
    (64, 8, 48, 2, "MoPoE", (0, 72), 5, 9, False),
    (208, 32, 104, 6, "PoE", (0, 64, 64), 3, 65, False),
]


@pytest.mark.parametrize("case", BWD_CASES)
def test_rollout_tc_bptt_matches_fp32(case):
    """BPTT on the tensor cores (csrc/rollout_tc.cu, backward) against the exact fp32 BPTT kernel: same stash, same upstream
    gradients.  Compared: gradients of the initial state / belief, of the actions and embeddings, and every parameter
    gradient (the deferred weight-gradient GEMMs run on the kernels' pre-activation gradients).  Tolerance (stated): relative
    Frobenius error <= 3e-2 per tensor (bf16 operand rounding of gradients and weights, T recurrent steps)."""
    from mrssm_b200 import _lib as L, ops
    D, S, H, A, fusion, emb_sizes, T, B, det = case
    gen = torch.Generator(device=DEV).manual_seed(11 + D + B)
    observe = fusion is not None
    E = len(emb_sizes)
    table = ops.FusionTable(E, S, fusion if observe else "single")
    spec = ops.RolloutSpec(D, S, H, A, ops.RELU, 0.1, table, [e > 0 for e in emb_sizes])
    params = [p.requires_grad_(True) for p in _params(gen, D, S, H, A, emb_sizes)]
    rn = lambda *s: torch.randn(*s, device=DEV, generator=gen)
    nonterm = (torch.rand(T, B, device=DEV, generator=gen) > 0.15).float()
    ins = [rn(B, S).requires_grad_(True), rn(T, B, A).requires_grad_(True), (rn(B, D) * 0.5).requires_grad_(True), nonterm,
           rn(T, B, S), rn(T, B, S) if observe else None]
    embs = [rn(T, B, e).requires_grad_(True) for e in emb_sizes if e > 0]
    n_out = 7 + 2 * E if observe else 4
    gouts = None
    results = []
    for tc in (False, True):
        for t in ins[:3] + embs + params:
            t.grad = None
        ops.set_bf16_mode(True)
        ops.set_rollout_tc(False)                       # identical (fp32-kernel) forward and stash for both runs
        try:
            outs = ops.RolloutFn.apply(spec, observe, det, *ins, *embs, *params)
            assert len(outs) == n_out
            if gouts is None:
                gouts = [rn(*o.shape) / o.shape[-1] ** 0.5 for o in outs]
            ops.set_rollout_tc(tc)
            torch.autograd.backward(outs, gouts)
        finally:
            ops.set_rollout_tc(True)
            ops.set_bf16_mode(False)
        results.append([t.grad.clone() for t in ins[:3] + embs + params])
    names = ["g_prev_state", "g_actions", "g_prev_belief"] + [f"g_emb{i}" for i in range(len(embs))] + [f"g_param{i}" for i in range(len(params))]
    for n, r, o in zip(names, *results):
        assert torch.isfinite(o).all(), n
        rel = float((o - r).norm() / (r.norm() + 1e-12))
        assert rel <= 3e-2, (n, rel, float(r.norm()))


@pytest.mark.parametrize("rows", [16, 8])
def test_rollout_tc_both_cta_shapes(rows):
    """64 sequences per CTA (both rows of every TMEM fragment carry a sequence; chosen automatically for B >= 4700) and 32
    per CTA give the same results: forward outputs and BPTT gradients against the fp32 kernels, with the shape forced."""
    from mrssm_b200 import _lib as L
    L.call_host("mrssm_rollout_tc_set_rows", rows)
    try:
        test_rollout_tc_matches_fp32(CASES[0])
        test_rollout_tc_matches_fp32(CASES[2])
        test_rollout_tc_bptt_matches_fp32(BWD_CASES[0])
        test_rollout_tc_bptt_matches_fp32(BWD_CASES[3])
    finally:
        L.call_host("mrssm_rollout_tc_set_rows", 0)
