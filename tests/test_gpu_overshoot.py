"""-m gpu: latent overshooting (reference base/algo.py:111-148, MRSSM_MoPoE/algo.py:69-108) through the C ABI.

1. the gather and masked-KL kernels against the reference's own construction (F.pad + torch.cat + kl_divergence) restated
   with torch ops on the same device tensors: bit-exact for the gather (pure data movement), rtol 1e-4 for the KL value
   and its gradients (fp32 both sides; only the summation order differs);
2. whole training steps with overshooting on against the CPU oracle (rtol 1e-3, the north star's figure) — the fixtures
   from the unmodified reference are covered by test_gpu_parity.py::test_train_step_matches_reference_fixture;
3. bf16 tensor-core mode: within the stated bf16 tolerances of test_gpu_parity.py."""
import pytest
import torch
import torch.nn.functional as F
from torch.distributions import Normal, kl_divergence

from tests import parity_util as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3


def _reference_layout(x, T, OD, value=0.0):
    """[T', B, ...] -> [OD, (T-2)B, ...]: slices x[t:d] of every start t = 1..T-2, padded to OD steps, side by side."""
    cols = []
    for t in range(1, T - 1):
        d = min(t + OD, T - 1)
        pad = (0, 0) * (x.dim() - 1) + (0, t - d + OD)
        cols.append(F.pad(x[t:d], pad, value=value))
    return torch.cat(cols, 1)


@pytest.mark.parametrize("T,B,OD,targets", [(6, 5, 3, "post"), (9, 3, 4, "subset"), (4, 7, 5, "all"), (3, 2, 1, "post")])
def test_gather_and_masked_kl_match_reference_construction(T, B, OD, targets):
    from mrssm_b200 import ops
    S, A, free = 30, 3, 0.7
    g = torch.Generator(device=DEV).manual_seed(5 + T)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    actions, rewards = rn(T, B, A), rn(T, B)
    nonterm = (torch.rand(T, B, 1, device=DEV, generator=g) > 0.2).float()
    spec = ops.OvershootSpec(T, B, S, A, OD, free)
    act_o, nt_o, rw_o, mk_o = ops.overshoot_gather(spec, actions, nonterm, rewards, want_mask=True)
    assert torch.equal(act_o, _reference_layout(actions, T, OD))
    assert torch.equal(nt_o, _reference_layout(nonterm, T, OD))
    assert torch.equal(rw_o, _reference_layout(rewards, T, OD))
    ones = torch.ones(T, B, device=DEV)
    assert torch.equal(mk_o, _reference_layout(ones, T, OD))

    N = (T - 2) * B
    pm = rn(OD, N, S).requires_grad_(True)
    ps = (torch.rand(OD, N, S, device=DEV, generator=g) + 0.3).requires_grad_(True)
    E = 3
    ex_m = [rn(T - 1, B, S) for _ in range(E)]
    ex_s = [torch.rand(T - 1, B, S, device=DEV, generator=g) + 0.4 for _ in range(E)]
    if targets == "post":
        n_experts, mask, tg = 0, 0, [ex_m[0], ex_s[0]]
        qm, qs = ex_m[0], ex_s[0]
    else:
        mask = 0b101 if targets == "subset" else 0b111
        n_experts, tg = E, ex_m + ex_s
        sel = [e for e in range(E) if mask >> e & 1]
        prec = sum(1.0 / ex_s[e] for e in sel)                       # encoder.py:50-55 (precision = 1/std)
        qm, qs = sum(ex_m[e] / ex_s[e] for e in sel) / prec, 1.0 / prec
    scale = 0.37
    kspec = ops.OvershootSpec(T, B, S, A, OD, free, scale=scale, n_experts=n_experts, mask=mask)
    out = ops.OvershootKlFn.apply(kspec, pm, ps, *tg)
    out.backward()
    got = (out.detach().clone(), pm.grad.clone(), ps.grad.clone())
    pm.grad = ps.grad = None
    # T-1 posterior steps: index t_+1..d_ = t..d-1, the same slice bounds as the [T] tensors
    qm_o, qs_o = _reference_layout(qm, T, OD), _reference_layout(qs, T, OD, value=1.0)
    seq_mask = _reference_layout(torch.ones(T - 1, B, S, device=DEV), T, OD)
    ref = scale * torch.max((kl_divergence(Normal(qm_o, qs_o), Normal(pm, ps)) * seq_mask).sum(2),
                            torch.full((1,), free, device=DEV)).mean((0, 1))
    ref.backward()
    torch.testing.assert_close(got[0], ref.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(got[1], pm.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(got[2], ps.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("fusion,kw", [
    ("MoPoE", dict(overshooting_distance=3, overshooting_kl_beta=0.5, free_nats=0.5)),
    ("PoE", dict(overshooting_distance=2, overshooting_kl_beta=1.0, predict_reward=True, overshooting_reward_scale=0.5)),
    ("NN", dict(overshooting_distance=4, overshooting_kl_beta=0.3, free_nats=0.2)),
    ("single", dict(overshooting_distance=8, overshooting_kl_beta=1.0, free_nats=0.1)),      # distance > chunk: all runs padded
    ("MoPoE", dict(overshooting_distance=2, overshooting_kl_beta=0.5, predict_reward=False, overshooting_reward_scale=1.0)),
])
def test_train_step_with_overshooting_matches_oracle(fusion, kw):
    out = U.run_train_parity(fusion, B=4, T=6, steps=2, device=DEV, rtol=RTOL, **kw)
    assert out["worst_grad_err"] < 2 * RTOL


def test_overshooting_bf16_mode_within_stated_tolerance():
    rep = U.run_train_parity_bf16("MoPoE", B=4, T=6, steps=2, device=DEV, overshooting_distance=3, overshooting_kl_beta=0.5,
                                  free_nats=0.5)
    assert rep["state_err"] <= 2e-2 and rep["loss_rel"] <= 2e-2, rep
    assert rep["grad_rel_fro"] <= 5e-2 and rep["gnorm_rel"] <= 5e-2, rep
