"""Pin oracle/mrssm_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import mrssm_oracle as O

TRAIN = ["mopoe", "poe", "nn", "single", "mopoe_clip", "poe_noalpha", "mopoe_reward", "mopoe_over", "poe_over", "single_over",
         "mopoe_bn", "single_bn", "mopoe_sound", "mopoe_sound_bn", "mopoe_logprob", "single_logprob", "mopoe_lrramp",
         "mopoe_emb512", "mopoe_img128", "mopoe_img84", "mopoe_img256", "single_img84_bn"]


def _cfg(meta):
    return O.OracleConfig(**meta["cfg"])


def _close(a, b, rtol=2e-4, atol=2e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def _cmp_states(st, ref, **tol):
    for k, v in ref.items():
        if isinstance(v, dict):
            for n in v:
                _close(st[k][n], v[n], **tol)
        else:
            _close(st[k], v, **tol)


@pytest.mark.parametrize("name", TRAIN)
def test_train_steps_match_reference(name, golden_dir):
    rec = torch.load(os.path.join(golden_dir, f"train_{name}.pt"), weights_only=False)
    cfg = _cfg(rec["meta"])
    B, T = rec["meta"]["B"], rec["meta"]["T"]
    P = O.make_params(cfg, seed=rec["meta"]["param_seed"])
    assert abs(sum(v.double().abs().sum() for k, v in P.items() if not O.is_buffer(k)) - rec["param_checksum"]) \
        < 1e-6 * rec["param_checksum"]
    opt = {}
    for step in rec["steps"]:
        batch, noise = O.synthetic_batch(cfg, B, T, seed=step["data_seed"])
        chk = float(sum(v.double().abs().sum() for v in batch["obs"].values()) + batch["actions"].double().abs().sum())
        assert abs(chk - step["input_checksum"]) < 1e-9 * chk
        out = O.train_step(P, opt, cfg, batch, noise)
        _cmp_states(out["states"], step["states"])
        for k, v in step["loss_info"].items():
            assert out["loss_info"][k] == pytest.approx(v, rel=2e-5, abs=1e-6), k
        assert out["model_loss"] == pytest.approx(step["model_loss"], rel=2e-5)
        assert out["grad_norm"] == pytest.approx(step["grad_norm"], rel=1e-4)
        assert sorted(k for k in P if k not in out["grads"] and not O.is_buffer(k)) == step["grad_none"]
        for k, s in step["grads"].items():
            g = out["grads"][k].reshape(-1)
            assert float(g.double().norm()) == pytest.approx(s["norm"], rel=2e-4, abs=1e-7), k
            _close(g[s["idx"]], s["val"], rtol=1e-3, atol=1e-5 * max(1.0, s["norm"]))
        for k, s in step["params_after"].items():
            p = P[k].reshape(-1)
            _close(p[s["idx"]], s["val"], rtol=1e-4, atol=2e-7 if cfg.lr < 1e-4 else 2e-6)
            assert float(p.double().norm()) == pytest.approx(s["norm"], rel=1e-5), k
        for k, b in step.get("buffers_after", {}).items():          # BatchNorm running statistics / batch counters
            if b.dtype == torch.long:
                assert int(P[k]) == int(b), k
            else:
                _close(P[k], b, rtol=1e-4, atol=1e-6)
    if "eval" in rec:                                               # eval mode: the running statistics normalise
        ev = rec["eval"]
        batch, _ = O.synthetic_batch(cfg, B, T, seed=ev["data_seed"])
        tgt = {n: batch["obs"][n][1:] for n in cfg.names_enc}
        with torch.no_grad():
            st = O.estimate_state(P, cfg, tgt, batch["actions"][:-1], batch["nonterminals"][:-1], None, None, det=True,
                                  train=False)
            _cmp_states(st, ev["states"], rtol=2e-4, atol=2e-5)
            recon = O.decode(P, cfg, st["beliefs"], st["posterior_states"], train=False)
        for n, s in ev["recon"].items():
            r = recon[n].reshape(-1)
            _close(r[s["idx"]], s["val"], rtol=1e-3, atol=1e-5)
            assert float(r.double().norm()) == pytest.approx(s["norm"], rel=1e-4), n


@pytest.mark.parametrize("name", ["mopoe", "single"])
def test_det_and_imagination_match_reference(name, golden_dir):
    rec = torch.load(os.path.join(golden_dir, f"infer_{name}.pt"), weights_only=False)
    cfg = _cfg(rec["meta"])
    B, T, H = rec["meta"]["B"], rec["meta"]["T"], rec["meta"]["H"]
    P = O.make_params(cfg, seed=0)
    batch, _ = O.synthetic_batch(cfg, B, T, seed=99)
    tgt = {n: batch["obs"][n][1:] for n in cfg.names_enc}
    with torch.no_grad():
        st = O.estimate_state(P, cfg, tgt, batch["actions"][:-1], batch["nonterminals"][:-1], None, None, det=True)
        _cmp_states(st, rec["states_det"])
        g = torch.Generator().manual_seed(7)
        acts = torch.randn(H, B, cfg.action_size, generator=g)
        eps = torch.randn(H, B, cfg.state_size, generator=g)
        h0, s0 = st["beliefs"][-1], st["posterior_states"][-1]
        im = O.rollout(P, cfg, s0, acts, h0, None, None, eps, None)
        imd = O.rollout(P, cfg, s0, acts, h0, None, None, None, None, det=True)
    keys = ["beliefs", "prior_states", "prior_means", "prior_std_devs"]
    for k, r, rd in zip(keys, rec["imagine"], rec["imagine_det"]):
        _close(im[k], r)
        _close(imd[k], rd)


def test_known_answers():
    """KATs for the quirks a clean re-derivation gets wrong."""
    # Q3: PoE precision is 1/sigma
    mu, sd = O.poe(torch.tensor([[1.0], [3.0]]), torch.tensor([[1.0], [0.5]]))
    assert float(mu) == pytest.approx((1 * 1 + 3 * 2) / 3) and float(sd) == pytest.approx(1 / 3)
    # Q8: slice table
    assert O.mopoe_slices(30, 4) == [(0, 7), (7, 14), (14, 21), (21, 30)]
    assert O.mopoe_slices(64, 4) == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert O.mopoe_slices(128, 4) == [(0, 32), (32, 64), (64, 96), (96, 128)]
    assert O.subset_table(2) == [(), (1,), (2,), (1, 2)]
    assert O.subset_table(3) == [(), (1,), (2,), (3,), (1, 2), (1, 3), (2, 3), (1, 2, 3)]
    # Gaussian KL against torch.distributions
    from torch.distributions import Normal, kl_divergence
    a, b, c, d = torch.randn(5), torch.rand(5) + 0.1, torch.randn(5), torch.rand(5) + 0.1
    torch.testing.assert_close(O.kl_normal(a, b, c, d), kl_divergence(Normal(a, b), Normal(c, d)))
    # GRU gate order / b_hn placement against nn.GRUCell
    cell = torch.nn.GRUCell(6, 6)
    x, h = torch.randn(3, 6), torch.randn(3, 6)
    torch.testing.assert_close(O.gru_cell(x, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh), cell(x, h))
    # clip coefficient and Adam against torch
    p = torch.nn.Parameter(torch.randn(7))
    p.grad = torch.randn(7) * 10
    q, g = p.detach().clone(), p.grad.clone()
    tot, coef = O.clip_coef([g], 1.0)
    n = torch.nn.utils.clip_grad_norm_([p], 1.0)
    assert float(tot) == pytest.approx(float(n))
    opt = torch.optim.Adam([p], lr=1e-3, eps=1e-7)
    m, v = torch.zeros(7), torch.zeros(7)
    for step in (1, 2):
        opt.step()
        O.adam_update(q, g * coef, m, v, step, 1e-3, 1e-7)
    torch.testing.assert_close(q, p.detach())
