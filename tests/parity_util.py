"""Shared harness: run the CUDA product and the CPU oracle on identical weights, data and noise."""
import torch

from oracle import mrssm_oracle as O


def to_attr(d):
    from mrssm_b200.config import to_attr as f
    return f(d)


def oracle_cfg(fusion, **kw):
    if fusion == "single":
        kw.setdefault("names_enc", ("image_horizon",))
        kw.setdefault("names_rec", ("image_horizon",))
    return O.OracleConfig(fusion=fusion, **kw)


def product_cfg(oc, B, T, device, bf16=False):
    from mrssm_b200.config import hot_path_config
    cfg = _product_cfg(oc, B, T, device)
    cfg.train.use_amp = bool(bf16)
    return cfg


def _product_cfg(oc, B, T, device):
    from mrssm_b200.config import hot_path_config
    cfg = _hot_cfg(oc, B, T, device)
    # modality set and shapes exactly as the oracle configuration names them (128x128 images, other vector names)
    cfg.rssm.observation_names_enc = list(oc.names_enc)
    cfg.rssm.observation_names_rec = list(oc.names_rec)
    cfg.env.observation_shapes = to_attr({k: list(v) for k, v in oc.observation_shapes.items()})
    cfg.env.action_size = oc.action_size
    return cfg


def _hot_cfg(oc, B, T, device):
    from mrssm_b200.config import hot_path_config
    return hot_path_config(fusion=oc.fusion, batch_size=B, chunk_size=T, device=device,
                           belief_size=oc.belief_size, state_size=oc.state_size, hidden_size=oc.hidden_size,
                           free_nats=oc.free_nats, kl_balancing_alpha=oc.kl_balancing_alpha,
                           global_kl_beta=oc.global_kl_beta, kl_beta=oc.kl_beta, grad_clip_norm=oc.grad_clip_norm,
                           model_learning_rate=oc.lr, adam_epsilon=oc.adam_eps, predict_reward=oc.predict_reward,
                           overshooting_distance=oc.overshooting_distance, overshooting_kl_beta=oc.overshooting_kl_beta,
                           overshooting_reward_scale=oc.overshooting_reward_scale,
                           worldmodel_LogProbLoss=oc.worldmodel_LogProbLoss, learning_rate_schedule=oc.learning_rate_schedule,
                           embedding_size=dict(oc.embedding_size), normalization=oc.normalization)


def unflatten(flat):
    out = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


def load_params(model, P, oc):
    if oc.multimodal:
        model.load_state_dict(unflatten(P))
    else:
        torch.nn.Module.load_state_dict(model, P)


def named_params(model, oc):
    """flat oracle key -> nn.Parameter of the product model."""
    if oc.multimodal:
        sd = model.get_state_dict()
        sd.pop("model_optimizer")
        by_ptr = {p.data_ptr(): p for p in model.param_list}
        return {k: by_ptr[v.data_ptr()] for k, v in O.flatten_state(sd).items() if not O.is_buffer(k)}
    return dict(model.named_parameters())


class FakeD:
    """Replay-buffer stand-in with the reference's D.sample contract (memory.py:212-222)."""

    def __init__(self, batch, device):
        self.b = batch
        self.device = device

    def sample(self, n, L):
        d = self.device
        return ({k: v.to(d) for k, v in self.b["obs"].items()}, self.b["actions"].to(d),
                self.b["rewards"].to(d), self.b["nonterminals"].to(d))


def named_buffers(model, oc):
    """flat oracle key -> BatchNorm / InstanceNorm buffer of the product model."""
    if oc.multimodal:
        sd = model.get_state_dict()
        sd.pop("model_optimizer", None)
        return {k: v for k, v in O.flatten_state(sd).items() if O.is_buffer(k)}
    return {k: v for k, v in torch.nn.Module.state_dict(model).items() if O.is_buffer(k)}


def build_product(oc, B, T, device, seed=0, bf16=False):
    from algos.MRSSM.MRSSM.algo import build_RSSM
    model = build_RSSM(product_cfg(oc, B, T, device, bf16), torch.device(device))
    P = O.make_params(oc, seed=seed)
    load_params(model, P, oc)
    return model, P


def product_step(model, oc, batch, noise, device):
    from mrssm_b200.noise import FixedNoise
    dev = torch.device(device)
    streams = dict(prior=[noise["eps_prior"].to(dev)] + [e.to(dev) for e in noise.get("eps_over", [])],   # one per open-loop run
                   post=noise["eps_post"].to(dev))
    if oc.fusion in ("PoE", "MoPoE"):
        streams["dec"] = noise["eps_dec"].to(dev)
    cap = {}
    orig = model.estimate_state

    def es(*a, **k):
        cap["states"] = orig(*a, **k)
        return cap["states"]

    model.estimate_state = es
    try:
        with FixedNoise(**streams):
            model.optimize(FakeD(batch, dev))
    finally:
        model.estimate_state = orig
    return cap["states"]


def assert_states_close(st, ref, rtol, atol):
    for k, v in ref.items():
        if isinstance(v, dict):
            for n in v:
                torch.testing.assert_close(st[k][n].detach().cpu(), v[n], rtol=rtol, atol=atol, msg=lambda m: f"{k}[{n}]: {m}")
        elif v is not None:
            torch.testing.assert_close(st[k].detach().cpu(), v, rtol=rtol, atol=atol, msg=lambda m: f"{k}: {m}")


def run_train_parity(fusion, B, T, steps, device, rtol=1e-3, atol=2e-5, **cfg_kw):
    """Product vs oracle: states, losses, every gradient tensor, every updated parameter."""
    oc = oracle_cfg(fusion, **cfg_kw)
    model, P = build_product(oc, B, T, device)
    named = named_params(model, oc)
    opt = {}
    worst = {}
    for s in range(steps):
        batch, noise = O.synthetic_batch(oc, B, T, seed=1234 + s)
        ref = O.train_step(P, opt, oc, batch, noise)
        st = product_step(model, oc, batch, noise, device)
        assert_states_close(st, {k: v for k, v in ref["states"].items()}, rtol, atol)
        info = {k: float(v) for k, v in model.loss_info.items()}
        for k, v in ref["loss_info"].items():
            assert abs(info[k] - v) <= rtol * abs(v) + 1e-5, (k, info[k], v)
        gn = float(model.model_optimizer.grad_norm)
        assert abs(gn - ref["grad_norm"]) <= rtol * ref["grad_norm"], (gn, ref["grad_norm"])
        gmax = max(float(g.abs().max()) for g in ref["grads"].values())
        for k, g in ref["grads"].items():
            mine = named[k].grad.detach().cpu()
            scale = max(float(g.abs().max()), 1e-3 * gmax)
            err = float((mine - g).abs().max()) / scale
            worst[k] = max(worst.get(k, 0.0), err)
            assert err <= 2 * rtol, f"grad {k}: max err {err:.3e} (relative to max|g|)"
        for k in P:
            if k not in ref["grads"]:
                assert float(named[k].grad.abs().max()) == 0.0, f"{k} should get no gradient"
            torch.testing.assert_close(named[k].detach().cpu(), P[k], rtol=rtol, atol=1e-5, msg=lambda m: f"param {k}: {m}")
    return {"worst_grad_err": max(worst.values()), "model_loss": float(model.model_loss), "steps": steps}


def run_train_parity_bf16(fusion, B, T, steps, device, **cfg_kw):
    """bf16 tensor-core mode vs the fp32 oracle: reports errors instead of asserting per tensor; the caller
    applies the stated bf16 tolerances."""
    oc = oracle_cfg(fusion, **cfg_kw)
    model, P = build_product(oc, B, T, device, bf16=True)
    named = named_params(model, oc)
    opt = {}
    rep = dict(state_err=0.0, loss_rel=0.0, grad_rel_fro=0.0, gnorm_rel=0.0)
    for s in range(steps):
        batch, noise = O.synthetic_batch(oc, B, T, seed=1234 + s)
        ref = O.train_step(P, opt, oc, batch, noise)
        st = product_step(model, oc, batch, noise, device)
        for k, v in ref["states"].items():
            items = v.items() if isinstance(v, dict) else [(None, v)]
            for n, t in items:
                if t is None:
                    continue
                mine = (st[k][n] if n is not None else st[k]).detach().cpu()
                rep["state_err"] = max(rep["state_err"], float((mine - t.detach()).abs().max() / (t.abs().max() + 1e-6)))
        info = {k: float(v) for k, v in model.loss_info.items()}
        for k, v in ref["loss_info"].items():
            if abs(v) > 1e-6:
                rep["loss_rel"] = max(rep["loss_rel"], abs(info[k] - v) / abs(v))
        gn = float(model.model_optimizer.grad_norm)
        rep["gnorm_rel"] = max(rep["gnorm_rel"], abs(gn - ref["grad_norm"]) / ref["grad_norm"])
        num = sum(float(((named[k].grad.detach().cpu() - g) ** 2).sum()) for k, g in ref["grads"].items())
        den = sum(float((g ** 2).sum()) for g in ref["grads"].values())
        rep["grad_rel_fro"] = max(rep["grad_rel_fro"], (num / den) ** 0.5)
        # keep the oracle on the product's trajectory so step 2 compares like with like
        bufs = named_buffers(model, oc) if any(O.is_buffer(k) for k in P) else {}
        for k in P:
            P[k].copy_((bufs[k] if O.is_buffer(k) else named[k]).detach().cpu())
    return rep
