"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only (the reference is not present on the GPU box):
    python tests/golden/make_golden.py
The fixtures pin oracle/mrssm_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA
path.  Inputs and weights are NOT stored: they are regenerated from seeds by
oracle.mrssm_oracle.{make_params,synthetic_batch}; checksums of both are stored to detect drift.
Noise is injected by replacing torch.distributions.normal._standard_normal with a FIFO
(draw order, SURVEY Q4: per step prior (B,S) then posterior (1,B,S); afterwards one (T-1,B,S)).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import mrssm_oracle as O  # noqa: E402

import torch.distributions.normal as tdn  # noqa: E402
from algos.MRSSM.MRSSM.algo import build_RSSM  # noqa: E402  (the reference)


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = dict.__setitem__


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


def ref_cfg(oc: O.OracleConfig, B, T):
    return to_attr(dict(
        main=dict(device="cpu", wandb=False),
        env=dict(observation_shapes={k: list(v) for k, v in oc.observation_shapes.items()},
                 action_size=oc.action_size),
        train=dict(batch_size=B, chunk_size=T, use_amp=False),
        rssm=dict(
            observation_names_enc=list(oc.names_enc), observation_names_rec=list(oc.names_rec),
            predict_reward=oc.predict_reward, multimodal=oc.multimodal,
            multimodal_params=dict(fusion_method=oc.fusion, expert_dist="q(st|ht,ot)"),
            activation_function=dict(cnn="relu", dense=oc.act_dense, fusion="relu"),
            embedding_size=dict(oc.embedding_size), hidden_size=oc.hidden_size,
            belief_size=oc.belief_size, state_size=oc.state_size, normalization=oc.normalization,
            worldmodel_LogProbLoss=oc.worldmodel_LogProbLoss, overshooting_distance=oc.overshooting_distance,
            overshooting_kl_beta=oc.overshooting_kl_beta,
            overshooting_reward_scale=oc.overshooting_reward_scale, global_kl_beta=oc.global_kl_beta, free_nats=oc.free_nats,
            kl_beta=oc.kl_beta, kl_balancing_alpha=oc.kl_balancing_alpha, learning_rate_schedule=oc.learning_rate_schedule,
            adam_epsilon=oc.adam_eps, grad_clip_norm=oc.grad_clip_norm, model_learning_rate=oc.lr)))


def unflatten(flat):
    out = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


class NoiseFifo:
    def __init__(self):
        self.q = []
        self.orig = tdn._standard_normal

    def push(self, t):
        self.q.append(t)

    def __call__(self, shape, dtype, device):
        t = self.q.pop(0)
        assert t.numel() == torch.Size(shape).numel(), (t.shape, shape)
        return t.reshape(shape).to(dtype)

    def __enter__(self):
        tdn._standard_normal = self
        return self

    def __exit__(self, *a):
        tdn._standard_normal = self.orig
        assert not self.q, "unused noise"


def queue_train_noise(fifo, noise, fusion):
    Tm1 = noise["eps_prior"].shape[0]
    for t in range(Tm1):
        fifo.push(noise["eps_prior"][t])
        fifo.push(noise["eps_post"][t])
    if fusion in ("PoE", "MoPoE"):
        fifo.push(noise["eps_dec"])
    if "eps_over" in noise:                  # latent overshooting: (PoE: one more decoder-latent draw, unused) then, per
        if "eps_dec2" in noise:              # imagination rollout, one (N,S) draw per step
            fifo.push(noise["eps_dec2"])
        for eps in noise["eps_over"]:
            for t in range(eps.shape[0]):
                fifo.push(eps[t])


def summarize(t, n=24):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return dict(norm=float(f.norm()), sum=float(f.sum()), idx=idx, val=f[idx].float())


def load_into_ref(model, P, oc):
    if oc.multimodal:
        nested = unflatten(P)
        nested["model_optimizer"] = model.model_optimizer.state_dict()
        model.load_state_dict(nested)
    else:
        torch.nn.Module.load_state_dict(model, P)


def named_ref_params(model, oc):
    if oc.multimodal:
        sd = model.get_state_dict()
        sd.pop("model_optimizer")
        # state_dict tensors alias the parameters; map by data_ptr to find .grad
        by_ptr = {p.data_ptr(): p for p in model.param_list}
        return {k: by_ptr[v.data_ptr()] for k, v in O.flatten_state(sd).items() if not O.is_buffer(k)}
    return dict(model.named_parameters())


def ref_buffers(model, oc):
    """BatchNorm running statistics of the reference model, keyed like the oracle's flat dict."""
    sd = model.get_state_dict()
    sd.pop("model_optimizer", None)
    return {k: v.detach().clone() for k, v in O.flatten_state(sd).items() if O.is_buffer(k)}


class FakeD:
    def __init__(self, batch):
        self.b = batch

    def sample(self, n, L):
        return ({k: v.clone() for k, v in self.b["obs"].items()}, self.b["actions"].clone(),
                self.b["rewards"].clone(), self.b["nonterminals"].clone())


def states_to_plain(st):
    out = {}
    for k, v in st.items():
        if isinstance(v, dict):
            out[k] = {n: t.detach().clone() for n, t in v.items()}
        elif v is not None:
            out[k] = v.detach().clone()
    return out


ONLY = set()


def gen_train(name, oc, B=4, T=6, steps=2):
    if ONLY and name not in ONLY:
        return
    torch.manual_seed(0)
    model = build_RSSM(ref_cfg(oc, B, T), torch.device("cpu"))
    P = O.make_params(oc, seed=0)
    load_into_ref(model, P, oc)
    model._init_optimizer()
    named = named_ref_params(model, oc)
    rec = dict(meta=dict(name=name, B=B, T=T, cfg=oc.__dict__.copy(), param_seed=0),
               param_checksum=float(sum(v.double().abs().sum() for k, v in P.items() if not O.is_buffer(k))), steps=[])
    cap = {}
    has_buffers = any(O.is_buffer(k) for k in P)
    orig_es = model.estimate_state
    orig_ml = model._get_model_loss

    def es(*a, **k):
        cap["states"] = orig_es(*a, **k)
        return cap["states"]

    def ml(*a, **k):
        loss, info = orig_ml(*a, **k)
        cap["loss"], cap["info"] = float(loss), dict(info)
        return loss, info

    model.estimate_state, model._get_model_loss = es, ml
    orig_clip = torch.nn.utils.clip_grad_norm_

    def clip(params, max_norm, norm_type=2):
        params = list(params)
        cap["grads"] = {k: p.grad.detach().clone() for k, p in named.items() if p.grad is not None}
        cap["grad_none"] = sorted(k for k, p in named.items() if p.grad is None)
        n = orig_clip(params, max_norm, norm_type=norm_type)
        cap["grad_norm"] = float(n)
        return n

    torch.nn.utils.clip_grad_norm_ = clip
    try:
        for s in range(steps):
            batch, noise = O.synthetic_batch(oc, B, T, seed=1234 + s)
            with NoiseFifo() as fifo:
                queue_train_noise(fifo, noise, oc.fusion)
                model.optimize(FakeD(batch))
            rec["steps"].append(dict(
                data_seed=1234 + s,
                input_checksum=float(sum(v.double().abs().sum() for v in batch["obs"].values())
                                     + batch["actions"].double().abs().sum()),
                states=states_to_plain(cap["states"]), loss_info=cap["info"], model_loss=cap["loss"],
                grad_norm=cap["grad_norm"], grad_none=cap["grad_none"],
                grads={k: summarize(g) for k, g in cap["grads"].items()},
                params_after={k: summarize(p) for k, p in named.items()}))
            if has_buffers:
                rec["steps"][-1]["buffers_after"] = ref_buffers(model, oc)
        if has_buffers:                         # eval mode (running statistics) after the training steps: validation's forward
            model.estimate_state = orig_es
            model.eval()
            batch, _ = O.synthetic_batch(oc, B, T, seed=77)
            tgt = {n: batch["obs"][n][1:] for n in oc.names_enc}
            with torch.no_grad():
                st = model.estimate_state(tgt, batch["actions"][:-1], None, batch["nonterminals"][:-1], det=True)
                out = model.observation_model(h_t=st["beliefs"], s_t=st["posterior_states"])
                rec["eval"] = dict(data_seed=77, states=states_to_plain(st),
                                   recon={n: summarize((out[n] if oc.multimodal else out)["loc"], n=64) for n in oc.names_rec})
            model.train()
    finally:
        torch.nn.utils.clip_grad_norm_ = orig_clip
    torch.save(rec, os.path.join(HERE, f"train_{name}.pt"))
    print(name, "loss", [s["model_loss"] for s in rec["steps"]], "gnorm", [s["grad_norm"] for s in rec["steps"]])


def gen_infer(name, oc, B=3, T=5, H=7):
    """estimate_state(det=True) and the open-loop imagination call of check_model.ipynb cell 55
    (transition_model(s, actions[H], h, None, None), stochastic and det)."""
    if ONLY and name not in ONLY:
        return
    torch.manual_seed(0)
    model = build_RSSM(ref_cfg(oc, B, T), torch.device("cpu"))
    P = O.make_params(oc, seed=0)
    load_into_ref(model, P, oc)
    batch, noise = O.synthetic_batch(oc, B, T, seed=99)
    tgt = {n: batch["obs"][n][1:] for n in oc.names_enc}
    if not oc.multimodal:
        pass
    with torch.no_grad():
        st_det = model.estimate_state(tgt, batch["actions"][:-1], None, batch["nonterminals"][:-1], det=True)
        g = torch.Generator().manual_seed(7)
        acts = torch.randn(H, B, oc.action_size, generator=g)
        eps = torch.randn(H, B, oc.state_size, generator=g)
        h0, s0 = st_det["beliefs"][-1], st_det["posterior_states"][-1]
        with NoiseFifo() as fifo:
            for t in range(H):
                fifo.push(eps[t])
            im = model.transition_model(s0, acts, h0, None, None)
        im_det = model.transition_model(s0, acts, h0, None, None, det=True)
    rec = dict(meta=dict(name=name, B=B, T=T, H=H, cfg=oc.__dict__.copy()),
               states_det=states_to_plain(st_det),
               imagine=[t.clone() for t in im], imagine_det=[t.clone() for t in im_det])
    torch.save(rec, os.path.join(HERE, f"infer_{name}.pt"))
    print("infer", name, "ok")


if __name__ == "__main__":
    torch.set_num_threads(8)
    ONLY.update(sys.argv[1:])            # python make_golden.py [name ...]: regenerate only the named fixtures
    gen_train("mopoe", O.OracleConfig(fusion="MoPoE"))
    gen_train("poe", O.OracleConfig(fusion="PoE"))
    gen_train("nn", O.OracleConfig(fusion="NN"))
    gen_train("single", O.OracleConfig(fusion="single", names_enc=("image_horizon",), names_rec=("image_horizon",)))
    gen_train("mopoe_clip", O.OracleConfig(fusion="MoPoE", grad_clip_norm=0.5, kl_balancing_alpha=None))
    gen_train("poe_noalpha", O.OracleConfig(fusion="PoE", kl_balancing_alpha=None, global_kl_beta=0.0, free_nats=0.5))
    gen_train("mopoe_reward", O.OracleConfig(fusion="MoPoE", predict_reward=True))
    gen_train("mopoe_over", O.OracleConfig(fusion="MoPoE", overshooting_distance=3, overshooting_kl_beta=0.5))
    gen_train("poe_over", O.OracleConfig(fusion="PoE", overshooting_distance=2, overshooting_kl_beta=1.0, predict_reward=True,
                                         overshooting_reward_scale=0.5))
    gen_train("single_over", O.OracleConfig(fusion="single", names_enc=("image_horizon",), names_rec=("image_horizon",),
                                            overshooting_distance=4, overshooting_kl_beta=0.25))
    # BatchNorm over 20 samples amplifies parameter differences a few hundred times per layer (1/sqrt(var + 1e-5) on nearly
    # constant channels); a small learning rate keeps Adam's sign-like first update from turning fp32 summation-order
    # noise into visible step-2 differences
    gen_train("mopoe_bn", O.OracleConfig(fusion="MoPoE", normalization="BatchNorm", lr=1e-5))
    gen_train("single_bn", O.OracleConfig(fusion="single", names_enc=("image_horizon",), names_rec=("image_horizon",),
                                          normalization="BatchNorm", lr=1e-5))
    sound_shapes = {"image_horizon": [3, 64, 64], "sound": [128, 20]}
    gen_train("mopoe_sound", O.OracleConfig(fusion="MoPoE", names_enc=("image_horizon", "sound"), names_rec=("image_horizon", "sound"),
                                            observation_shapes=sound_shapes, lr=1e-5), B=2, T=4)
    gen_train("mopoe_sound_bn", O.OracleConfig(fusion="MoPoE", names_enc=("image_horizon", "sound"), names_rec=("image_horizon", "sound"),
                                               observation_shapes=sound_shapes, normalization="BatchNorm", lr=1e-5), B=2, T=4)
    gen_train("mopoe_logprob", O.OracleConfig(fusion="MoPoE", worldmodel_LogProbLoss=True, predict_reward=True))
    gen_train("single_logprob", O.OracleConfig(fusion="single", names_enc=("image_horizon",), names_rec=("image_horizon",),
                                               worldmodel_LogProbLoss=True))
    gen_train("mopoe_emb512", O.OracleConfig(fusion="MoPoE", embedding_size={"fusion": 1024, "image": 512, "sound": 256, "other": 64}))
    gen_train("mopoe_img128", O.OracleConfig(fusion="MoPoE", names_enc=("image_horizon_128", "pose_quat_v2"),
                                             names_rec=("image_horizon_128", "pose_quat_v2"),
                                             observation_shapes={"image_horizon_128": [3, 128, 128], "pose_quat_v2": [3]}), B=2, T=4)
    gen_train("mopoe_lrramp", O.OracleConfig(fusion="MoPoE", learning_rate_schedule=3), steps=3)
    # the remaining image stacks (encoder.py:362-413, 511-615; observation_model.py:108-160, 231-345) and normalisation variants
    for side, name in ((84, "image_horizon_84"), (256, "image_horizon_256")):
        gen_train(f"mopoe_img{side}", O.OracleConfig(fusion="MoPoE", names_enc=(name, "pose_quat_v2"), names_rec=(name, "pose_quat_v2"),
                                                      observation_shapes={name: [3, side, side], "pose_quat_v2": [3]}), B=2, T=4)
    gen_train("single_img84_bn", O.OracleConfig(fusion="single", names_enc=("image_horizon_84",), names_rec=("image_horizon_84",),
                                                observation_shapes={"image_horizon_84": [3, 84, 84]}, normalization="BatchNorm", lr=1e-5))
    # (an InstanceNorm fixture of the 256x256 stacks was tried and dropped: its last encoder layer normalises planes of 4 values,
    #  which amplifies summation-order noise beyond any useful pin; the product implements None and BatchNorm for the image stacks)
    gen_infer("mopoe", O.OracleConfig(fusion="MoPoE"))
    gen_infer("single", O.OracleConfig(fusion="single", names_enc=("image_horizon",), names_rec=("image_horizon",)))
