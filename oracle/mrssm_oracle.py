"""CPU oracle for the MRSSM training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, from-scratch restatement (plain torch fp32/fp64 on the CPU) of the
arithmetic of EmergentSystemLabStudent/Multimodal-RSSM's ``model.optimize(D)`` path.  It is the
checker for the CUDA product in ``multimodal-rssm_b200/``: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``multimodal-rssm_b200/`` imports it, and the product has no CPU path.

Pinning: the reference ships no tests, golden vectors or fixtures, so the oracle is
pinned against outputs of the *unmodified reference itself*, run in the build container with
injected noise: ``tests/golden/make_golden.py`` (committed) imports ``/root/reference`` and writes
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks every function here against them.

Every function cites the reference file:line it restates (paths relative to the reference root).
All noise is caller supplied (three streams): ``eps_prior``, ``eps_post`` of shape
[T-1,B,S] consumed inside the rollout and ``eps_dec`` [T-1,B,S] consumed by the PoE/MoPoE
decoder-latent resample.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """The hyper-parameters the hot path reads (SURVEY §8b 'Config keys actually read')."""

    belief_size: int = 200
    state_size: int = 30
    hidden_size: int = 200
    action_size: int = 3
    names_enc: Tuple[str, ...] = ("image_horizon", "pose_quat_v2")
    names_rec: Tuple[str, ...] = ("image_horizon", "pose_quat_v2")
    observation_shapes: Dict[str, List[int]] = field(
        default_factory=lambda: {"image_horizon": [3, 64, 64], "pose_quat_v2": [3]})
    embedding_size: Dict[str, int] = field(
        default_factory=lambda: {"fusion": 1024, "image": 1024, "sound": 256, "other": 128})
    act_dense: str = "elu"
    fusion: str = "MoPoE"  # MoPoE | PoE | NN | single   ("single" = algos/MRSSM/RSSM)
    min_std_dev: float = 0.1
    free_nats: float = 3.0
    kl_balancing_alpha: Optional[float] = 0.5
    global_kl_beta: float = 1.0
    kl_beta: float = 1.0
    lr: float = 1e-3
    adam_eps: float = 1e-7
    grad_clip_norm: float = 100.0
    predict_reward: bool = False     # base/algo.py:200-201: False zeroes the reward loss (shipped default)
    learning_rate_schedule: int = 0        # > 0: Adam starts at lr 0 and gains lr/schedule per loss evaluation up to lr (base/algo.py:40-41,195-198)
    worldmodel_LogProbLoss: bool = False   # -log N(o; loc, 1) instead of the squared error (base/algo.py:101-103, :375-378)
    normalization: Optional[str] = None   # None | "BatchNorm" | "InstanceNorm" (image stacks: IMAGE_NORMS below; encoder.py:324-337, observation_model.py:75-86)
    overshooting_distance: int = 0   # latent overshooting (base/algo.py:111-148, MoPoE/algo.py:69-108); kl_beta 0 = off
    overshooting_kl_beta: float = 0.0
    overshooting_reward_scale: float = 0.0

    @property
    def multimodal(self) -> bool:
        return self.fusion != "single"

    @property
    def act_transition(self) -> str:
        # Q1: MultimodalTransitionModel keeps its default 'relu' (transition_model.py:149,156);
        # the single-modal RSSM passes activation_function.dense (RSSM/algo.py:18-19).
        return "relu" if self.multimodal else self.act_dense

    def emb_size_of(self, name: str) -> int:
        # transition_model.py:171-179
        if "image" in name:
            return self.embedding_size["image"]
        if "sound" in name:
            return self.embedding_size["sound"]
        return self.embedding_size["other"]


def _act(name: str):
    return getattr(F, name)


# --------------------------------------------------------------------------------------------
# parameter handling: flat dict  "<group>/<sub>/<torch key>" -> tensor
# --------------------------------------------------------------------------------------------
def flatten_state(nested, prefix="") -> Dict[str, Tensor]:
    """Flatten the reference checkpoint layout (base/algo.py:328-335) into 'a/b/c.weight' keys."""
    out = {}
    for k, v in nested.items():
        key = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            out.update(flatten_state(v, key))
        else:
            out[key] = v
    return out


# (out channels, kernel) of the stride-2 Conv2d / ConvTranspose2d stacks per image side (encoder.py:307-615, observation_model.py:56-345;
# None = the image's channel count) and the normalisation variants the reference defines for each (GroupNorm is not restated).
IMAGE_ENCODER_LAYERS = {64: [(32, 4), (64, 4), (128, 4), (256, 4)], 84: [(32, 4), (64, 5), (128, 5), (256, 6)],
                        128: [(16, 4), (32, 4), (64, 4), (128, 4), (256, 4)], 256: [(8, 4), (16, 4), (32, 4), (64, 4), (128, 4), (256, 4)]}
IMAGE_DECODER_LAYERS = {64: [(128, 5), (64, 5), (32, 6), (None, 6)], 84: [(128, 3), (64, 4), (32, 4), (16, 6), (None, 6)],
                        128: [(256, 6), (128, 4), (64, 4), (32, 4), (None, 6)], 256: [(256, 6), (128, 4), (64, 4), (32, 4), (16, 4), (None, 6)]}
IMAGE_NORMS = {("enc", 64): (None, "BatchNorm"), ("enc", 84): (None, "BatchNorm"), ("enc", 128): (None, "BatchNorm", "InstanceNorm"),
               ("enc", 256): (None, "BatchNorm", "InstanceNorm"), ("dec", 64): (None, "BatchNorm"), ("dec", 84): (None, "BatchNorm"),
               ("dec", 128): (None, "BatchNorm"), ("dec", 256): (None, "BatchNorm", "InstanceNorm")}


def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Shapes of every learnable tensor of the model, keyed like the flattened reference
    checkpoint (SURVEY §5 'Checkpoint / resume'; PyTorch layouts)."""
    D, S, H, A = cfg.belief_size, cfg.state_size, cfg.hidden_size, cfg.action_size
    sh: Dict[str, Tuple[int, ...]] = {}

    def lin(prefix, n_out, n_in):
        sh[prefix + ".weight"] = (n_out, n_in)
        sh[prefix + ".bias"] = (n_out,)

    def head(prefix, n_in):
        lin(prefix + "fc1", H, n_in)
        lin(prefix + "fc2", 2 * S, H)

    def batch_norm(prefix, c):                      # affine parameters; the running statistics are buffers (make_buffers)
        sh[prefix + ".weight"] = (c,)
        sh[prefix + ".bias"] = (c,)

    def image_encoder(prefix, shape):
        c = shape[0]
        layers = IMAGE_ENCODER_LAYERS[shape[1]]
        norm = cfg.normalization
        assert norm in IMAGE_NORMS[("enc", shape[1])], f"the reference has no {norm} variant of the {shape[1]}x{shape[1]} encoder"
        cin = c
        for i, (cout, k) in enumerate(layers):
            if norm is None:                        # Conv2d + ReLU pairs
                sh[f"{prefix}conv.{2 * i}.weight"] = (cout, cin, k, k)
                sh[f"{prefix}conv.{2 * i}.bias"] = (cout,)
            else:                                   # Conv2d(bias=False), BatchNorm2d / InstanceNorm2d, ReLU triples (encoder.py:324-337 ...)
                sh[f"{prefix}conv.{3 * i}.weight"] = (cout, cin, k, k)
                batch_norm(f"{prefix}conv.{3 * i + 1}", cout)
            cin = cout
        if cfg.embedding_size["image"] != 1024:
            lin(prefix + "fc", cfg.embedding_size["image"], 1024)

    def image_decoder(prefix, shape):
        c, E = shape[0], cfg.embedding_size["image"]
        lin(prefix + ("fc" if shape[1] == 84 else "fc1"), E, D + S)          # ImageDecoder_84 names its Linear `fc` (observation_model.py:115)
        layers = IMAGE_DECODER_LAYERS[shape[1]]
        norm = cfg.normalization
        assert norm in IMAGE_NORMS[("dec", shape[1])], f"the reference has no {norm} variant of the {shape[1]}x{shape[1]} decoder"
        cin = E
        for i, (cout, k) in enumerate(layers):
            cout = c if cout is None else cout
            last = i == len(layers) - 1
            if norm is None:
                sh[f"{prefix}conv.{2 * i}.weight"] = (cin, cout, k, k)
                sh[f"{prefix}conv.{2 * i}.bias"] = (cout,)
            else:                                   # ConvT(bias=False), norm, ReLU triples; the last ConvT keeps its bias
                sh[f"{prefix}conv.{3 * i}.weight"] = (cin, cout, k, k)
                if last:
                    sh[f"{prefix}conv.{3 * i}.bias"] = (cout,)
                else:
                    batch_norm(f"{prefix}conv.{3 * i + 1}", cout)
            cin = cout

    def instance_norm(prefix, c):                   # affine parameters (running statistics: make_buffers)
        sh[prefix + ".weight"] = (c,)
        sh[prefix + ".bias"] = (c,)

    def sound_encoder(prefix):                      # SoundEncoder_v2 encoder.py:661-721, channels_base = 128
        cb, emb = 128, cfg.embedding_size["sound"]
        sh[prefix + "down_sample_1.0.weight"] = (cb, 1, 3, 9)
        for i, (ci, co, k) in enumerate([(cb // 2, cb * 2, (4, 8)), (cb, cb * 4, (4, 8)), (cb * 2, cb * 4, (3, 4))]):
            sh[f"{prefix}down_sample_{i + 2}.0.weight"] = (co, ci, *k)
            instance_norm(f"{prefix}down_sample_{i + 2}.1", co)
        sh[prefix + "down_conversion.0.weight"] = (emb // 2, cb * 64, 1)
        instance_norm(prefix + "down_conversion.1", emb // 2)       # InstanceNorm1d(affine=True): no running statistics

    def sound_decoder(prefix):                      # SoundDecoder_v2 observation_model.py:420-472
        cb = 128
        sh[prefix + "up_conversion.weight"] = (cb * 2 * 32 * 4, S + D, 1)
        for i, (ci, co, k) in enumerate([(cb * 2, cb * 4, (3, 4)), (cb * 2, cb * 2, (4, 4)), (cb, cb, (4, 4))]):
            sh[f"{prefix}up_sample_{i}.0.weight"] = (ci, co, *k)    # ConvTranspose2d: [in, out, kh, kw]
            instance_norm(f"{prefix}up_sample_{i}.1", co)
        sh[prefix + "out.weight"] = (1, cb // 2, 7, 7)

    def mlp3(prefix, n_in, n_hid, n_out):
        lin(prefix + "fc1", n_hid, n_in)
        lin(prefix + "fc2", n_hid, n_hid)
        lin(prefix + "fc3", n_out, n_hid)

    if cfg.multimodal:
        tm = "transition_model/main/"
        lin(tm + "fc_embed_state_action", D, S + A)
        sh[tm + "rnn.weight_ih"] = (3 * D, D)
        sh[tm + "rnn.weight_hh"] = (3 * D, D)
        sh[tm + "rnn.bias_ih"] = (3 * D,)
        sh[tm + "rnn.bias_hh"] = (3 * D,)
        head(tm + "stochastic_state_model.", D)
        head("transition_model/obs_encoder/prior_expert/", D)
        for n in cfg.names_enc:
            head(f"transition_model/obs_encoder/{n}/", D + cfg.emb_size_of(n))
        for n in cfg.names_rec:
            if "image" in n:
                image_decoder(f"observation_model/{n}/", cfg.observation_shapes[n])
            elif "sound" in n:
                sound_decoder(f"observation_model/{n}/")
            else:
                mlp3(f"observation_model/{n}/", D + S, cfg.embedding_size["other"],
                     cfg.observation_shapes[n][0])
        mlp3("reward_model/", D + S, H, 1)
        for n in cfg.names_enc:
            if "image" in n:
                image_encoder(f"encoder/{n}/", cfg.observation_shapes[n])
            elif "sound" in n:
                sound_encoder(f"encoder/{n}/")
            else:
                e = cfg.embedding_size["other"]
                mlp3(f"encoder/{n}/", cfg.observation_shapes[n][0], e, e)
    else:
        # single-modal RSSM: one nn.Module state_dict (RSSM/algo.py:48-49)
        tm = "transition_model."
        lin(tm + "fc_embed_state_action", D, S + A)
        sh[tm + "rnn.weight_ih"] = (3 * D, D)
        sh[tm + "rnn.weight_hh"] = (3 * D, D)
        sh[tm + "rnn.bias_ih"] = (3 * D,)
        sh[tm + "rnn.bias_hh"] = (3 * D,)
        head(tm + "stochastic_state_model.", D)
        head(tm + "obs_encoder.", D + cfg.embedding_size["fusion"])
        mlp3("reward_model.", D + S, H, 1)
        n = cfg.names_rec[0]
        if "image" in n:
            image_decoder("observation_model.", cfg.observation_shapes[n])
        else:
            mlp3("observation_model.", D + S, cfg.embedding_size["other"], cfg.observation_shapes[n][0])
        n = cfg.names_enc[0]
        if "image" in n:
            image_encoder("encoder.", cfg.observation_shapes[n])
        else:
            e = cfg.embedding_size["other"]
            mlp3("encoder.", cfg.observation_shapes[n][0], e, e)
    return sh


def make_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic synthetic weights, independent of nn.Module init order: for each tensor in
    sorted key order, U(-b, b) with b = 1/sqrt(fan_in) (the scale of torch's default Linear/Conv
    init).  Used by the golden generator *and* by tests so no weights need to be committed."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    shapes = param_shapes(cfg)
    for k in sorted(shapes):
        shp = shapes[k]
        if len(shp) == 1 and k.endswith(".weight"):             # BatchNorm scale: U(0.5, 1.5)
            out[k] = (torch.rand(shp, generator=g, dtype=torch.float64) + 0.5).to(dtype)
            continue
        if k.endswith("weight") or "weight_" in k:
            if "observation_model" in k and "conv." in k:      # ConvTranspose: [Ci,Co,k,k]
                fan_in = shp[1] * shp[2] * shp[3]
            else:
                fan_in = int(torch.tensor(shp[1:]).prod()) if len(shp) > 1 else shp[0]
        else:
            fan_in = max(shp[0], 16)
        b = 1.0 / math.sqrt(fan_in)
        out[k] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * b).to(dtype)
    if cfg.normalization is not None or any("sound" in n for n in (*cfg.names_enc, *cfg.names_rec)):
        out.update(make_buffers(cfg, dtype))
    return out


BN_EPS, BN_MOMENTUM = 1e-5, 0.1          # nn.BatchNorm2d defaults, as constructed by the reference


def is_buffer(key: str) -> bool:
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def make_buffers(cfg: OracleConfig, dtype=torch.float32) -> Dict[str, Tensor]:
    """Freshly constructed BatchNorm buffers (zeros / ones / 0), keyed like the checkpoint.  They travel in the same flat
    dict as the parameters; train_step leaves them out of the optimiser and lets batch_norm() update them."""
    out = {}
    for k, shp in param_shapes(cfg).items():
        if len(shp) == 1 and k.endswith(".weight") and "down_conversion" not in k:
            base = k[:-len("weight")]
            out[base + "running_mean"] = torch.zeros(shp, dtype=dtype)
            out[base + "running_var"] = torch.ones(shp, dtype=dtype)
            out[base + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return out


def batch_norm(p: "_P", x: Tensor, train: bool) -> Tensor:
    """nn.BatchNorm2d(affine=True, track_running_stats=True): train mode normalises with the batch statistics over
    (N, H, W) (biased variance) and moves the running statistics by momentum 0.1 (unbiased variance); eval mode normalises
    with the running statistics."""
    mean, var = p["running_mean"], p["running_var"]
    if not train:
        xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
        return xh * p["weight"][None, :, None, None] + p["bias"][None, :, None, None]
    m = x.mean(dim=(0, 2, 3))
    v = x.var(dim=(0, 2, 3), unbiased=False)
    n = x.numel() // x.shape[1]
    with torch.no_grad():
        mean.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * m)
        var.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * v * n / (n - 1))
        p["num_batches_tracked"].add_(1)
    xh = (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + BN_EPS)
    return xh * p["weight"][None, :, None, None] + p["bias"][None, :, None, None]


class _P:
    """Prefix view over a flat parameter dict."""

    def __init__(self, params: Dict[str, Tensor], prefix: str):
        self.p, self.prefix = params, prefix

    def __getitem__(self, k):
        return self.p[self.prefix + k]

    def sub(self, s):
        return _P(self.p, self.prefix + s)


def _groups(P: Dict[str, Tensor], cfg: OracleConfig):
    if cfg.multimodal:
        return dict(
            tm=_P(P, "transition_model/main/"),
            experts={"prior_expert": _P(P, "transition_model/obs_encoder/prior_expert/"),
                     **{n: _P(P, f"transition_model/obs_encoder/{n}/") for n in cfg.names_enc}},
            dec={n: _P(P, f"observation_model/{n}/") for n in cfg.names_rec},
            enc={n: _P(P, f"encoder/{n}/") for n in cfg.names_enc},
            reward=_P(P, "reward_model/"))
    return dict(
        tm=_P(P, "transition_model."),
        experts={cfg.names_enc[0]: _P(P, "transition_model.obs_encoder.")},
        dec={cfg.names_rec[0]: _P(P, "observation_model.")},
        enc={cfg.names_enc[0]: _P(P, "encoder.")},
        reward=_P(P, "reward_model."))


# --------------------------------------------------------------------------------------------
# encoders  (utils/models/encoder.py)
# --------------------------------------------------------------------------------------------
def image_encoder(p: _P, x: Tensor, emb: int, act_cnn: str = "relu", train: bool = True, norm: str = "BatchNorm") -> Tensor:
    """ImageEncoder (64x64) encoder.py:307-351 / ImageEncoder_128 :415-500, normalization=None:
    n x (Conv2d k4 s2 + ReLU) then reshape(-1,1024) (flatten order C,H,W).
    normalization="BatchNorm" (encoder.py:324-337): n x (Conv2d k4 s2 without bias + BatchNorm2d + ReLU)."""
    i = 0
    if (p.prefix + "conv.1.running_mean") in p.p:
        while (p.prefix + f"conv.{3 * i}.weight") in p.p:
            nfn = instance_norm if norm == "InstanceNorm" else batch_norm
            x = F.relu(nfn(p.sub(f"conv.{3 * i + 1}."), F.conv2d(x, p[f"conv.{3 * i}.weight"], None, stride=2), train))
            i += 1
        x = x.reshape(-1, 1024)
        if emb != 1024:
            x = _act(act_cnn)(F.linear(x, p["fc.weight"], p["fc.bias"]))
        return x
    while (p.prefix + f"conv.{2 * i}.weight") in p.p:
        x = F.relu(F.conv2d(x, p[f"conv.{2 * i}.weight"], p[f"conv.{2 * i}.bias"], stride=2))
        i += 1
    x = x.reshape(-1, 1024)
    if emb != 1024:                                   # encoder.py:348-349
        x = _act(act_cnn)(F.linear(x, p["fc.weight"], p["fc.bias"]))
    return x


def instance_norm(p: _P, x: Tensor, train: bool) -> Tensor:
    """nn.InstanceNorm2d(affine=True, track_running_stats=True) / nn.InstanceNorm1d(affine=True): train mode (and the 1d
    layer, which tracks nothing, always) normalises every (sample, channel) plane with its own biased variance; the tracked
    layers move their running statistics by momentum 0.1 towards the batch means of the per-plane mean / unbiased variance
    and, in eval mode, normalise with them."""
    dims = tuple(range(2, x.dim()))
    shape = (1, -1) + (1,) * (x.dim() - 2)
    tracked = (p.prefix + "running_mean") in p.p
    if tracked and not train:
        xh = (x - p["running_mean"].reshape(shape)) / torch.sqrt(p["running_var"].reshape(shape) + BN_EPS)
    else:
        m = x.mean(dim=dims, keepdim=True)
        v = x.var(dim=dims, unbiased=False, keepdim=True)
        if tracked:
            n = x[0, 0].numel()
            with torch.no_grad():
                p["running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * m.mean(0).reshape(-1))
                p["running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (v * n / (n - 1)).mean(0).reshape(-1))
                # (torch's instance norm leaves num_batches_tracked at 0)
        xh = (x - m) / torch.sqrt(v + BN_EPS)
    return xh * p["weight"].reshape(shape) + p["bias"].reshape(shape)


def sound_encoder(p: _P, x: Tensor, train: bool = True) -> Tensor:
    """SoundEncoder_v2.forward encoder.py:703-714: [N,128,20] spectrogram -> 4 x (Conv2d without bias [+ InstanceNorm2d] +
    GLU over channels) -> view(N, 8192, 4) -> Conv1d k1 + InstanceNorm1d + GLU -> [N, embedding]."""
    x = F.glu(F.conv2d(x.unsqueeze(1), p["down_sample_1.0.weight"], None, padding=(1, 4)), dim=1)
    for i, (stride, pad) in enumerate([((2, 2), (1, 3)), ((2, 2), (1, 3)), ((1, 1), (1, 1))]):
        q = p.sub(f"down_sample_{i + 2}.")
        x = F.glu(instance_norm(q.sub("1."), F.conv2d(x, q["0.weight"], None, stride=stride, padding=pad), train), dim=1)
    x = x.contiguous().view(-1, p["down_conversion.0.weight"].shape[1], 4)
    x = F.glu(instance_norm(p.sub("down_conversion.1."), F.conv1d(x, p["down_conversion.0.weight"]), train), dim=1)
    return x.contiguous().view(x.shape[0], -1)


def sound_decoder(p: _P, h: Tensor, s: Tensor, train: bool = True) -> Tensor:
    """SoundDecoder_v2.forward observation_model.py:456-472.  Its signature is (s_t, h_t) but every caller passes
    (beliefs, states) positionally, so the concat order is [state, belief] — reproduced here: h = beliefs,
    s = states, input = cat([s, h])."""
    Tn, B = h.shape[:2]
    x = torch.cat([s.reshape(Tn * B, -1, 1), h.reshape(Tn * B, -1, 1)], dim=1)
    x = F.conv1d(x, p["up_conversion.weight"]).view(Tn * B, -1, 32, 4)
    for i, (stride, pad) in enumerate([((1, 1), (1, 1)), ((2, 2), (1, 1)), ((2, 2), (1, 1))]):
        q = p.sub(f"up_sample_{i}.")
        x = F.glu(instance_norm(q.sub("1."), F.conv_transpose2d(x, q["0.weight"], None, stride=stride, padding=pad), train), dim=1)
    x = F.conv2d(x, p["out.weight"], None, padding=3).squeeze(1)
    return x.reshape(Tn, B, *x.shape[1:])


def symbolic_encoder(p: _P, x: Tensor, act: str) -> Tensor:
    """SymbolicEncoder encoder.py:282-296: three Linear+act."""
    a = _act(act)
    x = a(F.linear(x, p["fc1.weight"], p["fc1.bias"]))
    x = a(F.linear(x, p["fc2.weight"], p["fc2.bias"]))
    return a(F.linear(x, p["fc3.weight"], p["fc3.bias"]))


def encode(P, cfg: OracleConfig, obs: Dict[str, Tensor], train: bool = True) -> Dict[str, Tensor]:
    """bottle_tupele_multimodal + MultimodalEncoder.forward (encoder.py:25-48, 778-783):
    fold [T',B,...] -> [T'*B,...], encode every modality, unfold."""
    g = _groups(P, cfg)
    out = {}
    for n in cfg.names_enc:
        x = obs[n]
        Tn, B = x.shape[:2]
        flat = x.reshape(Tn * B, *x.shape[2:])
        if "image" in n:
            e = image_encoder(g["enc"][n], flat, cfg.embedding_size["image"], train=train, norm=cfg.normalization)
        elif "sound" in n:
            e = sound_encoder(g["enc"][n], flat, train)
        else:
            e = symbolic_encoder(g["enc"][n], flat, cfg.act_dense)
        out[n] = e.reshape(Tn, B, -1)
    return out


# --------------------------------------------------------------------------------------------
# Gaussian heads and fusion  (encoder.py:50-155)
# --------------------------------------------------------------------------------------------
def gaussian_head(p: _P, x: Tensor, act: str, min_std: float) -> Tuple[Tensor, Tensor]:
    """StochasticStateModel.forward / ObsEncoder.forward (encoder.py:136-141, 172-176):
    fc2(act(fc1 x)) -> chunk -> (loc, softplus(raw)+min_std)."""
    hid = _act(act)(F.linear(x, p["fc1.weight"], p["fc1.bias"]))
    loc, raw = torch.chunk(F.linear(hid, p["fc2.weight"], p["fc2.bias"]), 2, dim=-1)
    return loc, F.softplus(raw) + min_std


def poe(mu: Tensor, scale: Tensor) -> Tuple[Tensor, Tensor]:
    """encoder.py:50-55.  Q3: 'precision' is 1/sigma, not 1/sigma^2."""
    t = 1.0 / scale
    return (mu * t).sum(0) / t.sum(0), 1.0 / t.sum(0)


def subset_table(n_modalities: int) -> List[Tuple[int, ...]]:
    """Order of the 2^K MoPoE subsets (encoder.py:84-86): itertools.combinations(keys, n) for
    n = 0..K.  Entries index the *modality* experts 1..K; expert 0 (prior_expert) is in all."""
    out = []
    for n in range(n_modalities + 1):
        out += list(itertools.combinations(range(1, n_modalities + 1), n))
    return out


def mopoe_slices(state_size: int, n_subsets: int) -> List[Tuple[int, int]]:
    """Slice bounds of encoder.py:104-120 (Q8): floor(S * (1/n)) dims each, last takes the rest.
    The reference evaluates floor(S*w) with w a float32 tensor (1/n rounded to fp32)."""
    w = torch.tensor(1.0 / float(n_subsets), dtype=torch.float32)
    step = int(torch.floor(state_size * w))
    bounds, start = [], 0
    for k in range(n_subsets):
        end = state_size if k == n_subsets - 1 else start + step
        bounds.append((start, end))
        start = end
    return bounds


def subsets_poe(means: List[Tensor], stds: List[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
    """calc_subset_states (encoder.py:73-97): PoE of {prior_expert} + each modality subset."""
    K = len(means) - 1
    sm, ss = [], []
    for sub in subset_table(K):
        idx = [0] + list(sub)
        m, s = poe(torch.stack([means[i] for i in idx]), torch.stack([stds[i] for i in idx]))
        sm.append(m)
        ss.append(s)
    return sm, ss


def fuse(means: List[Tensor], stds: List[Tensor], fusion: str) -> Tuple[Tensor, Tensor]:
    """get_poe_state / get_mopoe_state without the sample (encoder.py:57-71, 99-124).
    `means[0]` is prior_expert, then modalities in observation_names_enc order."""
    if fusion == "MoPoE":
        sm, ss = subsets_poe(means, stds)
        bounds = mopoe_slices(means[0].shape[-1], len(sm))
        mu = torch.cat([sm[k][..., a:b] for k, (a, b) in enumerate(bounds)], dim=-1)
        sd = torch.cat([ss[k][..., a:b] for k, (a, b) in enumerate(bounds)], dim=-1)
        return mu, sd
    return poe(torch.stack(means), torch.stack(stds))   # PoE and "NN" (Q9)


# --------------------------------------------------------------------------------------------
# the rollout  (utils/models/transition_model.py)
# --------------------------------------------------------------------------------------------
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell semantics: rows (r,z,n); b_hn inside the r* term."""
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def rollout(P, cfg: OracleConfig, prev_state: Tensor, actions: Tensor, prev_belief: Tensor,
            obs_emb: Optional[Dict[str, Tensor]], nonterminals: Optional[Tensor],
            eps_prior: Optional[Tensor], eps_post: Optional[Tensor], det: bool = False) -> dict:
    """MultimodalTransitionModel.forward (transition_model.py:200-285) and TransitionModel.forward
    (:50-114) with explicit noise.  obs_emb=None is the imagination mode.
    Q2: the prior head is evaluated once (the reference's second evaluation is value-identical)."""
    g = _groups(P, cfg)
    tm, act, ms = g["tm"], cfg.act_transition, cfg.min_std_dev
    Tm1 = actions.shape[0]
    h, s = prev_belief, prev_state
    names = list(g["experts"].keys())
    out = {k: [] for k in ["beliefs", "prior_states", "prior_means", "prior_std_devs",
                           "posterior_states", "posterior_means", "posterior_std_devs"]}
    ex_m = {n: [] for n in names}
    ex_s = {n: [] for n in names}
    for t in range(Tm1):
        sm = s if nonterminals is None else s * nonterminals[t]                   # :228-230
        x = _act(act)(F.linear(torch.cat([sm, actions[t]], 1),
                               tm["fc_embed_state_action.weight"], tm["fc_embed_state_action.bias"]))
        h = gru_cell(x, h, tm["rnn.weight_ih"], tm["rnn.weight_hh"], tm["rnn.bias_ih"], tm["rnn.bias_hh"])
        pm, ps = gaussian_head(tm.sub("stochastic_state_model."), h, act, ms)    # :240-241
        pstate = pm if det else pm + ps * eps_prior[t]                            # :242-245
        out["beliefs"].append(h)
        out["prior_means"].append(pm)
        out["prior_std_devs"].append(ps)
        out["prior_states"].append(pstate)
        if obs_emb is None:
            s = pstate
            continue
        means, stds = [], []
        for n in names:                                                           # encoder.py:213-224
            inp = h if n == "prior_expert" else torch.cat([h, obs_emb[n][t]], 1)
            m_, s_ = gaussian_head(g["experts"][n], inp, act, ms)
            means.append(m_)
            stds.append(s_)
            ex_m[n].append(m_)
            ex_s[n].append(s_)
        if cfg.multimodal:
            qm, qs = fuse(means, stds, cfg.fusion)                                # :256-270
        else:
            qm, qs = means[0], stds[0]                                            # :96-107
        qstate = qm if det else qm + qs * eps_post[t]
        out["posterior_means"].append(qm)
        out["posterior_std_devs"].append(qs)
        out["posterior_states"].append(qstate)
        s = qstate
    res = {k: torch.stack(v) for k, v in out.items() if len(v)}
    if obs_emb is not None:
        if cfg.multimodal:
            res["expert_means"] = {n: torch.stack(v) for n, v in ex_m.items()}
            res["expert_std_devs"] = {n: torch.stack(v) for n, v in ex_s.items()}
        else:
            res["expert_means"], res["expert_std_devs"] = None, None              # :113
    return res


# --------------------------------------------------------------------------------------------
# decoders  (utils/models/observation_model.py), reward model
# --------------------------------------------------------------------------------------------
def image_decoder(p: _P, h: Tensor, s: Tensor, train: bool = True, norm: str = "BatchNorm") -> Tensor:
    """ImageDecoder.forward observation_model.py:91-105 (64x64) / ImageDecoder_128 :215-229:
    fc1([h,s]) (no activation) -> [N,E,1,1] -> ConvTranspose2d stack with ReLU between."""
    Tn, B = h.shape[:2]
    fc = "fc1" if (p.prefix + "fc1.weight") in p.p else "fc"            # ImageDecoder_84: `fc`
    x = F.linear(torch.cat([h.reshape(Tn * B, -1), s.reshape(Tn * B, -1)], 1),
                 p[fc + ".weight"], p[fc + ".bias"])
    x = x.reshape(Tn * B, -1, 1, 1)
    if (p.prefix + "conv.1.running_mean") in p.p:           # observation_model.py:75-86
        n = 0
        while (p.prefix + f"conv.{3 * n}.weight") in p.p:
            n += 1
        for i in range(n):
            last = i == n - 1
            x = F.conv_transpose2d(x, p[f"conv.{3 * i}.weight"], p[f"conv.{3 * i}.bias"] if last else None, stride=2)
            if not last:
                nfn = instance_norm if norm == "InstanceNorm" else batch_norm
                x = F.relu(nfn(p.sub(f"conv.{3 * i + 1}."), x, train))
        return x.reshape(Tn, B, *x.shape[1:])
    n = 0
    while (p.prefix + f"conv.{2 * n}.weight") in p.p:
        n += 1
    for i in range(n):
        x = F.conv_transpose2d(x, p[f"conv.{2 * i}.weight"], p[f"conv.{2 * i}.bias"], stride=2)
        if i < n - 1:
            x = F.relu(x)
    return x.reshape(Tn, B, *x.shape[1:])


def mlp3(p: _P, x: Tensor, act: str) -> Tensor:
    a = _act(act)
    x = a(F.linear(x, p["fc1.weight"], p["fc1.bias"]))
    x = a(F.linear(x, p["fc2.weight"], p["fc2.bias"]))
    return F.linear(x, p["fc3.weight"], p["fc3.bias"])


def dense_decoder(p: _P, h: Tensor, s: Tensor, act: str) -> Tensor:
    """DenseDecoder.forward observation_model.py:42-54."""
    Tn, B = h.shape[:2]
    y = mlp3(p, torch.cat([h.reshape(Tn * B, -1), s.reshape(Tn * B, -1)], 1), act)
    return y.reshape(Tn, B, -1)


def decode(P, cfg: OracleConfig, h: Tensor, s: Tensor, train: bool = True) -> Dict[str, Tensor]:
    """MultimodalObservationModel.forward observation_model.py:560-566 ('loc' only)."""
    g = _groups(P, cfg)
    out = {}
    for n in cfg.names_rec:
        if "image" in n:
            out[n] = image_decoder(g["dec"][n], h, s, train, norm=cfg.normalization)
        elif "sound" in n:
            out[n] = sound_decoder(g["dec"][n], h, s, train)
        else:
            out[n] = dense_decoder(g["dec"][n], h, s, cfg.act_dense)
    return out


def reward_model(P, cfg: OracleConfig, h: Tensor, s: Tensor) -> Tensor:
    """RewardModel.forward reward_model.py:20-35."""
    g = _groups(P, cfg)
    Tn, B = h.shape[:2]
    y = mlp3(g["reward"], torch.cat([h.reshape(Tn * B, -1), s.reshape(Tn * B, -1)], 1), cfg.act_dense)
    return y.squeeze(1).reshape(Tn, B)


# --------------------------------------------------------------------------------------------
# ELBO  (algos/MRSSM/base/algo.py, MRSSM_MoPoE/algo.py)
# --------------------------------------------------------------------------------------------
def kl_normal(mq: Tensor, sq: Tensor, mp: Tensor, sp: Tensor) -> Tensor:
    """torch.distributions.kl._kl_normal_normal: 0.5*(var_ratio + t1 - 1 - log var_ratio)."""
    var_ratio = (sq / sp) ** 2
    t1 = ((mq - mp) / sp) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def kl_loss(cfg: OracleConfig, st: dict) -> Tensor:
    """RSSM_base._calc_kl (base/algo.py:75-94) | MRSSM_MoPoE._calc_kl (MoPoE/algo.py:110-137).
    Q6: free-nats clamp per (t,b) after the sum over S, before the mean; MoPoE ignores alpha."""
    pm, ps = st["prior_means"], st["prior_std_devs"]
    free = torch.full((1,), cfg.free_nats, dtype=pm.dtype)
    if cfg.fusion == "MoPoE":
        names = list(st["expert_means"].keys())
        sm, ss = subsets_poe([st["expert_means"][n] for n in names],
                             [st["expert_std_devs"][n] for n in names])
        losses = [torch.max(kl_normal(m, s, pm, ps).sum(2), free).mean((0, 1)) for m, s in zip(sm, ss)]
        return torch.stack(losses).mean(0)
    qm, qs = st["posterior_means"], st["posterior_std_devs"]
    a = cfg.kl_balancing_alpha
    if a is None:
        div = kl_normal(qm, qs, pm, ps).sum(2)
    else:
        div = a * kl_normal(qm.detach(), qs.detach(), pm, ps).sum(2) \
            + (1 - a) * kl_normal(qm, qs, pm.detach(), ps.detach()).sum(2)
    return torch.max(div, free).mean((0, 1))


def decoder_latent(cfg: OracleConfig, st: dict, eps_dec: Optional[Tensor]):
    """_get_posterior_states: base (base/algo.py:157-163) returns the rollout's sample; PoE/MoPoE
    overrides (MRSSM_PoE/algo.py:63-68, MRSSM_MoPoE/algo.py:62-67) re-fuse the stacked experts and
    draw a FRESH sample (Q4)."""
    if cfg.fusion in ("PoE", "MoPoE"):
        names = list(st["expert_means"].keys())
        qm, qs = fuse([st["expert_means"][n] for n in names],
                      [st["expert_std_devs"][n] for n in names], cfg.fusion)
        return qm + qs * eps_dec, qm, qs
    return st["posterior_states"], st["posterior_means"], st["posterior_std_devs"]


def latent_overshooting(P, cfg: OracleConfig, st: dict, actions: Tensor, rewards: Tensor, nonterminals: Tensor,
                        eps_over: List[Tensor]):
    """RSSM_base._latent_overshooting (base/algo.py:111-148) and the MoPoE override (MRSSM_MoPoE/algo.py:69-108).
    actions [T,B,A], rewards [T,B], nonterminals [T,B,1] are the FULL chunk; st holds the T-1 model steps.
    For every start t = 1..T-2 an open-loop (imagination) rollout of up to `overshooting_distance` steps from
    (beliefs[t-1], prior_states[t-1]) is run — all starts concatenated along the batch, shorter runs zero-padded — and
    its priors are pulled towards the DETACHED posteriors of the same steps: max((KL * seq_mask).sum(S), free_nats)
    .mean((0,1)) * overshooting_kl_beta.  MoPoE: one such term (one rollout, fresh noise) per modality subset, averaged.
    eps_over: one [OD, (T-2)*B, S] noise tensor per rollout.  Returns (kl term, reward term)."""
    T, B = actions.shape[0], actions.shape[1]
    OD, S = cfg.overshooting_distance, cfg.state_size
    beliefs, prior_states = st["beliefs"], st["prior_states"]
    if cfg.fusion == "MoPoE":
        names = list(st["expert_means"].keys())
        tm, ts = subsets_poe([st["expert_means"][n] for n in names], [st["expert_std_devs"][n] for n in names])
        targets = list(zip(tm, ts))
    elif cfg.fusion == "PoE":        # _get_posterior_states of MRSSM_PoE re-fuses the experts (same values as the rollout's)
        names = list(st["expert_means"].keys())
        targets = [fuse([st["expert_means"][n] for n in names], [st["expert_std_devs"][n] for n in names], "PoE")]
    else:
        targets = [(st["posterior_means"], st["posterior_std_devs"])]
    free = torch.full((1,), cfg.free_nats, dtype=beliefs.dtype)
    kl_sum = torch.zeros((), dtype=beliefs.dtype)
    reward = torch.zeros((), dtype=beliefs.dtype)
    last = None
    for (qm_all, qs_all), eps in zip(targets, eps_over):
        acts, nts, rws, h0, s0, qm, qs, msk = [], [], [], [], [], [], [], []
        for t in range(1, T - 1):
            d = min(t + OD, T - 1)
            t_, d_ = t - 1, d - 1
            pad = t - d + OD                                                  # base:124
            padt = lambda x, v=0.0: F.pad(x, (0, 0) * (x.dim() - 1) + (0, pad), value=v)
            acts.append(padt(actions[t:d]))
            nts.append(padt(nonterminals[t:d]))
            rws.append(padt(rewards[t:d]))
            h0.append(beliefs[t_])
            s0.append(prior_states[t_])
            qm.append(padt(qm_all[t_ + 1:d_ + 1].detach()))
            qs.append(padt(qs_all[t_ + 1:d_ + 1].detach(), 1.0))              # std padded with 1 (no infinite KL)
            msk.append(padt(torch.ones(d - t, B, S, dtype=beliefs.dtype)))
        out = rollout(P, cfg, torch.cat(s0, 0), torch.cat(acts, 1), torch.cat(h0, 0), None, torch.cat(nts, 1), eps, None)
        seq_mask = torch.cat(msk, 1)
        div = kl_normal(torch.cat(qm, 1), torch.cat(qs, 1), out["prior_means"], out["prior_std_devs"])
        kl_sum = kl_sum + cfg.overshooting_kl_beta * torch.max((div * seq_mask).sum(2), free).mean((0, 1))
        last = (out, seq_mask, torch.cat(rws, 1))
    kl_sum = kl_sum / len(targets)                                             # MoPoE:101 (a single term otherwise)
    if cfg.overshooting_reward_scale != 0:                                     # base:143-146 (MoPoE: the LAST subset's rollout)
        out, seq_mask, rw = last
        r = reward_model(P, cfg, out["beliefs"], out["prior_states"])
        reward = (1.0 / OD) * cfg.overshooting_reward_scale * \
            F.mse_loss(r * seq_mask[:, :, 0], rw, reduction="none").mean((0, 1)) * (T - 1)
    return kl_sum, reward


def elbo(P, cfg: OracleConfig, st: dict, obs_target: Dict[str, Tensor], eps_dec: Optional[Tensor],
         rewards: Optional[Tensor] = None, actions: Optional[Tensor] = None, nonterminals: Optional[Tensor] = None,
         eps_over: Optional[List[Tensor]] = None, train: bool = True):
    """_calc_loss + _get_model_loss (base/algo.py:165-232): overshooting off, MSE observation loss
    mean over (t,b) then sum over features (base/algo.py:381-383); the reward loss (_calc_reward_loss
    base/algo.py:96-109: MSE of the reward head on [h, z] against rewards[:-1], mean over (t,b)) is
    zeroed unless predict_reward (base/algo.py:200-201)."""
    z, qm, qs = decoder_latent(cfg, st, eps_dec)
    rec = decode(P, cfg, st["beliefs"], z, train)
    def point_loss(pred, target):
        """Per-element loss: squared error, or -log N(target; pred, 1) = (target - pred)^2 / 2 + log sqrt(2 pi) with
        worldmodel_LogProbLoss (torch.distributions.Normal.log_prob, scale 1.0)."""
        if cfg.worldmodel_LogProbLoss:
            return 0.5 * (target - pred) ** 2 + 0.5 * math.log(2 * math.pi)
        return F.mse_loss(pred, target, reduction="none")

    obs_loss = {n: point_loss(rec[n], obs_target[n]).mean((0, 1)).sum() for n in cfg.names_rec}
    kl = kl_loss(cfg, st)
    kl_sum = kl.clone()
    if cfg.global_kl_beta != 0:                                                    # :186-188
        kl_sum = kl_sum + cfg.global_kl_beta * kl_normal(
            qm, qs, torch.zeros_like(qm), torch.ones_like(qs)).sum(2).mean((0, 1))
    obs_sum = sum(obs_loss.values())
    reward_loss = torch.zeros(())
    need_reward = cfg.predict_reward
    if need_reward:                                                                 # :96-109, :175
        r = reward_model(P, cfg, st["beliefs"], z)
        reward_loss = point_loss(r, rewards[:-1]).mean((0, 1))
    if cfg.overshooting_kl_beta != 0:                                               # :190-193
        kl_o, r_o = latent_overshooting(P, cfg, st, actions, rewards, nonterminals, eps_over)
        kl_sum = kl_sum + kl_o
        reward_loss = reward_loss + r_o
    if not cfg.predict_reward:                                                      # :200-201
        reward_loss = torch.zeros(())
    model_loss = obs_sum + reward_loss + cfg.kl_beta * kl_sum                      # :221
    info = {"observations_loss_sum": obs_sum, "reward_loss": reward_loss,
            "kl_loss_sum": kl_sum, "kl_loss": kl}
    for n in cfg.names_rec:
        info[f"observation_{n}_loss"] = obs_loss[n]
    return model_loss, info


def estimate_state(P, cfg: OracleConfig, obs_target, actions, nonterminals, eps_prior, eps_post,
                   det=False, train: bool = True) -> dict:
    """MRSSM_base.estimate_state base/algo.py:337-366 (zeros init, encode, rollout).  train: BatchNorm mode of the encoders
    (model.train() / model.eval())."""
    B = actions.shape[1]
    dt = actions.dtype
    emb = encode(P, cfg, obs_target, train)
    if not cfg.multimodal:
        emb = {cfg.names_enc[0]: emb[cfg.names_enc[0]]}
    return rollout(P, cfg, torch.zeros(B, cfg.state_size, dtype=dt), actions,
                   torch.zeros(B, cfg.belief_size, dtype=dt), emb, nonterminals, eps_prior, eps_post, det)


# --------------------------------------------------------------------------------------------
# optimiser  (base/algo.py:40-42, 255-260)
# --------------------------------------------------------------------------------------------
def clip_coef(grads: List[Tensor], max_norm: float) -> Tuple[Tensor, Tensor]:
    """torch.nn.utils.clip_grad_norm_: total L2 norm; coef = clamp(max_norm/(norm+1e-6), max=1)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adam_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, eps: float,
                b1: float = 0.9, b2: float = 0.999):
    """torch.optim.Adam (no amsgrad, no weight decay), in place."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def train_step(P: Dict[str, Tensor], opt: dict, cfg: OracleConfig, batch: dict, noise: dict) -> dict:
    """RSSM_base.optimize (base/algo.py:268-276) on one already-sampled batch.
    batch: obs {name:[T,B,...]}, actions [T,B,A], nonterminals [T,B,1]; noise: eps_prior/eps_post/
    eps_dec [T-1,B,S].  Mutates P and opt (keys 'm','v','step').  Returns loss_info, grads,
    grad_norm."""
    leaves = {k: (v if is_buffer(k) else v.detach().clone().requires_grad_(True)) for k, v in P.items()}   # BatchNorm buffers: shared, updated in place
    tgt = {n: o[1:] for n, o in batch["obs"].items()}                               # base:241
    st = estimate_state(leaves, cfg, {n: tgt[n] for n in cfg.names_enc}, batch["actions"][:-1],
                        batch["nonterminals"][:-1], noise["eps_prior"], noise["eps_post"])
    loss, info = elbo(leaves, cfg, st, tgt, noise.get("eps_dec"), batch.get("rewards"), batch["actions"],
                      batch["nonterminals"], noise.get("eps_over"))
    loss.backward()
    keys = [k for k in leaves if leaves[k].grad is not None]     # reward model: grad None (a24)
    grads = {k: leaves[k].grad for k in keys}
    total, coef = clip_coef(list(grads.values()), cfg.grad_clip_norm)
    opt["step"] = opt.get("step", 0) + 1
    lr = cfg.lr
    if cfg.learning_rate_schedule != 0:        # the ramp advances inside _calc_loss, i.e. before this step's update
        lr = opt["lr"] = min(opt.get("lr", 0.0) + cfg.lr / cfg.learning_rate_schedule, cfg.lr)
    for k in keys:
        if k not in opt.setdefault("m", {}):
            opt["m"][k] = torch.zeros_like(P[k])
            opt.setdefault("v", {})[k] = torch.zeros_like(P[k])
        adam_update(P[k], grads[k] * coef, opt["m"][k], opt["v"][k], opt["step"], lr, cfg.adam_eps)
    return {"loss_info": {k: float(v) for k, v in info.items()}, "model_loss": float(loss),
            "grads": grads, "grad_norm": float(total), "states": st}


# --------------------------------------------------------------------------------------------
# synthetic COBOTTA-shaped data
# --------------------------------------------------------------------------------------------
def synthetic_batch(cfg: OracleConfig, B: int, T: int, seed: int = 1234, dtype=torch.float32):
    """Seeded synthetic batch, time-major.  Images follow normalize_image(bit_depth=5)'s output
    distribution (utils/processing/image_processing.py:5-11): floor(u8/8)/32 - 0.5 + U(0,1/32)."""
    g = torch.Generator().manual_seed(seed)
    obs = {}
    for n in sorted(set(cfg.names_enc) | set(cfg.names_rec)):
        shp = cfg.observation_shapes[n]
        if "image" in n:
            u8 = torch.randint(0, 256, (T, B, *shp), generator=g)
            obs[n] = (torch.floor(u8 / 8) / 32 - 0.5 + torch.rand((T, B, *shp), generator=g) / 32).to(dtype)
        else:
            obs[n] = torch.randn((T, B, *shp), generator=g).to(dtype)
    actions = torch.randn((T, B, cfg.action_size), generator=g).to(dtype)
    nonterm = torch.ones((T, B, 1), dtype=dtype)
    drop = torch.rand(B, generator=g) < 0.1
    if B > 1:
        drop[1] = True            # small batches always exercise the mask
    tpos = torch.randint(0, T, (B,), generator=g)
    for b in range(B):
        if drop[b]:
            nonterm[tpos[b], b, 0] = 0
    S = cfg.state_size
    noise = {k: torch.randn((T - 1, B, S), generator=g).to(dtype)
             for k in ("eps_prior", "eps_post", "eps_dec")}
    if cfg.overshooting_kl_beta != 0:     # own generator again: one [OD, (T-2)B, S] tensor per imagination rollout
        go = torch.Generator().manual_seed(seed + 177)
        n_roll = 2 ** (len(cfg.names_enc)) if cfg.fusion == "MoPoE" else 1
        noise["eps_over"] = [torch.randn((cfg.overshooting_distance, (T - 2) * B, S), generator=go).to(dtype) for _ in range(n_roll)]
        if cfg.fusion == "PoE":               # the base _latent_overshooting calls _get_posterior_states once more: one unused draw
            noise["eps_dec2"] = torch.randn((T - 1, B, S), generator=go).to(dtype)
    rewards = torch.zeros(T, B, dtype=dtype)
    if cfg.predict_reward:           # own generator: the other tensors of a seed do not change with this switch
        rewards = torch.randn((T, B), generator=torch.Generator().manual_seed(seed + 77)).to(dtype)
    batch = {"obs": obs, "actions": actions, "rewards": rewards, "nonterminals": nonterm}
    return batch, noise
