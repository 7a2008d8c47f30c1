"""Encoders, Gaussian heads and expert fusion of the MRSSM — B200 host side.

Drop-in for the reference's ``utils/models/encoder.py`` (same public names, constructor arguments,
return structures, ``.modules`` list attributes and state-dict keys); the arithmetic runs in
libmrssm_b200.so through ``mrssm_b200.ops``.  torch.nn layers are used only as parameter
containers (identical names/shapes/initialisation to the reference, so its checkpoints load); their
own ``forward`` is never called.

Covered: bottle helpers a3, MultimodalEncoder a4, ImageEncoder 64/128 a5,
SymbolicEncoder a6, StochasticStateModel a11, ObsEncoder a12, MultimodalObsEncoder a13, fusion a14.
Round 2 added the rest of the shipped YAML (`SoundEncoder_v2`, BatchNorm image stacks) and the 84x84 / 256x256 stacks (general NCHW
kernels).  Not covered (raise NotImplementedError): the q(st|ot) expert family, InstanceNorm / GroupNorm image variants.
"""
import itertools

import torch
from torch import nn
from torch.nn import functional as F

from mrssm_b200 import noise, ops

_ACT = {"relu": ops.RELU, "elu": ops.ELU}


def act_code(activation):
    """Accept the reference's spellings: a name or a torch.nn.functional callable."""
    name = activation if isinstance(activation, str) else getattr(activation, "__name__", None)
    if name not in _ACT:
        raise NotImplementedError(f"activation {activation!r}: only relu / elu are implemented")
    return _ACT[name]


# ---- time folding (reference encoder.py:13-48) -------------------------------------------------------
def _fold(x):
    return x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])


def bottle_tupele(f, x_tuple, var_name="", kwargs={}):
    x = next(iter(x_tuple.values()))
    lead = x.shape[:2]
    y = f(_fold(x), **kwargs)
    if var_name:
        y = y[var_name]
    return y.reshape(*lead, *y.shape[1:])


def bottle_tupele_multimodal(f, x_tuples, var_name="", kwargs={}):
    lead = next(iter(x_tuples.values())).shape[:2]
    y = f({name: _fold(x) for name, x in x_tuples.items()}, **kwargs)
    if var_name:
        y = y[var_name]

    def unfold(v):
        if torch.is_tensor(v):
            return v.reshape(*lead, *v.shape[1:])
        return {k: unfold(t) for k, t in v.items()}

    return {name: unfold(v) for name, v in y.items()}


# ---- expert fusion (reference encoder.py:50-124) -------------------------------------------------------
def _fused(expert_means, expert_std_devs, fusion):
    names = list(expert_means.keys())
    if names[0] != "prior_expert":
        raise ValueError("expert dicts must start with 'prior_expert' (MultimodalObsEncoder order)")
    m0 = expert_means[names[0]]
    table = ops.FusionTable(len(names), m0.shape[-1], fusion)
    spec = ops.LatentSpec(m0.shape[-1], table, kl_mode=0, refuse=True, free_nats=0.0, alpha=None)
    eps = noise.draw("dec", m0.shape, m0.device)
    ones = torch.ones_like(m0)
    z, qm, qs, _ = ops.LatentFn.apply(spec, torch.zeros_like(m0), ones, m0, ones, eps,
                                      *[expert_means[n] for n in names], *[expert_std_devs[n] for n in names])
    return z, qm, qs


def poe(mu, scale):
    """Product of experts with 1/sigma weights (reference encoder.py:50-55, Q3).  mu/scale: [K,...]."""
    names = ["prior_expert"] + [f"e{i}" for i in range(1, mu.shape[0])]
    _, qm, qs = _fused(dict(zip(names, mu)), dict(zip(names, scale)), "PoE")
    return qm, qs


def get_poe_state(expert_means, expert_std_devs):
    """(sample, mean, std) of the product of all experts; the sample uses the 'dec' noise stream."""
    return _fused(expert_means, expert_std_devs, "PoE")


def get_mopoe_state(expert_means, expert_std_devs):
    """(sample, mean, std) of the MoPoE posterior: state dims are sliced across the 2^K subsets (Q8)."""
    return _fused(expert_means, expert_std_devs, "MoPoE")


def calc_subset_states(expert_means, expert_std_devs):
    """Means/stds of every subset PoE, in itertools.combinations order (reference encoder.py:73-97)."""
    keys = [k for k in expert_means.keys() if k != "prior_expert"]
    sub_m, sub_s = [], []
    for n in range(len(keys) + 1):
        for combo in itertools.combinations(keys, n):
            names = ["prior_expert", *combo]
            _, qm, qs = _fused({k: expert_means[k] for k in names}, {k: expert_std_devs[k] for k in names}, "PoE")
            sub_m.append(qm)
            sub_s.append(qs)
    return sub_m, sub_s


# ---- Gaussian heads ----------------------------------------------------------------------------------------
class _GaussianHead(nn.Module):
    """fc1 -> act -> fc2 -> (loc, softplus + min_std).  In the training path the head is evaluated
    inside the rollout kernel; this standalone forward exists for API parity."""

    def __init__(self, in_size, hidden_size, s_size, activation, min_std_dev):
        super().__init__()
        self.fc1 = nn.Linear(in_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, 2 * s_size)
        self.activation = activation
        self.min_std_dev = min_std_dev

    def _loc_scale(self, *parts):
        out = ops.MlpFn.apply(act_code(self.activation), False, len(parts), *parts,
                              self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        loc, raw = torch.chunk(out, 2, dim=1)
        return {"loc": loc, "scale": F.softplus(raw) + self.min_std_dev}

    def get_state_dict(self):
        return self.state_dict()

    def _load_state_dict(self, state_dict):
        self.load_state_dict(state_dict)

    def get_model_params(self):
        return list(self.parameters())


class StochasticStateModel(_GaussianHead):
    """p(s_t | h_t)  (reference encoder.py:126-155)."""

    def __init__(self, h_size, hidden_size, activation, s_size, min_std_dev):
        super().__init__(h_size, hidden_size, s_size, activation, min_std_dev)

    def forward(self, h_t):
        return self._loc_scale(h_t)

    def sample(self, h_t):
        d = self.forward(h_t)
        return d["loc"] + d["scale"] * noise.draw("prior", d["loc"].shape, h_t.device)


class ObsEncoder(_GaussianHead):
    """q(s_t | h_t, o_t)  (reference encoder.py:157-190)."""

    def __init__(self, h_size, s_size, activation, embedding_size, hidden_size, min_std_dev):
        super().__init__(h_size + embedding_size, hidden_size, s_size, activation, min_std_dev)
        self.modules = [self.fc1, self.fc2]

    def forward(self, h_t, o_t):
        return self._loc_scale(h_t, o_t)

    def get_loc_and_scale(self, h_t, obs_emb):
        return self.forward(h_t, obs_emb)

    def sample(self, h_t, o_t):
        d = self.forward(h_t, o_t)
        return d["loc"] + d["scale"] * noise.draw("post", d["loc"].shape, h_t.device)


class MultimodalObsEncoder:
    """Dict of expert heads: 'prior_expert' (no observation) then one ObsEncoder per modality
    (reference encoder.py:196-248)."""

    def __init__(self, expert_dist, h_size, s_size, activation, embedding_sizes, hidden_size, min_std_dev, device):
        if expert_dist != "q(st|ht,ot)":
            raise NotImplementedError(f"expert_dist {expert_dist!r} is outside the B200 hot path (SURVEY Q17)")
        self.expert_dist = expert_dist
        self.obs_encoder = {"prior_expert": StochasticStateModel(
            h_size=h_size, hidden_size=hidden_size, activation=activation, s_size=s_size, min_std_dev=min_std_dev).to(device)}
        self.modules = [self.obs_encoder["prior_expert"]]
        for name, emb in embedding_sizes.items():
            head = ObsEncoder(h_size=h_size, s_size=s_size, activation=activation, embedding_size=emb,
                              hidden_size=hidden_size, min_std_dev=min_std_dev).to(device)
            self.obs_encoder[name] = head
            self.modules += head.modules

    def get_loc_and_scale(self, h_t, obs_emb, t):
        out = {}
        for name, head in self.obs_encoder.items():
            out[name] = head(h_t=h_t) if name == "prior_expert" else head(h_t=h_t, o_t=obs_emb[name][t])
        return out

    def get_state_dict(self):
        return {name: head.state_dict() for name, head in self.obs_encoder.items()}

    def _load_state_dict(self, state_dict):
        for name, sd in state_dict.items():
            self.obs_encoder[name].load_state_dict(sd)

    def get_model_params(self):
        return [p for head in self.obs_encoder.values() for p in head.parameters()]

    def eval(self):
        for head in self.obs_encoder.values():
            head.eval()

    def train(self):
        for head in self.obs_encoder.values():
            head.train()


# ---- observation encoders -------------------------------------------------------------------------------------
class _EncoderBase(nn.Module):
    def get_state_dict(self):
        return self.state_dict()

    def _load_state_dict(self, state_dict):
        self.load_state_dict(state_dict)

    def get_model_params(self):
        return list(self.parameters())


class SymbolicEncoder(_EncoderBase):
    """Three Linear+act layers on a vector observation (reference encoder.py:282-305)."""

    def __init__(self, observation_size, embedding_size, activation_function="relu"):
        super().__init__()
        self.embedding_size = embedding_size
        self.activation_function = activation_function
        self.fc1 = nn.Linear(observation_size, embedding_size)
        self.fc2 = nn.Linear(embedding_size, embedding_size)
        self.fc3 = nn.Linear(embedding_size, embedding_size)
        self.modules = [self.fc1, self.fc2, self.fc3]

    def forward(self, observation):
        return ops.mlp(act_code(self.activation_function), True, 1, observation,
                               self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                               self.fc3.weight, self.fc3.bias)


class _ConvEncoder(_EncoderBase):
    """Stride-2 Conv2d stacks of the reference's image encoders (encoder.py:307-615): LAYERS = (out channels, kernel) per layer.
    normalization None: Conv2d + ReLU pairs (keys conv.{0,2,..}.{weight,bias}); "BatchNorm": Conv2d(bias=False) + BatchNorm2d + ReLU
    triples (keys conv.{0,3,..}.weight, conv.{1,4,..}.*).  The 64x64 / 128x128 stacks without normalisation run on the plane tcgen05
    kernels (bf16 mode) or the exact SIMT kernels; every other variant on the general NCHW kernels (csrc/generic_nchw.cu: exact
    fp32, tensor-core route for the wide layers in bf16 mode)."""
    LAYERS = ()
    PLANE_PATH = False          # the validated fast path of ops.ConvEncoder(TC)Fn (k = 4 everywhere)

    def __init__(self, embedding_size, activation_function="relu", image_dim=3, normalization=None):
        super().__init__()
        if normalization not in (None, "BatchNorm"):
            raise NotImplementedError(f"normalization={normalization!r}: None and BatchNorm are implemented (InstanceNorm / GroupNorm "
                                      "variants of the 128x128 / 256x256 stacks are not)")
        self.embedding_size = embedding_size
        self.activation_function = activation_function
        self.normalization = normalization
        layers, cin = [], image_dim
        for cout, k in self.LAYERS:
            if normalization == "BatchNorm":
                layers += [nn.Conv2d(cin, cout, k, stride=2, bias=False), nn.BatchNorm2d(cout, affine=True, track_running_stats=True), nn.ReLU()]
            else:
                layers += [nn.Conv2d(cin, cout, k, stride=2), nn.ReLU()]
            cin = cout
        self.conv = nn.Sequential(*layers)               # parameter / buffer container
        self.fc = nn.Identity() if embedding_size == 1024 else nn.Linear(1024, embedding_size)
        self.modules = [self.conv, self.fc]

    def forward(self, observation):
        mods = list(self.conv)
        if self.normalization == "BatchNorm":
            hidden = observation
            for i in range(0, len(mods), 3):
                hidden = ops.batch_norm(ops.conv2d_nobias(hidden, mods[i].weight, 2, 0), mods[i + 1], relu=True)
            hidden = hidden.reshape(-1, 1024)
        elif not self.PLANE_PATH:
            hidden = observation
            for i in range(0, len(mods), 2):
                hidden = ops.add_channel_bias(ops.conv2d_nobias(hidden, mods[i].weight, 2, 0), mods[i].bias, relu=True)
            hidden = hidden.reshape(-1, 1024)
        else:
            params = [p for m in self.conv if isinstance(m, nn.Conv2d) for p in (m.weight, m.bias)]
            fn = ops.ConvEncoderTCFn if ops.bf16_mode() else ops.ConvEncoderFn
            hidden = fn.apply(observation, *params)
        if hidden.shape[1] != 1024:
            raise ValueError(f"conv stack produced {hidden.shape[1]} features, the reference reshapes to 1024")
        if self.embedding_size != 1024:
            hidden = ops.mlp(act_code(self.activation_function), True, 1, hidden, self.fc.weight, self.fc.bias)
        return hidden


class ImageEncoder(_ConvEncoder):
    """64x64: Conv 3->32->64->128->256, k4 s2, ReLU (reference encoder.py:307-360)."""
    LAYERS = ((32, 4), (64, 4), (128, 4), (256, 4))
    PLANE_PATH = True


class ImageEncoder_84(_ConvEncoder):
    """84x84: Conv 3->32 k4 ->64 k5 ->128 k5 ->256 k6, s2, ReLU (reference encoder.py:362-413)."""
    LAYERS = ((32, 4), (64, 5), (128, 5), (256, 6))


class ImageEncoder_128(_ConvEncoder):
    """128x128: Conv 3->16->32->64->128->256, k4 s2, ReLU (reference encoder.py:415-509)."""
    LAYERS = ((16, 4), (32, 4), (64, 4), (128, 4), (256, 4))
    PLANE_PATH = True


class ImageEncoder_256(_ConvEncoder):
    """256x256: Conv 3->8->16->32->64->128->256, k4 s2, ReLU (reference encoder.py:511-615)."""
    LAYERS = ((8, 4), (16, 4), (32, 4), (64, 4), (128, 4), (256, 4))


class SoundEncoder_v2(_EncoderBase):
    """Spectrogram [N,128,20] -> embedding (reference encoder.py:661-721): four bias-free Conv2d (+ InstanceNorm2d with tracked
    statistics) + GLU stages, a k = 1 Conv1d + InstanceNorm1d + GLU down-conversion.  Same constructor, `.modules`, state-dict
    keys (`down_sample_k.{0,1}.*`, `down_conversion.{0,1}.*`).  Exact fp32 kernels (csrc/generic_nchw.cu)."""

    def __init__(self, embbed_size=250, channels_base=128):
        super().__init__()
        cb = channels_base
        self.embbed_size = embbed_size
        self.conversion_channels = int(cb * 64)
        inorm = lambda c: nn.InstanceNorm2d(num_features=c, affine=True, track_running_stats=True)
        self.down_sample_1 = nn.Sequential(nn.Conv2d(1, cb, kernel_size=(3, 9), padding=(1, 4), bias=False), nn.GLU(dim=1))
        self.down_sample_2 = nn.Sequential(nn.Conv2d(cb // 2, cb * 2, kernel_size=(4, 8), stride=(2, 2), padding=(1, 3), bias=False),
                                           inorm(cb * 2), nn.GLU(dim=1))
        self.down_sample_3 = nn.Sequential(nn.Conv2d(cb, cb * 4, kernel_size=(4, 8), stride=(2, 2), padding=(1, 3), bias=False),
                                           inorm(cb * 4), nn.GLU(dim=1))
        self.down_sample_4 = nn.Sequential(nn.Conv2d(cb * 2, cb * 4, kernel_size=(3, 4), stride=(1, 1), padding=(1, 1), bias=False),
                                           inorm(cb * 4), nn.GLU(dim=1))
        self.down_conversion = nn.Sequential(nn.Conv1d(self.conversion_channels, int(embbed_size / 2), kernel_size=1, bias=False),
                                             nn.InstanceNorm1d(num_features=int(embbed_size / 2), affine=True), nn.GLU(dim=1))
        self.modules = [self.down_sample_1, self.down_sample_2, self.down_sample_3, self.down_sample_4, self.down_conversion]

    def forward(self, x):
        x = x.unsqueeze(1)
        c = self.down_sample_1[0]
        x = ops.GluFn.apply(ops.conv2d_nobias(x, c.weight, c.stride, c.padding))
        for stage in (self.down_sample_2, self.down_sample_3, self.down_sample_4):
            c = stage[0]
            x = ops.GluFn.apply(ops.instance_norm(ops.conv2d_nobias(x, c.weight, c.stride, c.padding), stage[1]))
        x = x.contiguous().view(-1, self.conversion_channels, 4)
        x = ops.GluFn.apply(ops.instance_norm(ops.conv1d_k1(x, self.down_conversion[0].weight), self.down_conversion[1]))
        return x.contiguous().view(-1, self.embbed_size)


def build_ImageEncoder(observation_shape, visual_embedding_size, cnn_activation_function, normalization=None):
    size = list(observation_shape[1:])
    cls = {(64, 64): ImageEncoder, (84, 84): ImageEncoder_84, (128, 128): ImageEncoder_128, (256, 256): ImageEncoder_256}.get(tuple(size))
    if cls is None:
        raise NotImplementedError(f"image size {size}: the reference defines 64, 84, 128 and 256")
    return cls(visual_embedding_size, cnn_activation_function, image_dim=observation_shape[0], normalization=normalization)


def build_Encoder(name, observation_shapes, embedding_size, activation_function, normalization=None):
    shape = observation_shapes[name]
    if "image" in name:
        return build_ImageEncoder(shape, embedding_size["image"], activation_function["cnn"], normalization=normalization)
    if "sound" in name:
        return SoundEncoder_v2(embbed_size=embedding_size["sound"])
    return SymbolicEncoder(shape[0], embedding_size["other"], activation_function["dense"])


class MultimodalEncoder:
    """One encoder per name in observation_names_enc (reference encoder.py:746-810)."""

    def __init__(self, observation_names_enc, observation_shapes, embedding_size, activation_function,
                 normalization=None, device=torch.device("cpu")):
        self.observation_names_enc = observation_names_enc
        self.encoders = {}
        self.modules = []
        for name in observation_names_enc:
            enc = build_Encoder(name, observation_shapes, embedding_size, activation_function, normalization).to(device)
            self.encoders[name] = enc
            self.modules += enc.modules

    def get_obs(self, observations, name):
        if name in observations:
            return observations[name]
        alias = {"observation": "image", "image": "observation"}.get(name)
        if alias in observations:
            return observations[alias]
        print("{} is missing in {}".format(name, observations.keys()))
        raise NotImplementedError

    def forward(self, observations):
        return {name: enc(self.get_obs(observations, name)) for name, enc in self.encoders.items()}

    __call__ = forward

    def get_state_dict(self):
        return {name: enc.state_dict() for name, enc in self.encoders.items()}

    def _load_state_dict(self, state_dict):
        for name, sd in state_dict.items():
            self.encoders[name].load_state_dict(sd)

    def get_model_params(self):
        return [p for enc in self.encoders.values() for p in enc.parameters()]

    def eval(self):
        for enc in self.encoders.values():
            enc.eval()

    def train(self):
        for enc in self.encoders.values():
            enc.train()
