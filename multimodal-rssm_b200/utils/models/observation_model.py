"""Observation decoders — B200 host side.

Drop-in for the reference's ``utils/models/observation_model.py`` for the modalities on the hot path:
``DenseDecoder`` (vectors), ``ImageDecoder`` (64x64), ``ImageDecoder_128`` and the
``MultimodalObservationModel`` container, with the same constructor arguments, ``forward`` /
``get_mse`` / ``get_log_prob`` API, ``.modules`` lists and state-dict keys.  In addition every
decoder offers ``mse_loss(h_t, s_t, o_t)`` = sum_features mean_{t,b} (loc - o)^2 as a scalar, which is
what the training step uses (the reference materialises the element-wise tensor first).
"""
import torch
from torch import nn

from mrssm_b200 import ops
from utils.models.encoder import act_code

_LOG_SQRT_2PI = 0.9189385332046727


class ObservationModel_base(nn.Module):
    def get_state_dict(self):
        return self.state_dict()

    def _load_state_dict(self, state_dict):
        self.load_state_dict(state_dict)

    def get_model_params(self):
        return list(self.parameters())

    def get_mse(self, h_t, s_t, o_t):
        """Element-wise squared error (reference observation_model.py:28-31), differentiable."""
        loc = self.forward(h_t, s_t)["loc"]
        return (loc - o_t) ** 2

    def get_log_prob(self, h_t, s_t, o_t):
        loc = self.forward(h_t, s_t)["loc"]
        return -0.5 * (o_t - loc) ** 2 - _LOG_SQRT_2PI           # Normal(loc, 1).log_prob(o)

    def mse_loss(self, h_t, s_t, o_t):
        loc = self.forward(h_t, s_t)["loc"]
        T, B = h_t.shape[:2]
        return ops.MseLossFn.apply(loc, o_t, T * B)


class DenseDecoder(ObservationModel_base):
    """Linear(D+S,E)+act -> Linear(E,E)+act -> Linear(E,obs) (reference observation_model.py:33-54)."""

    def __init__(self, observation_size, belief_size, state_size, embedding_size, activation_function="relu"):
        super().__init__()
        self.activation_function = activation_function
        self.fc1 = nn.Linear(belief_size + state_size, embedding_size)
        self.fc2 = nn.Linear(embedding_size, embedding_size)
        self.fc3 = nn.Linear(embedding_size, observation_size)
        self.modules = [self.fc1, self.fc2, self.fc3]

    def forward(self, h_t, s_t):
        T, B = h_t.shape[:2]
        y = ops.mlp(act_code(self.activation_function), False, 2, h_t.reshape(T * B, -1), s_t.reshape(T * B, -1),
                            self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                            self.fc3.weight, self.fc3.bias)
        return {"loc": y.reshape(T, B, -1), "scale": 1.0}


class _ConvDecoder(ObservationModel_base):
    """fc([h, s]) -> stride-2 ConvTranspose2d stacks of the reference's image decoders (observation_model.py:56-345): LAYERS =
    (out channels or None for the image's, kernel).  normalization None: ConvT + ReLU pairs, the last ConvT without ReLU;
    "BatchNorm": ConvT(bias=False) + BatchNorm2d + ReLU triples, the last ConvT with bias.  64x64 / 128x128 without normalisation
    run on the plane tcgen05 kernels (bf16 mode) / exact SIMT kernels, every other variant on the general NCHW kernels."""
    __constants__ = ["embedding_size"]
    LAYERS = ()
    FC_NAME = "fc1"
    PLANE_PATH = False

    def __init__(self, belief_size, state_size, embedding_size, activation_function="relu", image_dim=3,
                 normalization=None):
        super().__init__()
        if normalization not in (None, "BatchNorm"):
            raise NotImplementedError(f"normalization={normalization!r}: None and BatchNorm are implemented (InstanceNorm / GroupNorm "
                                      "variants of the 256x256 stack are not)")
        self.embedding_size = embedding_size
        self.normalization = normalization
        setattr(self, self.FC_NAME, nn.Linear(belief_size + state_size, embedding_size))
        layers, cin = [], embedding_size
        for i, (cout, k) in enumerate(self.LAYERS):
            cout = image_dim if cout is None else cout
            last = i == len(self.LAYERS) - 1
            if normalization == "BatchNorm" and not last:
                layers += [nn.ConvTranspose2d(cin, cout, k, stride=2, bias=False), nn.BatchNorm2d(cout, affine=True, track_running_stats=True),
                           nn.ReLU()]
            else:
                layers.append(nn.ConvTranspose2d(cin, cout, k, stride=2))
                if not last:
                    layers.append(nn.ReLU())
            cin = cout
        self.conv = nn.Sequential(*layers)               # parameter / buffer container
        self.modules = [self._fc, self.conv]

    @property
    def _fc(self):
        return getattr(self, self.FC_NAME)

    def _forward_generic(self, h_t, s_t):
        T, B = h_t.shape[:2]
        x = ops.MlpFn.apply(0, False, 2, h_t.reshape(T * B, -1), s_t.reshape(T * B, -1), self._fc.weight, self._fc.bias)
        x = x.reshape(-1, self.embedding_size, 1, 1)
        mods = [m for m in self.conv if not isinstance(m, nn.ReLU)]
        i = 0
        while i < len(mods):
            ct = mods[i]
            x = ops.conv_transpose2d_nobias(x, ct.weight, 2, 0)
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d):
                x = ops.batch_norm(x, mods[i + 1], relu=True)
                i += 2
            else:
                x = ops.add_channel_bias(x, ct.bias, relu=i < len(mods) - 1)
                i += 1
        return {"loc": x.reshape(T, B, *x.shape[1:]), "scale": 1.0}

    def forward(self, h_t, s_t):
        if self.normalization is not None or not self.PLANE_PATH:
            return self._forward_generic(h_t, s_t)
        T, B = h_t.shape[:2]
        params = [self._fc.weight, self._fc.bias]
        params += [p for m in self.conv if isinstance(m, nn.ConvTranspose2d) for p in (m.weight, m.bias)]
        fn = ops.ConvDecoderTCFn if ops.bf16_mode() else ops.ConvDecoderFn
        y = fn.apply(h_t.reshape(T * B, -1), s_t.reshape(T * B, -1), *params)
        return {"loc": y.reshape(T, B, *y.shape[1:]), "scale": 1.0}

    def mse_loss(self, h_t, s_t, o_t):
        """sum_features mean_{t,b} (loc - o)^2.  bf16 mode, <= 4 image channels: decoder and loss are ONE autograd node — the
        last ConvTranspose2d's epilogue reduces the loss and keeps the bf16 residual, the reconstruction is never written."""
        if not (ops.bf16_mode() and o_t.shape[2] <= 4) or self.normalization is not None or not self.PLANE_PATH:
            return super().mse_loss(h_t, s_t, o_t)
        T, B = h_t.shape[:2]
        params = [self._fc.weight, self._fc.bias]
        params += [p for m in self.conv if isinstance(m, nn.ConvTranspose2d) for p in (m.weight, m.bias)]
        return ops.ConvDecoderMseTCFn.apply(h_t.reshape(T * B, -1), s_t.reshape(T * B, -1),
                                            o_t.reshape(T * B, *o_t.shape[2:]), *params)


class ImageDecoder(_ConvDecoder):
    """64x64: fc -> ConvT 1024->128 k5 -> 64 k5 -> 32 k6 -> C k6, stride 2 (reference :58-105)."""
    LAYERS = ((128, 5), (64, 5), (32, 6), (None, 6))
    PLANE_PATH = True


class ImageDecoder_84(_ConvDecoder):
    """84x84: fc -> ConvT 1024->128 k3 -> 64 k4 -> 32 k4 -> 16 k6 -> C k6 (reference :108-160; its Linear is named `fc`)."""
    LAYERS = ((128, 3), (64, 4), (32, 4), (16, 6), (None, 6))
    FC_NAME = "fc"


class ImageDecoder_128(_ConvDecoder):
    """128x128: fc -> ConvT 1024->256 k6 -> 128 k4 -> 64 k4 -> 32 k4 -> C k6 (reference :162-229)."""
    LAYERS = ((256, 6), (128, 4), (64, 4), (32, 4), (None, 6))
    PLANE_PATH = True


class ImageDecoder_256(_ConvDecoder):
    """256x256: fc -> ConvT 1024->256 k6 -> 128 k4 -> 64 k4 -> 32 k4 -> 16 k4 -> C k6 (reference :231-345)."""
    LAYERS = ((256, 6), (128, 4), (64, 4), (32, 4), (16, 4), (None, 6))


class SoundDecoder_v2(ObservationModel_base):
    """(belief, state) -> spectrogram [T,B,128,20] (reference observation_model.py:420-472): k = 1 Conv1d up-conversion, three
    bias-free ConvTranspose2d + InstanceNorm2d (tracked statistics) + GLU stages, a 7x7 Conv2d.  `forward(s_t, h_t)` keeps the
    reference's argument names: every caller passes (beliefs, states) positionally, so the concatenation is [states, beliefs]
    (SURVEY Q14).  Exact fp32 kernels (csrc/generic_nchw.cu)."""

    def __init__(self, belief_size, state_size, channels_base=128):
        super().__init__()
        cb = channels_base
        self.state_size, self.belief_size, self.channels_base = state_size, belief_size, cb
        inorm = lambda c: nn.InstanceNorm2d(num_features=c, affine=True, track_running_stats=True)
        self.up_conversion = nn.Conv1d(state_size + belief_size, int(cb * 2 * 32 * 4), kernel_size=1, bias=False)
        self.up_sample_0 = nn.Sequential(nn.ConvTranspose2d(cb * 2, cb * 4, kernel_size=(3, 4), stride=(1, 1), padding=(1, 1), bias=False),
                                         inorm(cb * 4), nn.GLU(dim=1))
        self.up_sample_1 = nn.Sequential(nn.ConvTranspose2d(cb * 2, cb * 2, kernel_size=4, stride=2, padding=1, bias=False),
                                         inorm(cb * 2), nn.GLU(dim=1))
        self.up_sample_2 = nn.Sequential(nn.ConvTranspose2d(cb, cb, kernel_size=4, stride=2, padding=1, bias=False), inorm(cb), nn.GLU(dim=1))
        self.out = nn.Conv2d(cb // 2, 1, kernel_size=7, stride=1, padding=3, bias=False)
        self.modules = [self.up_conversion, self.up_sample_0, self.up_sample_1, self.up_sample_2, self.out]

    def forward(self, s_t, h_t):
        T, B = h_t.shape[:2]
        x = torch.cat([h_t.reshape(T * B, -1, 1), s_t.reshape(T * B, -1, 1)], dim=1)
        x = ops.conv1d_k1(x, self.up_conversion.weight).view(-1, int(self.channels_base * 2), 32, 4)
        for stage in (self.up_sample_0, self.up_sample_1, self.up_sample_2):
            c = stage[0]
            x = ops.GluFn.apply(ops.instance_norm(ops.conv_transpose2d_nobias(x, c.weight, c.stride, c.padding), stage[1]))
        x = ops.conv2d_nobias(x, self.out.weight, 1, 3).squeeze(1)
        return {"loc": x.reshape(T, B, *x.shape[1:]), "scale": 1.0}


def build_ObservationModel(name, observation_shapes, belief_size, state_size, hidden_size, embedding_size,
                           activation_function, normalization=None):
    shape = observation_shapes[name]
    if "image" in name:
        cls = {(64, 64): ImageDecoder, (84, 84): ImageDecoder_84, (128, 128): ImageDecoder_128, (256, 256): ImageDecoder_256}.get(tuple(shape[1:]))
        if cls is None:
            raise NotImplementedError(f"image size {list(shape[1:])}: the reference defines 64, 84, 128 and 256")
        return cls(belief_size, state_size, embedding_size["image"], activation_function["cnn"],
                   image_dim=shape[0], normalization=normalization)
    if "sound" in name:
        return SoundDecoder_v2(belief_size=belief_size, state_size=state_size)
    if name == "draw_target":
        raise NotImplementedError(f"decoder for '{name}' is outside the B200 hot path (SURVEY §2/§8f)")
    return DenseDecoder(shape[0], belief_size, state_size, embedding_size["other"], activation_function["dense"])


class MultimodalObservationModel:
    """One decoder per name in observation_names_rec (reference observation_model.py:537-611)."""
    __constants__ = ["embedding_size"]

    def __init__(self, observation_names_rec, observation_shapes, embedding_size, belief_size, state_size,
                 hidden_size, activation_function, normalization=None, device=torch.device("cpu")):
        self.observation_names_rec = observation_names_rec
        self.observation_models = {}
        self.modules = []
        for name in observation_names_rec:
            m = build_ObservationModel(name, observation_shapes, belief_size, state_size, hidden_size,
                                       embedding_size, activation_function, normalization=normalization).to(device)
            self.observation_models[name] = m
            self.modules += m.modules

    def forward(self, h_t, s_t):
        return {name: m(h_t, s_t) for name, m in self.observation_models.items()}

    __call__ = forward

    def get_log_prob(self, h_t, s_t, o_t):
        return {n: self.observation_models[n].get_log_prob(h_t, s_t, o_t[n]) for n in self.observation_names_rec}

    def get_mse(self, h_t, s_t, o_t):
        return {n: self.observation_models[n].get_mse(h_t, s_t, o_t[n]) for n in self.observation_names_rec}

    def mse_loss(self, h_t, s_t, o_t):
        return {n: self.observation_models[n].mse_loss(h_t, s_t, o_t[n]) for n in self.observation_names_rec}

    def get_pred_value(self, h_t, s_t, key):
        return self.observation_models[key](h_t, s_t)

    def get_pred_key(self, h_t, s_t, key):
        return self.get_pred_value(h_t, s_t, key)

    def get_state_dict(self):
        return {name: m.state_dict() for name, m in self.observation_models.items()}

    def _load_state_dict(self, state_dict):
        for name in self.observation_names_rec:
            self.observation_models[name].load_state_dict(state_dict[name])

    def get_model_params(self):
        return [p for m in self.observation_models.values() for p in m.parameters()]

    def eval(self):
        for m in self.observation_models.values():
            m.eval()

    def train(self):
        for m in self.observation_models.values():
            m.train()
