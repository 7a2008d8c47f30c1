"""Noise source for the three reparameterisation streams.

The kernels never generate random numbers: every rsample consumes a caller-supplied N(0,1) tensor.
By default the host draws them with torch.randn on the device, one tensor per stream per call.
Tests (and anyone needing bit-identical noise with the reference) install a source with
``set_noise_source(fn)`` where ``fn(stream, shape, device) -> tensor``; ``stream`` is one of
"prior", "post", "dec" ("prior"/"post": [T-1,B,S] consumed inside the rollout in the reference's
per-step order prior then posterior; "dec": the post-rollout decoder-latent resample).
"""
import torch

_source = None


def set_noise_source(fn):
    global _source
    prev, _source = _source, fn
    return prev


def draw(stream, shape, device):
    if _source is not None:
        t = _source(stream, tuple(shape), device)
        assert tuple(t.shape) == tuple(shape), (stream, t.shape, shape)
        return t.to(device=device, dtype=torch.float32)
    return torch.randn(shape, device=device, dtype=torch.float32)


class FixedNoise:
    """Context manager serving prepared tensors: FixedNoise(prior=..., post=..., dec=...)."""

    def __init__(self, **streams):
        self.streams = {k: [v] if torch.is_tensor(v) else list(v) for k, v in streams.items() if v is not None}

    def __call__(self, stream, shape, device):
        q = self.streams.get(stream)
        if not q:
            raise RuntimeError(f"no prepared noise for stream '{stream}' {shape}")
        return q.pop(0)

    def __enter__(self):
        self.prev = set_noise_source(self)
        return self

    def __exit__(self, *exc):
        set_noise_source(self.prev)
