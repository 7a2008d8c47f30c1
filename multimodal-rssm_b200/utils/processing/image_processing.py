"""Image quantisation helpers (reference utils/processing/image_processing.py).

`normalize_image` of the reference runs on an fp32 copy of the frames; here frames stay uint8 until the kernel that
consumes them (`mrssm_normalize_image_u8`, or fused into the replay gather `mrssm_replay_gather_u8`)."""
import numpy as np
import torch

from mrssm_b200 import _lib as L


def normalize_image_u8(frames, bit_depth, noise=None, seed=0):
    """uint8 CUDA tensor -> fp32 in [-0.5, 0.5): floor(x / 2^(8-bits)) / 2^bits - 0.5 + u / 2^bits  (reference :5-11).
    noise: optional U[0,1) tensor of the same shape (otherwise drawn in the kernel from `seed`)."""
    assert frames.dtype == torch.uint8
    frames = frames.contiguous()
    out = torch.empty(frames.shape, device=frames.device, dtype=torch.float32)
    L.call("mrssm_normalize_image_u8", L.ptr_any(frames), frames.numel(), int(bit_depth),
           None if noise is None else L.ptr(noise.contiguous()), int(seed), L.ptr(out))
    return out


def reverse_normalized_image(observation, bit_depth=5):
    """fp32 array in [-0.5, 0.5] -> uint8 storage format (reference :15-16).  Host side (numpy)."""
    levels = np.floor((np.asarray(observation) + 0.5) * 2 ** bit_depth) * 2 ** (8 - bit_depth)
    return np.clip(levels, 0, 2 ** 8 - 1).astype(np.uint8)
