"""Small host-side helpers the evaluation notebooks use on the arrays `estimate_states.get_states` returns
(interface of the reference's utils/evaluation/visualize_utils.py).  Nothing here touches the device."""
import numpy as np
import torch

from utils.processing.image_processing import reverse_normalized_image


def np2tensor(data, dtype=torch.float32):
    return data if torch.is_tensor(data) else torch.as_tensor(np.asarray(data), dtype=dtype)


def tensor2np(tensor):
    return tensor.detach().cpu().numpy().copy() if torch.is_tensor(tensor) else tensor


def reverse_image_observation(image, bit_depth=5):
    """Normalised CHW frame -> displayable HWC uint8."""
    return reverse_normalized_image(tensor2np(image), bit_depth=bit_depth).transpose(1, 2, 0)


def flat(feat):
    return feat.reshape(-1, feat.shape[-1])


def get_xyz(feat):
    f = flat(feat)
    return f[:, 0], f[:, 1], f[:, 2]


def get_pca_model(feat, n_components=3):
    """PCA over all (t, b) rows of a latent tensor, for 3-d scatter plots of beliefs / states."""
    from sklearn.decomposition import PCA
    rows = flat(tensor2np(feat))
    print(rows.shape)
    return PCA(n_components=n_components).fit(rows)
