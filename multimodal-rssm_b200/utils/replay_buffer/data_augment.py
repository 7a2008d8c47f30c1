"""Augmentation parameters of the replay buffer (reference utils/replay_buffer/data_augment.py).

Only the HOST side lives here — which crop window, which colour shift, which noise scale a batch gets.  Applying them is
part of the fused gather kernel (`mrssm_replay_gather_u8`): crop = origin of the read window, colour shift = one offset per
channel, Gaussian noise = added before the clip."""
import numpy as np
import torch


def spiral_offset(idx):
    """(dx, dy) of crop position `idx` (reference get_dx / get_dy :96-146): positions wind outwards from the centre,
        |12|13|14|15|
        |11| 2| 3| 4|
        |10| 1| 0| 5|
        | 9| 8| 7| 6|
    i.e. 1 step left, 1 up, 2 right, 2 down, 3 left, 3 up, ..."""
    x = y = 0
    step = 1
    left = idx
    moves = ((-1, 0), (0, -1), (1, 0), (0, 1))
    m = 0
    while left > 0:
        for _ in range(2):                      # each run length is used for two consecutive directions
            run = min(step, left)
            x += moves[m][0] * run
            y += moves[m][1] * run
            left -= run
            m = (m + 1) % 4
            if left == 0:
                break
        step += 1
    return x, y


def crop_origin(idx, stored_hw, size, dh_base, dw_base):
    """Top-left corner (dh, dw) of crop `idx` (reference idx_to_idx_w_h + crop_image :148-166).  As in the reference the
    centre of the W axis is derived from the H extent / dh_base and vice versa (they coincide for the square frames used)."""
    dx, dy = spiral_offset(idx)
    centre = (np.array(stored_hw) - np.array(size)) / (dh_base, dw_base)
    cx, cy = np.floor(centre / 2)
    return dh_base * int(cy + dy), dw_base * int(cx + dx)


def crop_size_of(name):
    """Side of the training crop for an image modality, chosen by name (reference :183-194)."""
    if "_256" in name or "high_resolution" in name:
        return 256
    if "_128" in name:
        return 128
    return 64


def crop_image(image, idx=0, size=(64, 64), dh_base=2, dw_base=2):
    """Host-side crop of [..., H, W] arrays / tensors (used when episodes are stored, reference :158-174)."""
    dh, dw = crop_origin(idx, image.shape[-2:], size, dh_base, dw_base)
    if dh < 0 or dw < 0 or dh + size[0] > image.shape[-2] or dw + size[1] > image.shape[-1]:
        raise ValueError(f"crop {idx} ({dh},{dw})+{tuple(size)} leaves the {tuple(image.shape[-2:])} frame")
    return image[..., dh:dh + size[0], dw:dw + size[1]]


def crop_image_data(data, n_crop=None, dh_base=None, dw_base=None):
    """Episode images are stored with a margin of k = int(sqrt(n_crop - 1)) crop steps (reference :212-230)."""
    if n_crop is not None:
        k = int(np.sqrt(n_crop - 1))
        for name in data.keys():
            if "image" in name:
                side = crop_size_of(name)
                data[name] = crop_image(data[name], idx=0, size=(side + k * dh_base, side + k * dw_base),
                                        dh_base=dh_base, dw_base=dw_base)
    return data


def calc_params_of_pca(image, dt=100):
    """Eigen-decomposition used by the PCA colour augmentation (reference :53-63): every dt-th stored frame, flattened to
    three rows (the reference's plain reshape, not a channel split), standardised, 3x3 covariance, eigh."""
    x = image[::dt].cpu().reshape(3, -1).to(torch.float32)      # a handful of frames: done on the host, once per load
    x = (x.transpose(1, 0) - torch.mean(x, dim=1)) / torch.std(x, dim=1)
    xm = x - torch.mean(x, dim=0)
    cov = torch.mm(xm.t(), xm) / (x.shape[0] - 1)               # columns are the variables, unbiased
    return torch.linalg.eigh(cov, UPLO="U")


def pca_delta(lambd_eigen_value, p_eigen_vector, rand):
    """Per-channel colour offset in 0..255 units (reference calc_delta :65-69).  Three numbers: computed on the host."""
    return torch.matmul(p_eigen_vector.cpu(), rand.cpu() * lambd_eigen_value.cpu()) * 255.0


def draw_pca_rand(pca_scales):
    """The three N(0, scale) coefficients of one batch; consumes numpy's global RNG like the reference (:71-79)."""
    scale = pca_scales[np.random.randint(0, len(pca_scales))]
    if scale > 0:
        return np.random.randn(3) * scale
    return np.zeros(3)
