"""Synthetic episode files in the reference's on-disk format (utils/replay_buffer/memory.py:90-110), and the two buffer
configurations the replay fixtures were generated with."""
import os

import numpy as np
import torch

IMAGE, BIN, VEC, ACTION = "image_horizon", "image_bin", "pose_quat_v2", "d_pose_quat_v2"

CONFIGS = {
    # the shipped defaults (train.yaml:22-27): one crop position, zero noise, zero colour shift
    "default": dict(side=64, names=[IMAGE, VEC], n_crop=1, dh_base=1, dw_base=1, noise_scales=[0.0], pca_scales=[0.0]),
    # everything on: 9 crop positions over 68x68 stored frames, Gaussian noise, PCA colour shift, a binary mask modality
    "augment": dict(side=72, names=[IMAGE, BIN, VEC], n_crop=9, dh_base=2, dw_base=2, noise_scales=[0.02, 0.05],
                    pca_scales=[0.1]),
    # the modality set of the shipped rssm YAML: image + sound spectrogram [128, 20]
    "shipped": dict(side=64, names=[IMAGE, "sound"], n_crop=1, dh_base=1, dw_base=1, noise_scales=[0.0], pca_scales=[0.0]),
}
SIZE, N, L = 40, 3, 4
EPISODES = [9, 7, 11]


def stored_side(cfg):
    return 64 + int(np.sqrt(cfg["n_crop"] - 1)) * cfg["dh_base"]


def shapes(cfg):
    out = {IMAGE: [3, 64, 64], VEC: [3]}
    if BIN in cfg["names"]:
        out[BIN] = [1, 64, 64]
    if "sound" in cfg["names"]:
        out["sound"] = [128, 20]
    return out


def write_dataset(root, cfg, seed=5):
    """Three episodes: uint8 HWC, uint8 CHW, and normalised-float HWC frames; one stream one step longer than the rest
    (clip_episode); a 'seed' entry (dropped by the loader).  Returns the file names in writing order."""
    rng = np.random.RandomState(seed)
    side = cfg["side"]
    os.makedirs(root, exist_ok=True)
    files = []
    for e, n in enumerate(EPISODES):
        hwc = rng.randint(0, 256, size=(n, side, side, 3)).astype(np.uint8)
        if e == 0:
            img = hwc
        elif e == 1:
            img = np.ascontiguousarray(hwc.transpose(0, 3, 1, 2))
        else:
            img = (np.floor(hwc / 8.0) / 32.0 - 0.5 + rng.rand(*hwc.shape) / 32.0).astype(np.float32)
        done = np.zeros(n, dtype=np.float32)
        done[-1] = 1.0
        data = {IMAGE: img, VEC: rng.randn(n + (1 if e == 1 else 0), 3).astype(np.float32),
                ACTION: rng.randn(n, 3).astype(np.float32), "reward": rng.randn(n).astype(np.float32), "done": done,
                "seed": np.arange(n + 3)}
        if BIN in cfg["names"]:
            data[BIN] = (rng.rand(n, 1, side, side) > 0.5).astype(np.uint8) * 255
        if "sound" in cfg["names"]:
            data["sound"] = rng.randn(n, 128, 20).astype(np.float32)
        path = os.path.join(root, "episode_%d.npy" % e)
        np.save(path, data, allow_pickle=True)
        files.append(path)
    return files


def buffer_kwargs(cfg, device):
    return dict(size=SIZE, observation_names=list(cfg["names"]), observation_shapes=shapes(cfg), n_crop=cfg["n_crop"],
                dh_base=cfg["dh_base"], dw_base=cfg["dw_base"], noise_scales=cfg["noise_scales"],
                pca_scales=cfg["pca_scales"], action_name=ACTION, action_size=3, bit_depth=5, device=device)


def digest(t):
    """Small stand-in for a full tensor: strided sample + sums."""
    import torch
    t = t.detach().cpu()
    flat = t.reshape(-1)
    return dict(shape=tuple(t.shape), sample=flat[::37].clone(), sum=float(flat.double().sum()),
                abs_sum=float(flat.double().abs().sum()))


def assert_digest(t, d, exact=True):
    import torch
    t = t.detach().cpu()
    assert tuple(t.shape) == tuple(d["shape"]), (t.shape, d["shape"])
    flat = t.reshape(-1)
    if exact:
        assert torch.equal(flat[::37], d["sample"])
        assert abs(float(flat.double().sum()) - d["sum"]) <= 1e-9 * max(1.0, d["abs_sum"])
        assert abs(float(flat.double().abs().sum()) - d["abs_sum"]) <= 1e-9 * max(1.0, d["abs_sum"])
    else:
        torch.testing.assert_close(flat[::37], d["sample"], rtol=1e-5, atol=1e-6)


def make_buffer(cfg, files, device="cpu"):
    from utils.replay_buffer.memory import ExperienceReplay_Multimodal
    D = ExperienceReplay_Multimodal(**buffer_kwargs(cfg, torch.device(device)))
    D.file_names += files
    for f in files:
        D._set_data_to_buffer(f)
    if D.pca_scales is not None:
        D._set_color_aug_params()
    return D


def load_fixture_buffer(name, golden_dir, tmp_path, device="cpu"):
    rec = torch.load(os.path.join(golden_dir, f"replay_{name}.pt"), weights_only=False)
    cfg = CONFIGS[name]
    files = write_dataset(str(tmp_path), cfg)
    assert [os.path.basename(f) for f in files] == rec["files"]
    return rec, cfg, make_buffer(cfg, files, device)
