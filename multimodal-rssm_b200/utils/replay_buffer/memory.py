"""Experience replay with the frame store in HBM (reference utils/replay_buffer/memory.py).

Same interface as the reference's `ExperienceReplay_Multimodal` (constructor arguments, `load_dataset`, `append`, `sample`,
the public counters and the `observations / actions / rewards / nonterminals` stores that utils/evaluation reads), same
on-disk episode format (one pickled dict per `.npy`: `<observation names>`, `<action name>`, `reward`, `done`; images
uint8 HWC or CHW, or fp32 in [-0.5, 0.5]) and the same consumption of numpy's global RNG for chunk starts and augmentation
choices, so a seeded run picks the same chunks.

What differs is where the work happens.  The reference keeps the stores on the host and, per `sample`, gathers with fancy
indexing, converts the frames to fp32, copies ~4 bytes per pixel to the device and then runs crop / colour shift / noise /
clip / quantise / dequantise as separate passes.  Here the stores live on `device` (uint8 frames: 12 KB per 3x64x64
frame, so 10^6 frames take 12 GB of the 180 GB), `sample` ships only the L*n int64 slot numbers, and one kernel per image
modality (`mrssm_replay_gather_u8`) reads each stored byte once and writes the normalised fp32 batch once; vectors,
actions, rewards and nonterminals go through `mrssm_gather_rows`.
"""
import ctypes
import glob
import os

import numpy as np
import torch

from mrssm_b200 import _lib as L
from utils.processing.image_processing import reverse_normalized_image
from utils.replay_buffer.data_augment import (calc_params_of_pca, crop_image_data, crop_origin, crop_size_of, draw_pca_rand,
                                              pca_delta)

try:                                    # dataset lists come as omegaconf ListConfig in the reference's hydra configs
    from omegaconf import ListConfig as _ListConfig
    _LIST_TYPES = (list, tuple, _ListConfig)
except ImportError:                     # omegaconf is optional here
    _LIST_TYPES = (list, tuple)


# ---- episode files (reference :13-110) -----------------------------------------------------------------------------
def _load_dataset(cfg, cwd, dataset_path, D):
    dataset_dir = os.path.join(cwd, dataset_path)
    if not os.path.exists(dataset_dir):
        raise NotImplementedError("{} is not exist".format(dataset_dir))
    print("load dataset from {}".format(dataset_dir))
    if not os.path.isdir(dataset_dir):
        raise NotImplementedError("{}: only episode directories are supported".format(dataset_dir))
    D.load_dataset(dataset_dir=dataset_dir)


def load_dataset(cfg, cwd, D, dataset_path):
    """dataset_path: one directory or a list of directories, relative to cwd (reference :26-32)."""
    if isinstance(dataset_path, str):
        _load_dataset(cfg, cwd, dataset_path, D)
    elif isinstance(dataset_path, _LIST_TYPES):
        for path in dataset_path:
            _load_dataset(cfg, cwd, path, D)


def clip_episode(data):
    """Cut every stream to the shortest one; the 'seed' entry is dropped (reference :34-45)."""
    streams = {k: v for k, v in data.items() if k != "seed"}
    episode_length = int(min(len(v) for v in streams.values()))
    return {k: v[:episode_length] for k, v in streams.items()}, episode_length


def preprocess_data(data):
    """Storage format: images CHW uint8, 'nonterminals' = 1 - done (reference :47-65)."""
    data, episode_length = clip_episode(data)
    for name in data.keys():
        if "image" not in name:
            continue
        if data[name].shape[1] > data[name].shape[3]:                 # HWC on disk
            data[name] = data[name].transpose(0, 3, 1, 2)
        if data[name].dtype != np.uint8:                              # normalised floats on disk
            data[name] = reverse_normalized_image(data[name])
    if "image" in data:
        side = data["image"].shape[2]
        if side != 64:
            data["image_{}".format(side)] = data.pop("image")
    data["nonterminals"] = 1.0 - np.expand_dims(data["done"], -1)
    return data, episode_length


def calc_image_shape(shape, n_crop=None, dw_base=2, dh_base=2):
    """Stored frame shape: the crop margin of k = int(sqrt(n_crop - 1)) steps is kept (reference :67-74)."""
    if n_crop is None:
        return shape
    d, h, w = shape
    k = int(np.sqrt(n_crop - 1))
    return [d, int(h + k * dh_base), int(w + k * dw_base)]


def np2tensor(data, device=torch.device("cpu")):
    if torch.is_tensor(data):
        return data.to(device)
    return torch.as_tensor(np.ascontiguousarray(data), dtype=torch.uint8 if data.dtype == np.uint8 else torch.float32).to(device)


def get_file_names(dataset_dir):
    return glob.glob(os.path.join(dataset_dir, "*.npy"))


def get_data(file_name, n_crop=1, dh_base=1, dw_base=1, encoding="ASCII"):
    raw = np.load(file_name, allow_pickle=True, encoding=encoding).item()
    if encoding != "ASCII":
        raw = {key.decode("utf-8"): v for key, v in raw.items()}
    data, episode_length = preprocess_data(raw)
    data = crop_image_data(data, n_crop=n_crop, dh_base=dh_base, dw_base=dw_base)
    return {k: v[:episode_length] for k, v in data.items()}, episode_length


# ---- the buffer ---------------------------------------------------------------------------------------------------
class ExperienceReplay_Multimodal:
    _instances = 0

    def __init__(self, size, observation_names=["image"], observation_shapes=dict(image=[3, 64, 64]), n_crop=None,
                 dh_base=None, dw_base=None, noise_scales=None, pca_scales=None, action_name="action", action_size=None,
                 bit_depth=5, device=torch.device("cpu")):
        self.device = torch.device(device)
        self.size = size
        self.observation_names = observation_names
        self.observation_shapes = observation_shapes
        self.action_name = action_name
        self.action_size = action_size
        self.file_names = []
        self.idx = 0
        self.full = False                   # every slot holds valid experience
        self.steps, self.episodes = 0, 0
        self.bit_depth = bit_depth
        self.n_crop, self.dh_base, self.dw_base = n_crop, dh_base, dw_base
        self.noise_scales = noise_scales
        self.pca_scales = pca_scales
        self.lambd_eigen_values = {name: None for name in observation_names}
        self.p_eigen_vectors = {name: None for name in observation_names}
        self.noise_source = None            # tests: fn(kind, name, shape) -> tensor for kind in {"uniform", "gauss"}
        self._draws = 0                     # counter mixed into the kernels' hash seed
        self.seed = 0
        ExperienceReplay_Multimodal._instances += 1
        self._salt = ExperienceReplay_Multimodal._instances      # train / validation buffers with one seed still draw different noise
        self._init_buffer(size)

    def _init_buffer(self, size):
        dev = self.device
        self.observations = {}
        for name in self.observation_names:
            if "image" in name:
                shape = calc_image_shape(self.observation_shapes[name], self.n_crop, self.dw_base, self.dh_base)
                self.observations[name] = torch.empty((size, *shape), dtype=torch.uint8, device=dev)
            else:
                self.observations[name] = torch.empty((size, *self.observation_shapes[name]), dtype=torch.float32, device=dev)
        self.actions = torch.empty((size, self.action_size), dtype=torch.float32, device=dev)
        self.rewards = torch.empty((size,), dtype=torch.float32, device=dev)
        self.nonterminals = torch.empty((size, 1), dtype=torch.float32, device=dev)

    # ---- sampling (reference :175-222) ---------------------------------------------------------------------------
    def _sample_start(self, L_, idx_max=None):
        """First slot of one chunk of L consecutive steps that does not run over the write position: one
        np.random.randint per attempt, exactly like the reference's rejection loop (:177-189)."""
        upper = self.size if self.full else self.idx - L_
        if idx_max is not None:
            upper = np.min([idx_max, upper])
        while True:
            start = np.random.randint(0, upper)
            # slots start+1 .. start+L-1 (mod size) must not contain the write position
            if (self.idx - start - 1) % self.size >= L_ - 1:
                return start

    def _sample_chunks(self, n, L_):
        """Slots [n, L] of n chunks: the draws stay sequential (RNG parity), the slot table is built in one step."""
        starts = np.asarray([self._sample_start(L_) for _ in range(n)])
        return (starts[:, None] + np.arange(L_)[None, :]) % self.size

    def _sample_idx(self, L_, idx_max=None):
        """Slots of one chunk (the reference's helper, kept for callers that use it directly)."""
        return (self._sample_start(L_, idx_max) + np.arange(L_)) % self.size

    def _noise(self, kind, name, shape):
        if self.noise_source is None:
            return None
        t = self.noise_source(kind, name, tuple(shape))
        return None if t is None else t.to(device=self.device, dtype=torch.float32).contiguous()

    def _require_cuda(self):
        if self.device.type != "cuda":
            # the constructor keeps the reference's host default so that loading / filling can be exercised anywhere, but
            # sampling is a CUDA gather kernel and the product has no CPU path (oracle/replay_oracle.py is test infrastructure)
            raise RuntimeError(f"ExperienceReplay_Multimodal.sample: the store is on {self.device}; construct the buffer with "
                               "device=torch.device('cuda:<n>') — the B200 replay buffer samples on the GPU only")

    def _gather_image(self, name, slots, rows, crop, delta, gauss_scale, normalise):
        self._require_cuda()
        store = self.observations[name]
        C, Hs, Ws = store.shape[1:]
        side = crop_size_of(name) if crop is not None else None
        H, W = (side, side) if crop is not None else (Hs, Ws)
        dh, dw = crop if crop is not None else (0, 0)
        out = torch.empty((rows, C, H, W), device=self.device, dtype=torch.float32)
        a = L.ReplayGatherArgs()
        a.frames, a.idx, a.rows = L.ptr_any(store), L.ptr_any(slots), rows
        a.C, a.Hs, a.Ws, a.H, a.W, a.dh, a.dw = C, Hs, Ws, H, W, dh, dw
        a.bit_depth = int(self.bit_depth) if normalise else 0
        gauss = self._noise("gauss", name, out.shape) if gauss_scale > 0 else None
        uniform = self._noise("uniform", name, out.shape) if normalise else None
        a.delta, a.out = L.ptr(delta), L.ptr(out)
        a.gauss, a.gauss_scale, a.uniform = L.ptr(gauss), float(gauss_scale), L.ptr(uniform)
        self._draws += 1
        a.seed = ((int(self.seed) * 0x9E3779B1 + self._salt) * 0x85EBCA6B + self._draws) & 0xFFFFFFFFFFFFFFFF
        with torch.cuda.device(self.device):          # launch on the buffer's device, whichever device is current
            L.call("mrssm_replay_gather_u8", ctypes.byref(a))
        return out

    def _gather_rows(self, store, slots, rows):
        self._require_cuda()
        K = store.numel() // store.shape[0]
        out = torch.empty((rows, K), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.call("mrssm_gather_rows", L.ptr(store), L.ptr_any(slots), rows, K, L.ptr(out))
        return out

    def _plan_batch(self, idxs, crop_idx=None, pca_rand=None):
        """Host half of _retrieve_batch: the slot list (row = l * n + b) and, per image modality, the augmentation choices,
        drawn from numpy's global RNG in the reference's order (memory.py:191-208, data_augment.py:178-208).
        A given crop_idx / pca_rand (three coefficients) replaces the corresponding draw, as in the reference's
        augment_image_data (used by utils/evaluation/estimate_states.py).
        -> (vec_idxs int64 [L*n], {name: dict(crop=(dh, dw)|None, side, delta=[C] cpu tensor|None, gauss_scale, plain)})"""
        vec_idxs = np.ascontiguousarray(idxs.transpose().reshape(-1)).astype(np.int64)
        assert vec_idxs.min() >= 0 and vec_idxs.max() < self.size
        if pca_rand is not None:
            pca_rand = np.asarray(torch.as_tensor(pca_rand).cpu(), dtype=np.float64)
        plan = {}
        for name in self.observation_names:
            if "image" not in name:
                continue
            store = self.observations[name]
            side = crop_size_of(name)
            crop = None
            if self.n_crop is not None:
                pos = np.random.randint(0, self.n_crop) if crop_idx is None else crop_idx
                crop = crop_origin(pos, store.shape[-2:], (side, side), self.dh_base, self.dw_base)
            plain = "bin" in name                   # binary masks: cropped only, neither augmented nor normalised
            gauss_scale, delta = 0.0, None
            if not plain:
                if self.noise_scales is not None:
                    gauss_scale = float(self.noise_scales[np.random.randint(0, len(self.noise_scales))])
                if self.pca_scales is not None:
                    if pca_rand is None:            # one colour shift per batch, shared by the image modalities
                        pca_rand = draw_pca_rand(self.pca_scales)
                    if np.any(pca_rand != 0):
                        rand = torch.as_tensor(pca_rand, dtype=torch.float32)
                        delta = pca_delta(self.lambd_eigen_values[name], self.p_eigen_vectors[name], rand)
            plan[name] = dict(crop=crop, side=side, delta=delta, gauss_scale=gauss_scale, plain=plain)
        return vec_idxs, plan

    def _retrieve_batch(self, idxs, n, L_, crop_idx=None, pca_rand=None):
        vec_idxs, plan = self._plan_batch(idxs, crop_idx, pca_rand)
        rows = n * L_
        slots = torch.from_numpy(vec_idxs).to(self.device, non_blocking=True)
        observations = {}
        for name in self.observation_names:
            store = self.observations[name]
            if name not in plan:
                observations[name] = self._gather_rows(store, slots, rows).reshape(L_, n, *store.shape[1:])
                continue
            p = plan[name]
            delta = None if p["delta"] is None else p["delta"].to(self.device).contiguous()
            out = self._gather_image(name, slots, rows, p["crop"], delta, p["gauss_scale"], normalise=not p["plain"])
            observations[name] = out.reshape(L_, n, *out.shape[1:])
        actions = self._gather_rows(self.actions, slots, rows).reshape(L_, n, -1)
        rewards = self._gather_rows(self.rewards, slots, rows).reshape(L_, n)
        nonterminals = self._gather_rows(self.nonterminals, slots, rows).reshape(L_, n, 1)
        return observations, actions, rewards, nonterminals

    def sample(self, n, L_):
        """n chunks of L consecutive steps, time-major: [obs dict [L,n,...], actions [L,n,A], rewards [L,n],
        nonterminals [L,n,1]] on `device`."""
        return list(self._retrieve_batch(self._sample_chunks(n, L_), n, L_))

    # ---- filling (reference :224-268) ------------------------------------------------------------------------------
    def append(self, observation, action, reward, done):
        for name in self.observation_names:
            if "image" in name:
                frame = reverse_normalized_image(np.asarray(observation[name]), self.bit_depth)
                self.observations[name][self.idx] = torch.as_tensor(frame).to(self.device)
            else:
                self.observations[name][self.idx] = torch.as_tensor(observation[name], dtype=torch.float32).to(self.device)
        self.actions[self.idx] = torch.as_tensor(action, dtype=torch.float32).to(self.device)
        self.rewards[self.idx] = float(reward)
        self.nonterminals[self.idx] = float(not done)
        self.idx = (self.idx + 1) % self.size
        self.full = self.full or self.idx == 0
        self.steps, self.episodes = self.steps + 1, self.episodes + (1 if done else 0)

    def _set_data_to_buffer(self, file_name):
        data, episode_length = get_data(file_name, self.n_crop, self.dh_base, self.dw_base)
        if self.idx + episode_length > self.size:
            raise IndexError("episode of {} steps does not fit behind slot {} of {}".format(episode_length, self.idx, self.size))
        sl = slice(self.idx, self.idx + episode_length)
        dev = self.device
        for name in self.observation_names:
            self.observations[name][sl] = np2tensor(data[name], dev)
        if self.action_name == "dummy":
            self.actions[sl] = 0.0
        else:
            self.actions[sl] = np2tensor(data[self.action_name], dev)
        self.rewards[sl] = np2tensor(data["reward"], dev)
        self.nonterminals[sl] = np2tensor(data["nonterminals"], dev)
        self.full = self.full or (self.idx + episode_length) / self.size >= 1
        self.idx = (self.idx + episode_length) % self.size
        self.steps += episode_length
        self.episodes += 1

    def load_dataset(self, dataset_dir):
        file_names = get_file_names(dataset_dir)
        print("find %d npy files!" % len(file_names))
        self.file_names += file_names
        for file_name in file_names:
            self._set_data_to_buffer(file_name)
        if self.pca_scales is not None:
            print("set color augment params")
            self._set_color_aug_params()

    def _set_color_aug_params(self):
        self.lambd_eigen_values, self.p_eigen_vectors = {}, {}
        for name in self.observations.keys():
            if "image" in name and "bin" not in name:
                lam, vec = calc_params_of_pca(self.observations[name][:self.idx])
                self.lambd_eigen_values[name] = lam.to(self.device)
                self.p_eigen_vectors[name] = vec.to(self.device)
