"""Training driver over device-resident replay buffers.

Interface of the reference's algos/MRSSM/MRSSM/train.py (`get_dataset_loader`, `train`, `run`; :9-66): two buffers are filled
from episode directories, then `optimize` runs `train_iteration` times with `validation` every `validation_interval` and a
checkpoint every `checkpoint_interval` iterations.  The experiment logger (hydra / wandb, utils/logger.py) is control plane
and not part of this package, so `run` takes the three values it would hand over.  One process drives one GPU; under
torchrun pass `dp=mrssm_b200.dist.DataParallel` to `run` / `train`: numpy's global RNG (chunk starts, augmentation choices)
is then re-seeded per rank (`cfg.main.seed + RANK`) so the ranks draw different batches, and only rank 0 writes checkpoints
(`save_model`)."""
import os

import numpy as np
import torch

from algos.MRSSM.MRSSM.algo import build_RSSM
from utils.replay_buffer.memory import ExperienceReplay_Multimodal, load_dataset


def _buffer_options(cfg):
    """Constructor arguments of the replay buffer, gathered from the three config groups that hold them."""
    aug, env, rssm = cfg.train.augmentation, cfg.env, cfg.rssm
    wanted = dict.fromkeys(list(rssm.observation_names_enc) + list(rssm.observation_names_rec))     # union, first-seen order
    opts = {key: aug[key] for key in ("n_crop", "dh_base", "dw_base", "noise_scales", "pca_scales")}
    opts.update(size=cfg.train.experience_size, observation_names=list(wanted), observation_shapes=env.observation_shapes,
                action_name=env.action_name, action_size=env.action_size, bit_depth=env.bit_depth)
    return opts


def get_dataset_loader(cfg, cwd, device, dataset_path):
    """A buffer on `device` holding every episode found under dataset_path (one directory or a list of them)."""
    buffer = ExperienceReplay_Multimodal(device=device, **_buffer_options(cfg))
    buffer.seed = int(os.environ.get("RANK", 0))        # in-kernel dequantisation / Gaussian noise streams differ per rank
    load_dataset(cfg, cwd, buffer, dataset_path)
    return buffer


def _restore(model, cfg, cwd):
    if cfg.train.model_path is None:
        return
    path = os.path.join(cwd, cfg.train.model_path)
    if not os.path.exists(path):
        raise NotImplementedError("{} is not exist".format(path))
    model.load_model(path)


def _due(itr, every):
    return bool(every) and itr % every == 0


def train(cfg, cwd, results_dir, device, dp=None):
    buffers = {split: get_dataset_loader(cfg, cwd, device, cfg.train[split + "_data_path"]) for split in ("train", "validation")}
    model = build_RSSM(cfg, device)
    _restore(model, cfg, cwd)
    if dp is not None:
        dp(model)
        rank = int(os.environ.get("RANK", 0))
        np.random.seed(int(cfg.main.get("seed", 0) or 0) + rank)       # identical seeds would make N ranks compute one rank's gradient
    last = cfg.train.train_iteration
    for itr in range(1, last + 1):
        model.optimize(buffers["train"])
        if _due(itr, cfg.train.validation_interval):
            model.validation(buffers["validation"])
        if _due(itr, cfg.train.checkpoint_interval):
            model.save_model(results_dir, itr)
    return model


def run(cfg, cwd=None, results_dir=None, device=None, dp=None):
    """Reference run(cfg) (train.py:58-66) minus the hydra / wandb logger set-up: seeds as utils/logger.py:93-100 does."""
    cwd = cwd or os.getcwd()
    results_dir = results_dir or cwd
    os.makedirs(results_dir, exist_ok=True)
    seed = int(cfg.main.get("seed", 0) or 0)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    return train(cfg, cwd, results_dir, device or torch.device(cfg.main.device), dp=dp)
