"""Training driver (reference algos/MRSSM/MRSSM/train.py): replay buffers from episode directories, then
optimize / validation / checkpoint on the reference's schedule.  The experiment logger (hydra / wandb set-up,
utils/logger.py) is control plane and not part of this package: `run` takes the three values it would produce."""
import os

import torch

from utils.replay_buffer.memory import ExperienceReplay_Multimodal, load_dataset

from algos.MRSSM.MRSSM.algo import build_RSSM


def get_dataset_loader(cfg, cwd, device, dataset_path):
    observation_names = list(set(list(cfg.rssm.observation_names_enc) + list(cfg.rssm.observation_names_rec)))
    aug = cfg.train.augmentation
    D = ExperienceReplay_Multimodal(size=cfg.train.experience_size, observation_names=observation_names,
                                    observation_shapes=cfg.env.observation_shapes, n_crop=aug.n_crop, dh_base=aug.dh_base,
                                    dw_base=aug.dw_base, noise_scales=aug.noise_scales, pca_scales=aug.pca_scales,
                                    action_name=cfg.env.action_name, action_size=cfg.env.action_size,
                                    bit_depth=cfg.env.bit_depth, device=device)
    load_dataset(cfg, cwd, D, dataset_path)
    return D


def train(cfg, cwd, results_dir, device, dp=None):
    """dp: optional mrssm_b200.dist.DataParallel factory applied to the model (one process per GPU)."""
    print("Initialize training environment and experience replay memory")
    D = get_dataset_loader(cfg, cwd, device, cfg.train.train_data_path)
    D_val = get_dataset_loader(cfg, cwd, device, cfg.train.validation_data_path)
    print("Initialise model parameters randomly")
    model = build_RSSM(cfg, device)
    if cfg.train.model_path is not None:
        model_path = os.path.join(cwd, cfg.train.model_path)
        if not os.path.exists(model_path):
            raise NotImplementedError("{} is not exist".format(model_path))
        model.load_model(model_path)
    if dp is not None:
        dp(model)
    for itr in range(1, cfg.train.train_iteration + 1):
        model.optimize(D)
        if itr % cfg.train.validation_interval == 0:
            model.validation(D_val)
        if itr % cfg.train.checkpoint_interval == 0:
            model.save_model(results_dir, itr)
    return model


def run(cfg, cwd=None, results_dir=None, device=None):
    cwd = os.getcwd() if cwd is None else cwd
    results_dir = cwd if results_dir is None else results_dir
    device = torch.device(cfg.main.device) if device is None else device
    os.makedirs(results_dir, exist_ok=True)
    return train(cfg, cwd, results_dir, device)
