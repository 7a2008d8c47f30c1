"""Latent states of every stored episode (reference utils/evaluation/estimate_states.py): the B = 1, T = episode-length
consumer of `estimate_state`.

An episode is a run of consecutive slots of the device-resident store, so its observations come out of the same fused
gather kernel as training batches (n = 1, L = T, fixed crop position, optional fixed colour shift); nothing is staged
through the host."""
import numpy as np
import torch

from mrssm_b200 import ops


def tensor2numpy_state(state):
    for key in state.keys():
        if state[key] is None:
            continue
        if "expert" in key:
            for k in state[key].keys():
                state[key][k] = state[key][k].detach().cpu().numpy()
        else:
            state[key] = state[key].detach().cpu().numpy()
    return state


def estimate_state(model, observations, actions, rewards, nonterminals):
    return model.estimate_state(observations, actions, rewards, nonterminals)


def get_all_data(D):
    observations = {name: D.observations[name][:D.idx] for name in D.observation_names}
    return observations, D.actions[:D.idx], D.rewards[:D.idx], D.nonterminals[:D.idx]


def episode_bounds(D):
    """Slot ranges of the stored episodes: an episode ends where nonterminals == 0 (reference :36-39)."""
    done = np.where(D.nonterminals[:D.idx, 0].cpu().numpy() == 0)[0]
    return np.hstack([[0], done + 1])


def get_episode_data(D, epi_idx, crop_idx=None, pca_rand=None):
    """-> observations {name: [T,1,...]}, actions [T,1,A], rewards [T,1], nonterminals [T,1,1] of one episode, augmented and
    normalised like a training batch (reference :35-58)."""
    bounds = episode_bounds(D)
    idx_start, idx_end = int(bounds[epi_idx]), int(bounds[epi_idx + 1])
    idxs = np.arange(idx_start, idx_end)[None, :]
    return D._retrieve_batch(idxs, 1, idx_end - idx_start, crop_idx=crop_idx, pca_rand=pca_rand)


def get_states(D, model, device, crop_idx=0, pca_rand=None):
    ops.set_bf16_mode(bool(model.cfg.train.use_amp))
    states = dict()
    with torch.no_grad():
        for epi_idx in range(D.episodes):
            observations, actions, rewards, nonterminals = get_episode_data(D, epi_idx=epi_idx, crop_idx=crop_idx, pca_rand=pca_rand)
            _observations = model._clip_obs(observations, idx_start=1)
            state = estimate_state(model, _observations, actions[:-1], rewards, nonterminals[:-1])
            states[D.file_names[epi_idx]] = tensor2numpy_state(state)
    return states


def run(cfg, cwd, device, model_class, model_path):
    from algos.MRSSM.MRSSM.train import get_dataset_loader
    D = get_dataset_loader(cfg, cwd, device, cfg.train.train_data_path)
    model = model_class(cfg, device)
    model.load_model(model_path)
    model.eval()
    print("model_path: {}".format(model_path))
    states = get_states(D, model, device)
    save_file_name = model_path.replace(".pth", ".npy").replace("/models_", "/states_models_")
    print("save to {}".format(save_file_name))
    np.save(save_file_name, states)
    return states
