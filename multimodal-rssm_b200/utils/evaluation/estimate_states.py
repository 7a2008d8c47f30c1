"""Latent states of every stored episode: the B = 1, T = episode-length consumer of `estimate_state`
(interface of the reference's utils/evaluation/estimate_states.py).

An episode is a run of consecutive slots of the device-resident store, so its observations come out of the same fused
gather kernel as training batches (n = 1, L = T, fixed crop position, optional fixed colour shift); nothing is staged
through the host."""
import numpy as np
import torch

from mrssm_b200 import ops


def _to_numpy(value):
    if isinstance(value, dict):
        return {k: _to_numpy(v) for k, v in value.items()}
    return value.detach().cpu().numpy() if torch.is_tensor(value) else value


def tensor2numpy_state(state):
    """Every tensor of a state dict (per-expert dicts included) as a host array, in place."""
    for key in list(state.keys()):
        state[key] = _to_numpy(state[key])
    return state


def estimate_state(model, observations, actions, rewards, nonterminals):
    return model.estimate_state(observations, actions, rewards, nonterminals)


def get_all_data(D):
    """Views of the filled part of the stores (they stay on the device)."""
    filled = slice(0, D.idx)
    return ({name: D.observations[name][filled] for name in D.observation_names}, D.actions[filled], D.rewards[filled],
            D.nonterminals[filled])


def episode_bounds(D):
    """Slot ranges [b[e], b[e+1]) of the stored episodes: an episode ends where nonterminals is 0 (reference :36-39)."""
    ends = torch.nonzero(D.nonterminals[:D.idx, 0] == 0).flatten().cpu().numpy() + 1
    return np.concatenate([[0], ends])


def get_episode_data(D, epi_idx, crop_idx=None, pca_rand=None):
    """One episode as a batch of one sequence: observations {name: [T,1,...]}, actions [T,1,A], rewards [T,1],
    nonterminals [T,1,1], augmented and normalised like a training batch (reference :35-58)."""
    bounds = episode_bounds(D)
    first, stop = int(bounds[epi_idx]), int(bounds[epi_idx + 1])
    slots = np.arange(first, stop).reshape(1, -1)
    return D._retrieve_batch(slots, 1, stop - first, crop_idx=crop_idx, pca_rand=pca_rand)


def get_states(D, model, device, crop_idx=0, pca_rand=None):
    """{episode file name: state dict of host arrays} for every stored episode (reference :60-71)."""
    ops.set_bf16_mode(bool(model.cfg.train.use_amp))
    states = {}
    with torch.no_grad():
        for e, file_name in enumerate(D.file_names[:D.episodes]):
            obs, actions, rewards, nonterminals = get_episode_data(D, e, crop_idx=crop_idx, pca_rand=pca_rand)
            state = estimate_state(model, model._clip_obs(obs, idx_start=1), actions[:-1], rewards, nonterminals[:-1])
            states[file_name] = tensor2numpy_state(state)
    return states


def run(cfg, cwd, device, model_class, model_path):
    from algos.MRSSM.MRSSM.train import get_dataset_loader
    D = get_dataset_loader(cfg, cwd, device, cfg.train.train_data_path)
    model = model_class(cfg, device)
    model.load_model(model_path)
    model.eval()
    states = get_states(D, model, device)
    out = model_path.replace(".pth", ".npy").replace("/models_", "/states_models_")
    np.save(out, states)
    print("states of {} -> {}".format(model_path, out))
    return states
