"""-m gpu: the callers either side of the hot path, end to end on the device —
`algos/MRSSM/MRSSM/train.py` (episode directories -> device-resident replay buffers -> optimize / validation / checkpoint,
reference train.py:9-58) and `utils/evaluation/estimate_states.py` (B = 1, T = episode length `estimate_state` over every
stored episode, reference estimate_states.py:60-91), plus B = 1 parity of that path against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import mrssm_oracle as O
from tests import parity_util as U
from tests import replay_util as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cfg(use_amp):
    from mrssm_b200.config import hot_path_config, to_attr
    cfg = hot_path_config(fusion="MoPoE", batch_size=3, chunk_size=4, device=DEV)
    cfg.train.use_amp = use_amp
    cfg.train.experience_size = R.SIZE
    cfg.train.augmentation = to_attr(dict(n_crop=1, dh_base=1, dw_base=1, noise_scales=[0.0], pca_scales=[0.0]))
    cfg.train.train_data_path, cfg.train.validation_data_path = "train", "val"
    cfg.train.train_iteration, cfg.train.validation_interval, cfg.train.checkpoint_interval = 4, 2, 4
    cfg.train.model_path = None
    return cfg


@pytest.mark.parametrize("use_amp", [False, True])
def test_train_driver_then_estimate_states(tmp_path, use_amp):
    from algos.MRSSM.MRSSM.algo import build_RSSM
    from algos.MRSSM.MRSSM.train import train
    from utils.evaluation import estimate_states as ES
    cfg = _cfg(use_amp)
    R.write_dataset(str(tmp_path / "train"), R.CONFIGS["default"], seed=5)
    R.write_dataset(str(tmp_path / "val"), R.CONFIGS["default"], seed=6)
    results = tmp_path / "results"
    os.makedirs(results)
    np.random.seed(0)
    model = train(cfg, str(tmp_path), str(results), torch.device(DEV))
    ckpt = str(results / "models_4.pth")
    assert os.path.exists(ckpt)
    assert model.itr_optim == 4 and torch.isfinite(model.model_loss).item()
    assert all(np.isfinite(float(v)) for v in model.validation_info.values())

    states = ES.run(cfg, str(tmp_path), torch.device(DEV), build_RSSM, ckpt)
    assert os.path.exists(str(results / "states_models_4.npy"))
    assert len(states) == len(R.EPISODES)
    lengths = sorted(st["beliefs"].shape[0] + 1 for st in states.values())
    assert lengths == sorted(R.EPISODES)
    for name, st in states.items():
        T1 = st["beliefs"].shape[0]
        assert name.endswith(".npy") and st["beliefs"].shape == (T1, 1, 200) and st["posterior_means"].shape == (T1, 1, 30)
        assert all(np.isfinite(v).all() for k, v in st.items() if isinstance(v, np.ndarray))
        assert set(st["expert_means"].keys()) == {"prior_expert", R.IMAGE, R.VEC}


def test_episode_data_is_the_stored_episode(tmp_path):
    from algos.MRSSM.MRSSM.train import get_dataset_loader
    from utils.evaluation import estimate_states as ES
    cfg = _cfg(False)
    R.write_dataset(str(tmp_path / "train"), R.CONFIGS["default"], seed=5)
    D = get_dataset_loader(cfg, str(tmp_path), torch.device(DEV), "train")
    bounds = ES.episode_bounds(D)
    assert bounds[0] == 0 and bounds[-1] == D.idx == sum(R.EPISODES) and len(bounds) == len(R.EPISODES) + 1
    for e in range(len(R.EPISODES)):
        obs, actions, rewards, nonterminals = ES.get_episode_data(D, e, crop_idx=0)
        s, t = int(bounds[e]), int(bounds[e + 1])
        T = t - s
        assert obs[R.IMAGE].shape == (T, 1, 3, 64, 64) and actions.shape == (T, 1, 3)
        assert rewards.shape == (T, 1) and nonterminals.shape == (T, 1, 1)
        # 5-bit quantisation level of every stored pixel + dequantisation noise in [0, 1] levels (1 only when fp32 rounds
        # level + u, u -> 1, up — the reference's normalize_image has the same property)
        noise = (obs[R.IMAGE][:, 0] + 0.5) * 32 - torch.floor(D.observations[R.IMAGE][s:t].float() / 8)
        assert float(noise.min()) >= 0.0 and float(noise.max()) <= 1.0 and 0.45 < float(noise.mean()) < 0.55
        assert torch.equal(obs[R.VEC][:, 0], D.observations[R.VEC][s:t])
        assert torch.equal(actions[:, 0], D.actions[s:t]) and torch.equal(rewards[:, 0], D.rewards[s:t])
        assert float(nonterminals[-1]) == 0.0 and float(nonterminals[:-1].min()) == 1.0


@pytest.mark.parametrize("bf16", [False, True])
def test_single_sequence_estimate_state_matches_oracle(tmp_path, bf16):
    """B = 1: one stored episode through encoders + observe rollout against the oracle with shared noise.  fp32 mode: rtol
    1e-3; bf16 mode: the stated bf16 tolerance of test_gpu_rollout_configs.py (mean abs error <= 1e-2, max <= 0.25)."""
    from algos.MRSSM.MRSSM.train import get_dataset_loader
    from mrssm_b200 import ops
    from mrssm_b200.noise import FixedNoise
    from utils.evaluation import estimate_states as ES
    R.write_dataset(str(tmp_path / "train"), R.CONFIGS["default"], seed=5)
    D = get_dataset_loader(_cfg(False), str(tmp_path), torch.device(DEV), "train")
    oc = U.oracle_cfg("MoPoE")
    obs, actions, rewards, nonterminals = ES.get_episode_data(D, 0, crop_idx=0)
    T = actions.shape[0]
    model, P = U.build_product(oc, 1, T, DEV, bf16=bf16)
    tgt = model._clip_obs(obs, idx_start=1)
    g = torch.Generator().manual_seed(4)
    eps_prior, eps_post = torch.randn(T - 1, 1, oc.state_size, generator=g), torch.randn(T - 1, 1, oc.state_size, generator=g)
    with torch.no_grad():
        ref = O.estimate_state(P, oc, {k: v.cpu() for k, v in tgt.items()}, actions[:-1].cpu(), nonterminals[:-1].cpu(),
                               eps_prior, eps_post, det=False)
        ops.set_bf16_mode(bf16)
        try:
            with FixedNoise(prior=eps_prior.to(DEV), post=eps_post.to(DEV)):
                st = model.estimate_state(tgt, actions[:-1], rewards, nonterminals[:-1])
        finally:
            ops.set_bf16_mode(False)
    if not bf16:
        U.assert_states_close(st, ref, 1e-3, 2e-5)
        return
    for k, v in ref.items():
        items = v.items() if isinstance(v, dict) else [(None, v)]
        for n, t in items:
            if t is None:
                continue
            mine = (st[k][n] if n is not None else st[k]).cpu()
            err = (mine - t).abs()
            assert torch.isfinite(mine).all() and float(err.mean()) <= 1e-2 and float(err.max()) <= 0.25, (k, n, float(err.max()))


def test_main_entry_point_trains_from_the_yaml_tree(tmp_path):
    """main.py (reference train/COBOTTA/SingleHoleDrilling/MRSSM/MRSSM/main.py:37-49): shipped YAML + `group.key=value`
    overrides -> run(): a few iterations on synthetic episode files, a checkpoint in results/<experiment>/<date>/run_0 that
    loads back into a fresh model (optimizer state included)."""
    import importlib.util
    import glob
    from algos.MRSSM.MRSSM.algo import build_RSSM
    from tests.test_config_entry import ENTRY
    spec = importlib.util.spec_from_file_location("mrssm_main", os.path.join(ENTRY, "main.py"))
    main = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(main)
    R.write_dataset(str(tmp_path / "train"), R.CONFIGS["default"], seed=5)
    R.write_dataset(str(tmp_path / "val"), R.CONFIGS["default"], seed=6)
    names = f"[{R.IMAGE},{R.VEC}]"
    ov = [f"rssm.observation_names_enc={names}", f"rssm.observation_names_rec={names}", "rssm.normalization=null",
          "rssm.hidden_size=200", "rssm.belief_size=200", "rssm.state_size=30", "main.wandb=False", f"main.device={DEV}",
          f"env.action_name={R.ACTION}", "env.action_size=3",
          "train.train_data_path=train", "train.validation_data_path=val", f"train.experience_size={R.SIZE}",
          "train.batch_size=3", "train.chunk_size=4", "train.train_iteration=4", "train.validation_interval=2",
          "train.checkpoint_interval=4"]
    (model,) = main.main(["--cwd", str(tmp_path)] + ov)
    assert model.itr_optim == 4 and torch.isfinite(model.model_loss).item()
    assert model.cfg.main.experiment_name == "RSSM-seed_0" and model.cfg.train.use_amp is True      # shipped default: tensor-core mode
    runs = glob.glob(str(tmp_path / "results" / "RSSM-seed_0" / "*" / "run_0"))
    assert len(runs) == 1 and os.path.exists(os.path.join(runs[0], "hydra_config.yaml"))
    ckpt = os.path.join(runs[0], "models_4.pth")
    assert os.path.exists(ckpt)
    fresh = build_RSSM(model.cfg, torch.device(DEV))
    fresh.load_model(ckpt)
    for a, b in zip(fresh.param_list, model.param_list):
        assert torch.equal(a, b)


def test_main_entry_point_runs_the_unmodified_shipped_yaml(tmp_path):
    """The shipped YAML as it is — image + sound, BatchNorm, MoPoE, deter = hidden = 1024, stoch = 128, use_amp
    — with only the data paths, sizes and iteration counts overridden: trains, validates (eval-mode normalisation), checkpoints,
    and the checkpoint (running statistics included) loads back."""
    import importlib.util
    import glob
    from algos.MRSSM.MRSSM.algo import build_RSSM
    from tests.test_config_entry import ENTRY
    spec = importlib.util.spec_from_file_location("mrssm_main", os.path.join(ENTRY, "main.py"))
    main = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(main)
    R.write_dataset(str(tmp_path / "train"), R.CONFIGS["shipped"], seed=5)
    R.write_dataset(str(tmp_path / "val"), R.CONFIGS["shipped"], seed=6)
    ov = ["main.wandb=False", f"main.device={DEV}", "train.train_data_path=train", "train.validation_data_path=val",
          f"train.experience_size={R.SIZE}", "train.batch_size=3", "train.chunk_size=4", "train.train_iteration=3",
          "train.validation_interval=2", "train.checkpoint_interval=3"]
    (model,) = main.main(["--cwd", str(tmp_path)] + ov)
    c = model.cfg.rssm
    assert list(c.observation_names_enc) == ["image_horizon", "sound"] and c.normalization == "BatchNorm"
    assert (c.belief_size, c.hidden_size, c.state_size) == (1024, 1024, 128) and c.multimodal_params.fusion_method == "MoPoE"
    assert model.itr_optim == 3 and torch.isfinite(model.model_loss).item()
    bn = model.encoder.encoders["image_horizon"].conv[1]
    assert int(bn.num_batches_tracked) == 3 and float(bn.running_var.sub(1).abs().max()) > 0
    inorm = model.encoder.encoders["sound"].down_sample_2[1]
    assert float(inorm.running_mean.abs().max()) > 0
    runs = glob.glob(str(tmp_path / "results" / "*" / "*" / "run_0"))
    ckpt = os.path.join(runs[0], "models_3.pth")
    assert os.path.exists(ckpt)
    fresh = build_RSSM(model.cfg, torch.device(DEV))
    fresh.load_model(ckpt)
    for a, b in zip(fresh.param_list, model.param_list):
        assert torch.equal(a, b)
    assert torch.equal(fresh.encoder.encoders["image_horizon"].conv[1].running_mean, bn.running_mean)
    assert torch.equal(fresh.observation_model.observation_models["sound"].up_sample_1[1].running_var,
                       model.observation_model.observation_models["sound"].up_sample_1[1].running_var)
