"""CPU: the hydra-free composition of the YAML tree shipped next to main.py (mrssm_b200/config.py) and main.py's own
host logic.  The GPU half (main.py driving a few iterations on synthetic episode files) is tests/test_gpu_train_driver.py."""
import importlib.util
import os

import pytest

from mrssm_b200.config import apply_overrides, load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRY = os.path.join(ROOT, "multimodal-rssm_b200", "train", "COBOTTA", "SingleHoleDrilling", "MRSSM", "MRSSM")
BASELINE_OVERRIDES = ["rssm.observation_names_enc=[image_horizon,pose_quat_v2]", "rssm.observation_names_rec=[image_horizon,pose_quat_v2]",
                      "rssm.normalization=null", "rssm.hidden_size=200", "rssm.belief_size=200", "rssm.state_size=30", "main.wandb=False"]


def _main_module():
    spec = importlib.util.spec_from_file_location("mrssm_main", os.path.join(ENTRY, "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shipped_yaml_tree_composes_with_reference_keys_and_types():
    cfg = load_config(os.path.join(ENTRY, "config"))
    assert set(cfg) == {"main", "env", "rssm", "train"}
    r = cfg.rssm
    # shipped defaults of the reference (config/rssm/multimodal.yaml): image + sound, BatchNorm, 1024 / 128 latents
    assert r.observation_names_enc == ["image_horizon", "sound"] and r.normalization == "BatchNorm"
    assert (r.hidden_size, r.belief_size, r.state_size) == (1024, 1024, 128)
    assert r.multimodal_params.fusion_method == "MoPoE" and r.multimodal_params.expert_dist == "q(st|ht,ot)"
    assert isinstance(r.adam_epsilon, float) and r.adam_epsilon == 1e-7             # `1e-7` must not stay a string
    assert isinstance(r.model_learning_rate, float) and r.model_learning_rate == 1e-3
    assert r.kl_balancing_alpha == 0.5 and r.free_nats == 3 and r.grad_clip_norm == 100.0
    assert dict(r.embedding_size) == {"fusion": 1024, "image": 1024, "sound": 256, "other": 128}
    assert cfg.env.observation_shapes["image_horizon_128"] == [3, 128, 128] and cfg.env.action_size == 3 and cfg.env.bit_depth == 5
    t = cfg.train
    assert (t.batch_size, t.chunk_size, t.train_iteration, t.checkpoint_interval, t.validation_interval) == (50, 50, 10000, 1000, 10)
    assert t.experience_size == 500000 and t.use_amp is True and t.model_path is None
    assert t.augmentation.noise_scales == [0.0] and cfg.main.device == "cuda:0" and cfg.main.seed == 0


def test_overrides_follow_hydra_syntax():
    cfg = load_config(os.path.join(ENTRY, "config"), BASELINE_OVERRIDES + ["train.batch_size=7", "rssm.adam_epsilon=1e-5",
                                                                            "train.augmentation.pca_scales=[0.1,0.2]", "new.key=x"])
    assert cfg.rssm.normalization is None and cfg.rssm.observation_names_rec == ["image_horizon", "pose_quat_v2"]
    assert cfg.train.batch_size == 7 and cfg.rssm.adam_epsilon == 1e-5 and cfg.train.augmentation.pca_scales == [0.1, 0.2]
    assert cfg.main.wandb is False and cfg.new.key == "x"
    again = apply_overrides(cfg, ["rssm.multimodal_params.fusion_method=PoE"])
    assert again.rssm.multimodal_params.fusion_method == "PoE"


def test_main_prepares_the_experiment_like_the_reference(tmp_path):
    m = _main_module()
    cfg = load_config(os.path.join(ENTRY, "config"), ["rssm.overshooting_distance=80"])
    out = m.prepare(cfg, "RSSM", ["RSSM"], 0)
    assert out.main.experiment_name == "RSSM-seed_0" and out.main.tags == ["RSSM"] and out.main.seed == 0
    assert out.rssm.overshooting_distance == 50                      # clamped to chunk_size (utils/logger.py:42)
    assert cfg.main.experiment_name is None                          # the raw config is not mutated
    a = m.results_folder(str(tmp_path), "RSSM-seed_0")
    b = m.results_folder(str(tmp_path), "RSSM-seed_0")
    assert a.endswith("run_0") and b.endswith("run_1") and os.path.isdir(b)


def test_shipped_default_yaml_builds_the_model():
    """The unmodified YAML names BatchNorm + sound, deter = hidden = 1024, stoch = 128: the factory builds it,
    with the reference's state-dict keys for the normalisation layers."""
    from algos.MRSSM.MRSSM.algo import build_RSSM
    import torch
    cfg = load_config(os.path.join(ENTRY, "config"), ["main.wandb=False"])
    model = build_RSSM(cfg, torch.device("cpu"))
    sd = model.get_state_dict()
    enc, dec = sd["encoder"], sd["observation_model"]
    assert {"conv.0.weight", "conv.1.weight", "conv.1.bias", "conv.1.running_mean", "conv.1.running_var", "conv.1.num_batches_tracked",
            "conv.9.weight", "conv.10.running_var"} <= set(enc["image_horizon"]) and "conv.0.bias" not in enc["image_horizon"]
    assert {"down_sample_1.0.weight", "down_sample_2.1.running_mean", "down_sample_4.1.weight", "down_conversion.0.weight",
            "down_conversion.1.bias"} <= set(enc["sound"]) and "down_conversion.1.running_mean" not in enc["sound"]
    assert {"fc1.weight", "conv.0.weight", "conv.1.running_mean", "conv.7.bias", "conv.9.weight", "conv.9.bias"} <= set(dec["image_horizon"])
    assert {"up_conversion.weight", "up_sample_0.0.weight", "up_sample_2.1.running_var", "out.weight"} <= set(dec["sound"])
    assert tuple(dec["sound"]["up_conversion.weight"].shape) == (128 * 2 * 32 * 4, 1024 + 128, 1)


REF_CONFIG = "/root/reference/train/COBOTTA/SingleHoleDrilling/MRSSM/MRSSM/config"


@pytest.mark.skipif(not os.path.isdir(REF_CONFIG), reason="the reference tree is only mounted in the build container")
def test_yaml_tree_parses_to_the_reference_tree():
    """The five YAML files are written in their own layout; what they parse to is, key for key and value for value, what the
    reference's files parse to (so `python main.py` without overrides is the reference's shipped experiment)."""
    import yaml
    for rel in ("config.yaml", "main/main.yaml", "env/SingleHoleDrilling.yaml", "rssm/multimodal.yaml", "train/train.yaml"):
        with open(os.path.join(ENTRY, "config", rel)) as f:
            mine = yaml.safe_load(f)
        with open(os.path.join(REF_CONFIG, rel)) as f:
            ref = yaml.safe_load(f)
        assert mine == ref, rel
