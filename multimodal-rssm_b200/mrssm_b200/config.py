"""Configuration without hydra/omegaconf (neither is installable here): an attribute-dict over the same
YAML tree the reference composes with hydra (config/config.yaml -> main/ env/ rssm/ train/ groups) plus
``group.key=value`` command-line overrides.  Only attribute access, ``dict(cfg.x)`` and list values are
needed by the hot path (SURVEY §8b 'Config keys actually read')."""
import copy
import os
import re

import yaml

# PyYAML (YAML 1.1) reads `1e-7` as a string; omegaconf — which the reference's hydra entry uses — reads it as a float
_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)[eE][+-]?\d+$")


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, list):
        return [to_attr(v) for v in d]
    if isinstance(d, str) and _FLOAT.match(d):
        return float(d)
    return d


def apply_overrides(cfg, overrides):
    for ov in overrides:
        key, _, raw = ov.partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, AttrDict())
        node[parts[-1]] = to_attr(yaml.safe_load(raw))
    return cfg


def load_config(config_dir, overrides=()):
    """Compose <config_dir>/config.yaml's ``defaults`` list (``- group: file``) like hydra does."""
    with open(os.path.join(config_dir, "config.yaml")) as f:
        root = yaml.safe_load(f)
    cfg = AttrDict()
    for entry in root.get("defaults", []):
        for group, fname in entry.items():
            with open(os.path.join(config_dir, group, f"{fname}.yaml")) as f:
                cfg[group] = to_attr(yaml.safe_load(f) or {})
    return apply_overrides(cfg, overrides)


def hot_path_config(fusion="MoPoE", batch_size=50, chunk_size=50, device="cuda:0", belief_size=200,
                    state_size=30, hidden_size=200, image_name="image_horizon", image_size=64,
                    vector_name="pose_quat_v2", **rssm_overrides):
    """The classic-default MRSSM of BASELINE.json (image + joint-state vector, normalization None).
    fusion: MoPoE | PoE | NN | single."""
    multimodal = fusion != "single"
    names = [image_name, vector_name] if multimodal else [image_name]
    rssm = dict(
        observation_names_enc=list(names), observation_names_rec=list(names), condition_names=["d_pose_quat_v2"],
        predict_reward=False, multimodal=multimodal,
        multimodal_params=dict(fusion_method=fusion if multimodal else "MoPoE", expert_dist="q(st|ht,ot)"),
        activation_function=dict(cnn="relu", dense="elu", fusion="relu"),
        embedding_size=dict(fusion=1024, image=1024, sound=256, other=128),
        hidden_size=hidden_size, belief_size=belief_size, state_size=state_size, normalization=None,
        worldmodel_LogProbLoss=False, overshooting_distance=0, overshooting_kl_beta=0, overshooting_reward_scale=0,
        global_kl_beta=1, free_nats=3, kl_beta=1, kl_balancing_alpha=0.5, learning_rate_schedule=0,
        adam_epsilon=1e-7, grad_clip_norm=100.0, model_learning_rate=1e-3)
    rssm.update(rssm_overrides)
    return to_attr(dict(
        main=dict(experiment_name="bench", tags=None, log_dir=None, seed=0, disable_cuda=False, device=device,
                  wandb=False, git_hash=None),
        env=dict(observation_shapes={image_name: [3, image_size, image_size], vector_name: [3]},
                 action_name="d_pose_quat_v2", action_size=3, action_repeat=1, bit_depth=5),
        rssm=rssm,
        train=dict(batch_size=batch_size, chunk_size=chunk_size, use_amp=False, train_iteration=10000,
                   checkpoint_interval=1000, validation_interval=10, model_path=None)))
