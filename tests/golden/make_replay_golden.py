"""Generates tests/golden/replay_*.pt by running the UNMODIFIED reference replay buffer
(/root/reference/utils/replay_buffer/memory.py) on the synthetic episode files of tests/replay_util.py, CPU only.

    python tests/golden/make_replay_golden.py

Recorded per configuration: the buffer after load_dataset (stores, counters, PCA parameters), the spiral crop table, and
two seeded sample(n, L) calls (slot numbers, batch digests).  numpy's and torch's global RNGs are seeded before each call;
the reference draws chunk starts / augmentation choices from numpy and the Gaussian / dequantisation noise from torch, in a
fixed order the tests reproduce."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import replay_util as R  # noqa: E402

if "omegaconf" not in sys.modules:                      # only used for an isinstance check on dataset path lists
    stub = types.ModuleType("omegaconf")
    stub.ListConfig = list
    sys.modules["omegaconf"] = stub
sys.path.insert(0, "/root/reference")
from utils.replay_buffer import memory as ref_memory  # noqa: E402
from utils.replay_buffer import data_augment as ref_aug  # noqa: E402


def main():
    for cname, cfg in R.CONFIGS.items():
        with tempfile.TemporaryDirectory() as root:
            files = R.write_dataset(root, cfg)
            D = ref_memory.ExperienceReplay_Multimodal(**R.buffer_kwargs(cfg, torch.device("cpu")))
            D.file_names += files
            for f in files:                             # load_dataset's loop, in a fixed file order (glob order is not)
                D._set_data_to_buffer(f)
            if D.pca_scales is not None:
                D._set_color_aug_params()
            rec = dict(config=cname, files=[os.path.basename(f) for f in files], idx=D.idx, full=D.full, steps=D.steps,
                       episodes=D.episodes,
                       stores={k: R.digest(v[:D.idx].float()) for k, v in D.observations.items()},
                       actions=D.actions[:D.idx].clone(), rewards=D.rewards[:D.idx].clone(),
                       nonterminals=D.nonterminals[:D.idx].clone(),
                       pca={k: (D.lambd_eigen_values[k].clone(), D.p_eigen_vectors[k].clone())
                            for k in D.observation_names if D.lambd_eigen_values.get(k) is not None},
                       spiral=[(ref_aug.get_dx(i), ref_aug.get_dy(i)) for i in range(60)], samples=[])
            for seed in (11, 12):
                np.random.seed(seed)
                torch.manual_seed(seed)
                picked = []
                orig = D._retrieve_batch

                def spy(idxs, n, L, _orig=orig, _picked=picked):
                    _picked.append(np.array(idxs))
                    return _orig(idxs, n, L)
                D._retrieve_batch = spy
                obs, actions, rewards, nonterminals = D.sample(R.N, R.L)
                D._retrieve_batch = orig
                rec["samples"].append(dict(seed=seed, idxs=picked[0], obs={k: R.digest(v) for k, v in obs.items()},
                                           actions=actions.clone(), rewards=rewards.clone(), nonterminals=nonterminals.clone(),
                                           np_next=float(np.random.rand()), torch_next=float(torch.rand(()))))
            torch.save(rec, os.path.join(HERE, "replay_%s.pt" % cname))
            print(cname, "idx", D.idx, "steps", D.steps, {k: v["sum"] for k, v in rec["samples"][0]["obs"].items()})


if __name__ == "__main__":
    main()
