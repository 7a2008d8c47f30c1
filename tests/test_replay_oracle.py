"""CPU: the replay buffer's host logic and oracle/replay_oracle.py against tests/golden/replay_*.pt (outputs of the
unmodified reference buffer, tests/golden/make_replay_golden.py).

* episode files -> stores: the product's loader fills a (host) store bit-identical to the reference's;
* chunk starts and augmentation choices: the product's `_sample_idx` / `_plan_batch` consume numpy's RNG exactly as the
  reference does (same slots, and the RNG ends in the same state);
* the batch math: the oracle fed with those choices and torch's RNG stream in the reference's order reproduces the
  reference batches bit for bit.
The device half (`sample` itself) is covered by tests/test_gpu_replay.py; on a host-only store it must fail loudly."""
import os

import numpy as np
import pytest
import torch

from oracle import replay_oracle as RO
from tests import replay_util as R


from tests.replay_util import load_fixture_buffer as _load  # noqa: E402


def oracle_batch(D, cfg, vec_idxs, plan, n, L):
    """The oracle on the buffer's stores with the planned choices; noise from torch's global RNG in the reference's order
    (per image modality: randn for the Gaussian noise when its scale is > 0, then rand for the dequantisation)."""
    obs = {}
    for name in D.observation_names:
        store = D.observations[name].cpu()
        if name not in plan:
            obs[name] = RO.gather_rows(store, vec_idxs, n, L)
            continue
        p = plan[name]
        C = store.shape[1]
        hw = (p["side"], p["side"]) if p["crop"] is not None else tuple(store.shape[-2:])
        shape = (L, n, C, *hw)
        gauss = torch.randn(*shape) if p["gauss_scale"] > 0 else None
        uniform = None if p["plain"] else torch.rand(shape)
        obs[name] = RO.gather_image(store, vec_idxs, n, L, crop=p["crop"], side=p["side"], delta=p["delta"], gauss=gauss,
                                    gauss_scale=p["gauss_scale"], uniform=uniform, bit_depth=D.bit_depth, normalise=not p["plain"])
    return (obs, RO.gather_rows(D.actions.cpu(), vec_idxs, n, L), RO.gather_rows(D.rewards.cpu(), vec_idxs, n, L),
            RO.gather_rows(D.nonterminals.cpu(), vec_idxs, n, L))


@pytest.mark.parametrize("name", ["default", "augment"])
def test_loader_fills_the_store_like_the_reference(name, golden_dir, tmp_path):
    rec, cfg, D = _load(name, golden_dir, tmp_path)
    assert (D.idx, D.full, D.steps, D.episodes) == (rec["idx"], rec["full"], rec["steps"], rec["episodes"])
    for k, d in rec["stores"].items():
        assert D.observations[k].dtype == (torch.uint8 if "image" in k else torch.float32)
        R.assert_digest(D.observations[k][:D.idx].float(), d)
    assert torch.equal(D.actions[:D.idx], rec["actions"])
    assert torch.equal(D.rewards[:D.idx], rec["rewards"])
    assert torch.equal(D.nonterminals[:D.idx], rec["nonterminals"])
    for k, (lam, vec) in rec["pca"].items():
        torch.testing.assert_close(D.lambd_eigen_values[k].cpu(), lam, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(D.p_eigen_vectors[k].cpu(), vec, rtol=1e-5, atol=1e-6)


def test_spiral_crop_table_matches_reference(golden_dir):
    from utils.replay_buffer.data_augment import spiral_offset
    rec = torch.load(os.path.join(golden_dir, "replay_augment.pt"), weights_only=False)
    assert [spiral_offset(i) for i in range(len(rec["spiral"]))] == [tuple(x) for x in rec["spiral"]]


@pytest.mark.parametrize("name", ["default", "augment"])
def test_host_plan_and_oracle_reproduce_reference_batches(name, golden_dir, tmp_path):
    rec, cfg, D = _load(name, golden_dir, tmp_path)
    for s in rec["samples"]:
        np.random.seed(s["seed"])
        torch.manual_seed(s["seed"])
        idxs = np.asarray([D._sample_idx(R.L) for _ in range(R.N)])
        assert np.array_equal(idxs, s["idxs"])
        np.random.seed(s["seed"])                                # the batched form `sample` uses draws the same chunks
        assert np.array_equal(D._sample_chunks(R.N, R.L), s["idxs"])
        vec_idxs, plan = D._plan_batch(idxs)
        obs, actions, rewards, nonterminals = oracle_batch(D, cfg, vec_idxs, plan, R.N, R.L)
        for k, d in s["obs"].items():
            R.assert_digest(obs[k], d)
        assert torch.equal(actions, s["actions"]) and torch.equal(rewards, s["rewards"])
        assert torch.equal(nonterminals, s["nonterminals"])
        assert float(np.random.rand()) == s["np_next"]          # both RNG streams were consumed exactly as the reference does
        assert float(torch.rand(())) == s["torch_next"]


def test_sample_idx_never_crosses_the_write_position():
    from utils.replay_buffer.memory import ExperienceReplay_Multimodal
    D = ExperienceReplay_Multimodal(size=20, observation_names=["v"], observation_shapes={"v": [2]}, action_size=1)
    np.random.seed(0)
    D.idx, D.full = 7, True                    # wrapped buffer: chunks may wrap around the end but not run over slot 7
    for _ in range(300):
        idxs = D._sample_idx(6)
        assert len(idxs) == 6 and D.idx not in idxs[1:]
        assert np.array_equal(idxs, np.arange(idxs[0], idxs[0] + 6) % 20)
    D.idx, D.full = 13, False                  # partly filled: only slots [0, idx) hold data
    for _ in range(300):
        idxs = D._sample_idx(5)
        assert idxs[0] >= 0 and idxs[-1] < 13


def test_append_and_episode_overflow():
    from utils.replay_buffer.memory import ExperienceReplay_Multimodal
    D = ExperienceReplay_Multimodal(size=3, observation_names=["image", "v"], observation_shapes={"image": [3, 64, 64], "v": [2]},
                                    action_size=2)
    frame = np.random.RandomState(0).rand(3, 64, 64).astype(np.float32) - 0.5
    for i in range(4):
        D.append({"image": frame, "v": np.array([i, -i], dtype=np.float32)}, np.array([1.0, 2.0]), 0.5 * i, i == 2)
    assert (D.idx, D.full, D.steps, D.episodes) == (1, True, 4, 1)
    expect = np.clip(np.floor((frame + 0.5) * 32) * 8, 0, 255).astype(np.uint8)     # image_processing.py:15-16
    assert np.array_equal(D.observations["image"][0].numpy(), expect)
    assert D.observations["v"][0].tolist() == [3.0, -3.0] and float(D.rewards[0]) == 1.5
    assert D.nonterminals[:, 0].tolist() == [1.0, 1.0, 0.0]


def test_sample_fails_loudly_without_a_device(golden_dir, tmp_path):
    """No CPU fallback: the gather kernels are the only implementation of `sample`."""
    if torch.cuda.is_available():
        pytest.skip("host-only check")
    _, _, D = _load("default", golden_dir, tmp_path)
    np.random.seed(0)
    with pytest.raises((RuntimeError, AssertionError)):
        D.sample(R.N, R.L)


def test_visualize_helpers():
    from utils.evaluation import visualize_utils as V
    rng = np.random.RandomState(0)
    frame = rng.rand(3, 8, 8).astype(np.float32) - 0.5
    img = V.reverse_image_observation(torch.as_tensor(frame))
    assert img.shape == (8, 8, 3) and img.dtype == np.uint8
    assert np.array_equal(img, np.clip(np.floor((frame + 0.5) * 32) * 8, 0, 255).astype(np.uint8).transpose(1, 2, 0))
    feat = rng.randn(5, 2, 6).astype(np.float32)
    x, y, z = V.get_xyz(feat)
    assert x.shape == (10,) and np.array_equal(z, feat.reshape(-1, 6)[:, 2])
    pca = V.get_pca_model(torch.as_tensor(feat), n_components=3)
    assert pca.transform(V.flat(feat)).shape == (10, 3)
    assert torch.equal(V.np2tensor(feat), torch.as_tensor(feat)) and isinstance(V.tensor2np(torch.ones(2)), np.ndarray)


def test_episode_preprocessing_edge_cases():
    """preprocess_data (reference memory.py:47-65): shortest stream wins, 'seed' is dropped, HWC -> CHW, normalised floats ->
    uint8, a plain 'image' key of another size is renamed, nonterminals = 1 - done; calc_image_shape keeps the crop margin."""
    from utils.replay_buffer.memory import calc_image_shape, clip_episode, preprocess_data
    rng = np.random.RandomState(1)
    hwc = rng.randint(0, 256, size=(5, 32, 32, 3)).astype(np.uint8)
    data = {"image": hwc.copy(), "v": rng.randn(7, 2).astype(np.float32), "done": np.array([0, 0, 0, 0, 1, 0], dtype=np.float32),
            "reward": np.zeros(5, dtype=np.float32), "seed": np.arange(9)}
    clipped, n = clip_episode(dict(data))
    assert n == 5 and "seed" not in clipped and all(len(v) == 5 for v in clipped.values())
    out, n = preprocess_data(dict(data))
    assert n == 5 and "image" not in out and out["image_32"].shape == (5, 3, 32, 32)
    assert np.array_equal(out["image_32"], hwc.transpose(0, 3, 1, 2))
    assert out["nonterminals"].shape == (5, 1) and out["nonterminals"][:, 0].tolist() == [1, 1, 1, 1, 0]
    as_float = (np.floor(hwc / 8.0) / 32.0 - 0.5).astype(np.float32)            # what normalize_image produces, noise-free
    out2, _ = preprocess_data({"image_horizon": as_float, "done": np.zeros(5, dtype=np.float32)})
    assert out2["image_horizon"].dtype == np.uint8
    assert np.array_equal(out2["image_horizon"], (hwc // 8 * 8).transpose(0, 3, 1, 2))
    assert calc_image_shape([3, 64, 64]) == [3, 64, 64]
    assert calc_image_shape([3, 64, 64], n_crop=9, dw_base=2, dh_base=3) == [3, 70, 68]      # k = 2: +k*dh rows, +k*dw columns


@pytest.mark.parametrize("name", ["default", "augment"])
def test_sample_plumbing_reproduces_reference_batches(name, golden_dir, tmp_path):
    """`sample` end to end on the host with the two kernel calls swapped for the oracle's gathers (test-only stand-ins; the
    product itself has no host path): chunk draws, slot order, per-modality plans, noise order and output shapes must give
    the reference batches bit for bit.  The kernels themselves are checked against the same fixtures in test_gpu_replay.py."""
    from utils.replay_buffer.data_augment import crop_size_of
    rec, cfg, D = _load(name, golden_dir, tmp_path)

    def gather_image(mod, slots, rows, crop, delta, gauss_scale, normalise):
        store = D.observations[mod]
        side = crop_size_of(mod)
        hw = (side, side) if crop is not None else tuple(store.shape[-2:])
        shape = (R.L, R.N, store.shape[1], *hw)
        gauss = torch.randn(*shape) if gauss_scale > 0 else None          # the reference's order: Gaussian noise, then dequantisation
        uniform = torch.rand(shape) if normalise else None
        out = RO.gather_image(store, slots.numpy(), R.N, R.L, crop=crop, side=side, delta=delta, gauss=gauss,
                              gauss_scale=gauss_scale, uniform=uniform, bit_depth=D.bit_depth, normalise=normalise)
        return out.reshape(rows, *out.shape[2:])

    D._gather_image = gather_image
    D._gather_rows = lambda store, slots, rows: store[slots].reshape(rows, -1)
    for s in rec["samples"]:
        np.random.seed(s["seed"])
        torch.manual_seed(s["seed"])
        obs, actions, rewards, nonterminals = D.sample(R.N, R.L)
        for k, d in s["obs"].items():
            R.assert_digest(obs[k], d)
        assert torch.equal(actions, s["actions"]) and torch.equal(rewards, s["rewards"])
        assert torch.equal(nonterminals, s["nonterminals"])
        assert float(np.random.rand()) == s["np_next"] and float(torch.rand(())) == s["torch_next"]
