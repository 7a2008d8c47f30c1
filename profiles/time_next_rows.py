"""Device timings of the two §8(f) rows added after the headline path (not part of bench.py's metric):
  1. replay sampling: `ExperienceReplay_Multimodal.sample(1024, 50)` on a device-resident store, CUDA events around the image
     gather kernel (mrssm_replay_gather_u8) -> achieved HBM GB/s on its algorithmic bytes (1 B read + 4 B written per pixel);
  2. one bf16 train step at B = 1024, T = 50 with latent overshooting on (distance 3, MoPoE: 4 open-loop runs of 49 152
     sequences) next to the same step with it off.
Run on the GPU box:  python profiles/time_next_rows.py > gpurun_out/next_rows.json"""
import json
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-rssm_b200"))
out = {}


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def replay():
    from mrssm_b200 import _lib as L
    from utils.replay_buffer.memory import ExperienceReplay_Multimodal
    dev = torch.device("cuda:0")
    size, n, Lc = 200_000, 1024, 50                     # 200k frames = 2.5 GB of uint8, far beyond the 126 MB L2
    D = ExperienceReplay_Multimodal(size=size, observation_names=["image_horizon", "pose_quat_v2"],
                                    observation_shapes={"image_horizon": [3, 64, 64], "pose_quat_v2": [3]}, n_crop=1, dh_base=1,
                                    dw_base=1, noise_scales=[0.0], pca_scales=None, action_name="a", action_size=3, device=dev)
    D.observations["image_horizon"].random_(0, 256)
    D.observations["pose_quat_v2"].normal_()
    D.actions.normal_(); D.rewards.normal_(); D.nonterminals.fill_(1.0)
    D.idx, D.full = 0, True
    np.random.seed(0)
    ms_sample = timed(lambda: D.sample(n, Lc), reps=5)                       # host index draws + all gathers
    idxs = np.asarray([D._sample_idx(Lc) for _ in range(n)])
    vec, plan = D._plan_batch(idxs)
    slots = torch.from_numpy(vec).to(dev)
    ms_kernel = timed(lambda: D._gather_image("image_horizon", slots, n * Lc, (0, 0), None, 0.0, True), reps=10)
    px = n * Lc * 3 * 64 * 64
    out["replay"] = dict(sample_ms=ms_sample, gather_u8_ms=ms_kernel, algorithmic_bytes=5 * px,
                         achieved_GBps=5 * px / ms_kernel / 1e6, frames=n * Lc, store_frames=size)


def overshooting():
    import bench
    from mrssm_b200.config import hot_path_config
    from algos.MRSSM.MRSSM.algo import build_RSSM
    dev = torch.device("cuda:0")
    res = {}
    for label, kw in (("off", {}), ("distance3", dict(overshooting_distance=3, overshooting_kl_beta=1.0))):
        cfg = hot_path_config(fusion="MoPoE", batch_size=1024, chunk_size=50, device="cuda:0", **kw)
        cfg.train.use_amp = True
        torch.manual_seed(0)
        model = build_RSSM(cfg, dev)
        D = bench.SyntheticReplay(cfg, "cuda:0", seed=0)
        res[label] = timed(lambda: model.optimize(D), reps=3, warm=3)
        res[label + "_loss"] = float(model.model_loss)
        del model, D
        torch.cuda.empty_cache()
    out["overshooting_step_ms"] = res


parts = [p for p in (replay, overshooting) if len(sys.argv) < 2 or p.__name__ in sys.argv[1:]]
for part in parts:
    try:
        part()
    except Exception:
        out[part.__name__ + "_error"] = traceback.format_exc()
print(json.dumps(out, indent=1))
