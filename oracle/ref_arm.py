"""Reference arm: the UNMODIFIED reference code timed on the host cores (and, for the rollout-only configs, on the GPU
under stock PyTorch) — TEST / BENCH INFRASTRUCTURE, never imported by the product.

The reference is a source tree of scripts (no setup.py / pyproject, `pip install` has nothing to install), so
`__graft_entry__.build()` stages its hot-path packages (`algos/`, `utils/models/`) under the git-ignored `oracle/_ref/`
where the tree is mounted; the directory travels to the GPU box with the snapshot.  This module imports them from there —
in its OWN process, because the product mirrors the same top-level package names (`algos`, `utils`) — builds the model
through the reference's own factory `algos.MRSSM.MRSSM.algo.build_RSSM` (reference algos/MRSSM/MRSSM/algo.py:6-18) from an
attribute-dict carrying the YAML keys the path reads, and drives `model.optimize(D)` (base/algo.py:268-276).

    python -m oracle.ref_arm train   --batch 16 --chunk 50 --steps 3 --warmup 1      # CPU train step (BASELINE config 1)
    python -m oracle.ref_arm samebox                                                   # configs 2 and 4 on cuda:0, stock PyTorch
Each prints one JSON line.
"""
import argparse
import json
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "algos", "MRSSM")) and os.path.isdir(os.path.join(REF, "utils", "models"))


class _Attr(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = dict.__setitem__


def _attr(d):
    return _Attr({k: _attr(v) for k, v in d.items()}) if isinstance(d, dict) else d


def reference_cfg(B, T, device, fusion="MoPoE", belief=200, state=30, hidden=200, image=64):
    """The classic-default MRSSM of BASELINE.json (image + joint-state vector, normalization None), reference YAML key names
    (train/COBOTTA/SingleHoleDrilling/MRSSM/MRSSM/config/{main,env,rssm,train}/*.yaml)."""
    img = "image_horizon" if image == 64 else "image_horizon_128"
    names = [img, "pose_quat_v2"] if fusion != "single" else [img]
    return _attr(dict(
        main=dict(device=device, wandb=False),
        env=dict(observation_shapes={img: [3, image, image], "pose_quat_v2": [3]}, action_size=3),
        train=dict(batch_size=B, chunk_size=T, use_amp=False),
        rssm=dict(observation_names_enc=list(names), observation_names_rec=list(names), predict_reward=False,
                  multimodal=fusion != "single",
                  multimodal_params=dict(fusion_method=fusion if fusion != "single" else "MoPoE", expert_dist="q(st|ht,ot)"),
                  activation_function=dict(cnn="relu", dense="elu", fusion="relu"),
                  embedding_size=dict(fusion=1024, image=1024, sound=256, other=128),
                  hidden_size=hidden, belief_size=belief, state_size=state, normalization=None,
                  worldmodel_LogProbLoss=False, overshooting_distance=0, overshooting_kl_beta=0, overshooting_reward_scale=0,
                  global_kl_beta=1, free_nats=3, kl_beta=1, kl_balancing_alpha=0.5, learning_rate_schedule=0,
                  adam_epsilon=1e-7, grad_clip_norm=100.0, model_learning_rate=1e-3)))


def _import_reference():
    assert available(), "oracle/_ref is empty: run __graft_entry__.build() where /root/reference is mounted"
    for name in list(sys.modules):
        if name == "algos" or name.startswith("algos.") or name == "utils" or name.startswith("utils."):
            raise RuntimeError("oracle.ref_arm must run in its own process (the product's `algos` / `utils` are already imported)")
    sys.path.insert(0, REF)
    from algos.MRSSM.MRSSM.algo import build_RSSM
    return build_RSSM


def _synthetic(torch, B, T, image, device, seed):
    """COBOTTA-shaped batch, time-major fp32, as D.sample returns it (utils/replay_buffer/memory.py:212-222)."""
    g = torch.Generator().manual_seed(seed)
    img = "image_horizon" if image == 64 else "image_horizon_128"
    u8 = torch.randint(0, 256, (T, B, 3, image, image), generator=g)
    obs = {img: (torch.floor(u8 / 8) / 32 - 0.5 + torch.rand((T, B, 3, image, image), generator=g) / 32).to(device),
           "pose_quat_v2": torch.randn((T, B, 3), generator=g).to(device)}
    actions = torch.randn((T, B, 3), generator=g).to(device)
    nonterm = torch.ones((T, B, 1))
    drop = torch.rand(B, generator=g) < 0.1
    tpos = torch.randint(0, T, (B,), generator=g)
    for b in range(B):
        if drop[b]:
            nonterm[tpos[b], b, 0] = 0
    return obs, actions, torch.zeros(T, B, device=device), nonterm.to(device)


class _D:
    def __init__(self, batches):
        self.b, self.i = batches, 0

    def sample(self, n, L):
        self.i += 1
        obs, a, r, nt = self.b[self.i % len(self.b)]
        return [{k: v.clone() for k, v in obs.items()}, a.clone(), r.clone(), nt.clone()]


def time_train(B, T, fusion="MoPoE", steps=3, warmup=1, device="cpu", image=64, belief=200, state=30, hidden=200):
    """Median time of `model.optimize(D)` of the reference on `device` (all host threads when cpu) -> dict."""
    import torch
    build_RSSM = _import_reference()
    if device == "cpu":
        torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = build_RSSM(reference_cfg(B, T, device, fusion, belief, state, hidden, image), torch.device(device))
    D = _D([_synthetic(torch, B, T, image, device, 1234 + i) for i in range(2)])
    times = []
    for s in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.optimize(D)
        if device != "cpu":
            torch.cuda.synchronize()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return dict(kind="reference", seq_steps_per_s=B * T / med, ms_per_step=med * 1e3, cores=torch.get_num_threads(),
                B=B, T=T, device=device, steps=steps)


def same_box(device="cuda:0", reps=3):
    """the reference transition model on the B200 itself under stock PyTorch (library kernels), fp32:
    config 2 = observe rollout B=256, T=50 (forward, and forward + backward); config 4 = open-loop imagination B=4096, H=100."""
    import torch
    build_RSSM = _import_reference()
    torch.manual_seed(0)
    out = {}
    model = build_RSSM(reference_cfg(256, 50, device), torch.device(device))
    tm = model.transition_model
    g = torch.Generator(device=device).manual_seed(5)
    B, T = 256, 49

    def obs_inputs(req):
        emb = {"image_horizon": torch.randn(T, B, 1024, device=device, generator=g).requires_grad_(req),
               "pose_quat_v2": torch.randn(T, B, 128, device=device, generator=g).requires_grad_(req)}
        return (torch.zeros(B, 30, device=device), torch.randn(T, B, 3, device=device, generator=g),
                torch.zeros(B, 200, device=device), emb, torch.ones(T, B, 1, device=device))

    def timed(fn):
        ts = []
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts[1:]) * 1e3

    def fwd():
        with torch.no_grad():
            tm(*obs_inputs(False))

    def fwd_bwd():
        res = tm(*obs_inputs(True))
        sum(t.sum() for t in res[:7]).backward()

    ms = timed(fwd)
    out["cfg2_observe_fwd"] = dict(ms=ms, seq_steps_per_s=256 * 50 / (ms * 1e-3))
    ms = timed(fwd_bwd)
    out["cfg2_observe_fwd_bwd"] = dict(ms=ms, seq_steps_per_s=256 * 50 / (ms * 1e-3))
    Bi, H = 4096, 100

    def imagine():
        with torch.no_grad():
            tm(torch.randn(Bi, 30, device=device, generator=g), torch.randn(H, Bi, 3, device=device, generator=g),
               torch.randn(Bi, 200, device=device, generator=g), None, None)

    ms = timed(imagine)
    out["cfg4_imagine"] = dict(ms=ms, seq_steps_per_s=Bi * H / (ms * 1e-3))
    out["what"] = "reference MultimodalTransitionModel, stock PyTorch CUDA kernels, fp32, wall clock with synchronize, median of %d" % reps
    return out


def write_checkpoint(path, B=2, T=4, steps=2, fusion="MoPoE"):
    """Let the unmodified reference train `steps` steps on CPU and save its own checkpoint (base/algo.py:55-58: nested
    state dicts + torch.optim.Adam's state_dict) — the file the product's load_model must accept."""
    import torch
    build_RSSM = _import_reference()
    torch.manual_seed(0)
    model = build_RSSM(reference_cfg(B, T, "cpu", fusion), torch.device("cpu"))
    D = _D([_synthetic(torch, B, T, 64, "cpu", 7 + i) for i in range(2)])
    for _ in range(steps):
        model.optimize(D)
    os.makedirs(path, exist_ok=True)
    model.save_model(path, steps)
    return dict(file=os.path.join(path, "models_%d.pth" % steps), n_params=sum(p.numel() for p in model.param_list))


def try_expert_dist(expert_dist):
    """Does the unmodified reference construct a model with this `multimodal_params.expert_dist`?  (q(st|ot): its MoPoE / PoE
    factories call MultimodalStochasticEncoder(observation_names_enc=...) while that class takes observation_names_rec
    — MRSSM_MoPoE/algo.py:51-60 vs utils/models/encoder.py:885-895 — so construction raises TypeError.)"""
    import torch
    build_RSSM = _import_reference()
    cfg = reference_cfg(2, 4, "cpu", "MoPoE")
    cfg.rssm.multimodal_params.expert_dist = expert_dist
    try:
        build_RSSM(cfg, torch.device("cpu"))
        return dict(expert_dist=expert_dist, constructed=True)
    except Exception as e:                                   # noqa: BLE001  (the point is to report whatever the reference raises)
        return dict(expert_dist=expert_dist, constructed=False, error=type(e).__name__, message=str(e))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["train", "samebox", "ckpt", "expert_dist"])
    ap.add_argument("--expert-dist", default="q(st|ot)")
    ap.add_argument("--out", default="/tmp/mrssm_ref_ckpt")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--chunk", type=int, default=50)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--fusion", default="MoPoE")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--image", type=int, default=64)
    ap.add_argument("--belief", type=int, default=200)
    ap.add_argument("--state", type=int, default=30)
    a = ap.parse_args()
    if a.what == "expert_dist":
        res = try_expert_dist(a.expert_dist)
    elif a.what == "ckpt":
        res = write_checkpoint(a.out, fusion=a.fusion)
    elif a.what == "train":
        res = time_train(a.batch, a.chunk, a.fusion, a.steps, a.warmup, a.device, a.image, a.belief, a.state, a.belief)
    else:
        res = same_box(a.device if a.device != "cpu" else "cuda:0")
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
