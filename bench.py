#!/usr/bin/env python
"""MRSSM train-step benchmark (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric  : MRSSM train seq-steps/s = (world * B * T) / step time      (BASELINE.json)
workload: full ELBO train step (encoders + T-step observe rollout + decoders + KL, fwd + bwd,
          global-norm clip + Adam), MoPoE, image 64x64 + 3-d joint state, D=H=200, S=30,
          B=1024 sequences x T=50 per GPU (BASELINE config 3; weak scaling), synthetic
          COBOTTA-shaped data and seeded weights.
value   : K timed steps with the batch resident in HBM (CUDA events, barrier + sync both sides,
          max over ranks).
e2e     : the same steps through the public API model.optimize(D) with D handing out PINNED HOST
          buffers: H2D of the batch and D2H of the loss inside the timed region.
roofline: the dominant kernel of the step (by device time, measured live with CUDA events around
          every C-ABI call of one extra step) — algorithmic FLOPs or bytes / its mean duration,
          against MEASURED_PEAKS.json.  `rollout_roofline` is the same for the rollout kernel.
cpu_baseline / --impl reference: the UNMODIFIED reference's model.optimize(D) (oracle/_ref, staged by
          __graft_entry__.build(); kind "reference") — or, when it is not staged, the oracle port of it (kind "port") —
          timed on the host cores on a bounded sample (B=16 sequences of the same T).
same_box_rollout: BASELINE configs 2 and 4 (rollout only) of this build next to the reference modules on the same
          B200 under stock PyTorch.  strong_scaling (N > 1): the same global batch split over the ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multimodal-rssm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "mrssm_train_seq_steps_per_s"
UNIT = "seq-steps/s (B*T per second)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_rate(T, fusion, sample_B=16, steps=3, warmup=1, model=None):
    from oracle import mrssm_oracle as O
    torch.set_num_threads(os.cpu_count())
    m = model or MODELS[3]
    names = (m["image_name"], "pose_quat_v2")
    oc = O.OracleConfig(fusion=fusion, belief_size=m["belief"], hidden_size=m["hidden"], state_size=m["state"], names_enc=names, names_rec=names,
                        observation_shapes={m["image_name"]: [3, m["image"], m["image"]], "pose_quat_v2": [3]})
    P = O.make_params(oc, seed=0)
    opt = {}
    times = []
    for s in range(warmup + steps):
        batch, noise = O.synthetic_batch(oc, sample_B, T, seed=1234 + s)
        t0 = time.perf_counter()
        O.train_step(P, opt, oc, batch, noise)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return sample_B * T / med, med, torch.get_num_threads()


def reference_rate(T, fusion, sample_B=16, steps=3, warmup=1, in_process=False, model=None):
    """The UNMODIFIED reference (oracle/_ref, staged by __graft_entry__.build) on the host cores: model.optimize(D) at
    B = sample_B.  Runs in its own process unless in_process (the product mirrors the names `algos` / `utils`).
    -> (rate, seconds per step, cores, kind); falls back to the oracle port when the reference is not staged."""
    from oracle import ref_arm
    m = model or MODELS[3]
    if ref_arm.available():
        try:
            if in_process:
                r = ref_arm.time_train(sample_B, T, fusion, steps, warmup, "cpu", m["image"], m["belief"], m["state"], m["hidden"])
            else:
                out = subprocess.run([sys.executable, "-m", "oracle.ref_arm", "train", "--batch", str(sample_B), "--chunk", str(T),
                                      "--steps", str(steps), "--warmup", str(warmup), "--fusion", fusion, "--image", str(m["image"]),
                                      "--belief", str(m["belief"]), "--state", str(m["state"])], cwd=ROOT,
                                     capture_output=True, text=True, timeout=900)
                r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
            return r["seq_steps_per_s"], r["ms_per_step"] * 1e-3, r["cores"], "reference"
        except Exception as e:                                   # pragma: no cover
            print(f"bench: reference arm failed ({e!r}); timing the oracle port instead", file=sys.stderr)
    rate, med, cores = cpu_oracle_rate(T, fusion, sample_B=sample_B, steps=steps, warmup=warmup, model=m)
    return rate, med, cores, "port"


def same_box_reference():
    """the reference transition model on this B200 under stock PyTorch, configs 2 and 4 (own process)."""
    from oracle import ref_arm
    if not ref_arm.available():
        return None
    try:
        out = subprocess.run([sys.executable, "-m", "oracle.ref_arm", "samebox"], cwd=ROOT, capture_output=True, text=True, timeout=600)
        return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    except Exception as e:                                       # pragma: no cover
        print(f"bench: same-box reference arm failed ({e!r})", file=sys.stderr)
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sb = 16 if args.config == 3 else 4
    rate, med, cores, kind = reference_rate(args.chunk, args.fusion, sample_B=sb, steps=max(1, min(args.steps, 5 if args.config == 3 else 2)),
                                            warmup=max(1, min(args.warmup, 1)), in_process=True, model=model_of(args))
    what = "unmodified reference model.optimize(D)" if kind == "reference" else "fp32 oracle port of the reference step"
    sample = f"B={sb} of the B={args.batch} sequences, T={args.chunk}, full train step, {what}, fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# BASELINE.json configs: 3 = the headline (full ELBO train step, D = H = 200, S = 30, 64x64 image + 3-d vector, B = 1024 x T = 50
# per GPU); 5 = the scaled model (D = H = 1024, S = 64, 128x128 image; B = 2048 x T = 64 GLOBAL over 8 GPUs = 256 per GPU —
# 129 024 frames of 128x128 activations do not fit one GPU's 180 GB)
MODELS = {3: dict(belief=200, state=30, hidden=200, image=64, image_name="image_horizon", batch=1024, chunk=50),
          5: dict(belief=1024, state=64, hidden=1024, image=128, image_name="image_horizon_128", batch=256, chunk=64)}


def model_of(args):
    return MODELS[args.config]


def build_cfg(args, device, batch=None):
    from mrssm_b200.config import hot_path_config
    m = model_of(args)
    cfg = hot_path_config(fusion=args.fusion, batch_size=batch or args.batch, chunk_size=args.chunk, device=device,
                          belief_size=m["belief"], state_size=m["state"], hidden_size=m["hidden"], image_name=m["image_name"],
                          image_size=m["image"])
    cfg.train.use_amp = args.mode == "bf16"
    return cfg


def workload_config(args, world):
    m = model_of(args)
    return {"workload": f"mrssm_{args.fusion.lower()}_train_step_B{args.batch}_T{args.chunk}_per_gpu" + ("" if args.config == 3 else f"_config{args.config}"),
            "baseline_config": args.config, "fusion": args.fusion, "per_gpu_batch": args.batch, "chunk_size": args.chunk,
            "global_batch": args.batch * world, "belief": m["belief"], "state": m["state"], "hidden": m["hidden"],
            "modalities": f"{m['image_name']}[3,{m['image']},{m['image']}]+pose_quat_v2[3]", "mode": args.mode,
            "parallelism": f"dp{world}", "l2_policy": "inputs_exceed_l2"}


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class SyntheticReplay:
    """D.sample contract of the reference (utils/replay_buffer/memory.py:212-222): time-major fp32
    [obs dict, actions, rewards, nonterminals] on `device`, handing out tensors that already live in HBM
    (the `value` measurement; the e2e measurement uses mrssm_b200.data.PinnedChunkSource)."""

    def __init__(self, cfg, device, seed, n_buffers=2):
        B, T = cfg.train.batch_size, cfg.train.chunk_size
        g = torch.Generator(device=device).manual_seed(seed)
        self.device, self.i = device, 0
        self.dev = []
        for _ in range(n_buffers):
            obs = {}
            for name in cfg.rssm.observation_names_enc:
                shp = cfg.env.observation_shapes[name]
                if "image" in name:
                    u8 = torch.randint(0, 256, (T, B, *shp), generator=g, device=device)
                    x = torch.floor(u8 / 8) / 32 - 0.5 + torch.rand((T, B, *shp), generator=g, device=device) / 32
                    del u8
                else:
                    x = torch.randn((T, B, *shp), generator=g, device=device)
                obs[name] = x
            actions = torch.randn((T, B, cfg.env.action_size), generator=g, device=device)
            rewards = torch.zeros((T, B), device=device)
            nonterm = torch.ones((T, B, 1), device=device)
            drop = torch.rand(B, generator=g, device=device) < 0.1
            tpos = torch.randint(0, T, (B,), generator=g, device=device)
            nonterm[tpos[drop], torch.nonzero(drop).flatten(), 0] = 0
            self.dev.append((obs, actions, rewards, nonterm))

    def sample(self, n, L):
        self.i += 1
        obs, a, r, nt = self.dev[self.i % len(self.dev)]
        return [obs, a, r, nt]


def host_chunks(cfg, seed, n_chunks=2):
    """Pre-gathered replay chunks as the reference keeps them on the HOST (utils/replay_buffer/memory.py:160-168):
    uint8 frames, fp32 vectors / actions / rewards / nonterminals, time-major."""
    B, T = cfg.train.batch_size, cfg.train.chunk_size
    g = torch.Generator().manual_seed(seed)
    chunks = []
    for _ in range(n_chunks):
        obs = {}
        for name in cfg.rssm.observation_names_enc:
            shp = cfg.env.observation_shapes[name]
            if "image" in name:
                obs[name] = torch.randint(0, 256, (T, B, *shp), generator=g, dtype=torch.uint8)
            else:
                obs[name] = torch.randn((T, B, *shp), generator=g)
        actions = torch.randn((T, B, cfg.env.action_size), generator=g)
        rewards = torch.zeros((T, B))
        nonterm = torch.ones((T, B, 1))
        drop = torch.rand(B, generator=g) < 0.1
        tpos = torch.randint(0, T, (B,), generator=g)
        nonterm[tpos[drop], torch.nonzero(drop).flatten(), 0] = 0
        chunks.append((obs, actions, rewards, nonterm))
    return chunks


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/mrssm_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def timed_steps(model, D, steps, sync_loss):
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for _ in range(steps):
        model.optimize(D)
        if sync_loss:
            loss = float(model.model_loss)          # D2H read of the step's result
    e1.record()
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist.is_initialized():
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms, loss


def profile_one_step(model, D):
    """CUDA-event time of every C-ABI call of one step -> per-kernel aggregates."""
    from mrssm_b200 import _lib as L
    L.profile = []
    model.optimize(D)
    torch.cuda.synchronize()
    rec, L.profile = L.profile, None
    agg = {}
    for name, tag, work, a, b in rec:
        key = name.replace("mrssm_", "") + (f":{tag}" if tag else "")
        d = agg.setdefault(key, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
        d["ms"] += a.elapsed_time(b)
        d["n"] += 1
        if work:
            d["flops"] += work.get("flops", 0.0)
            d["bytes"] += work.get("bytes", 0.0)
    return agg


def measured_traffic(key):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/r02_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    e = t.get("kernels", {}).get(key)
    return None if e is None else e.get("dram_bytes_per_launch")


def roofline_of(key, d, pk, total_ms):
    ms = d["ms"] / d["n"]
    flops, byts = d["flops"] / d["n"], d["bytes"] / d["n"]
    ridge = pk["tf_sust"] * 1e12 / (pk["hbm"] * 1e9)
    tensor_bound = flops > 0 and (byts <= 0 or flops / byts > ridge)      # (multi-launch calls report flops only)
    if tensor_bound:
        ach = flops / (ms * 1e-3) / 1e12
        out = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"]}
    else:
        ach = byts / (ms * 1e-3) / 1e9
        out = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    out.update({"traffic": measured_traffic(key), "kernel": key, "launch_ms": ms, "launches_per_step": d["n"],
                "share_of_step": d["ms"] / total_ms, "peak_source": pk["src"],
                "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": byts})
    return out


def rollout_configs(model, device):
    """BASELINE configs 2 and 4 (rollout only) through MultimodalTransitionModel.__call__, CUDA events, next to the reference
    modules run on the same GPU under stock PyTorch (SURVEY §8d 'same box' comparator)."""
    from mrssm_b200 import ops
    tm = model.transition_model
    g = torch.Generator(device=device).manual_seed(5)
    rn = lambda *s: torch.randn(*s, device=device, generator=g)

    def timed(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    B, T, Bi, H = 256, 49, 4096, 100
    emb = {"image_horizon": rn(T, B, 1024), "pose_quat_v2": rn(T, B, 128)}
    a2, nt2 = rn(T, B, 3), torch.ones(T, B, 1, device=device)
    s0, h0 = torch.zeros(B, 30, device=device), torch.zeros(B, 200, device=device)
    a4, s4, h4 = rn(H, Bi, 3), rn(Bi, 30), rn(Bi, 200)
    ours = {}
    was = ops.bf16_mode()
    try:
        for mode in ("bf16", "fp32"):
            ops.set_bf16_mode(mode == "bf16")

            def fwd():
                with torch.no_grad():
                    tm(s0, a2, h0, emb, nt2)

            def fwd_bwd():
                e = {k: v.detach().requires_grad_(True) for k, v in emb.items()}
                out = tm(s0, a2, h0, e, nt2)
                torch.autograd.backward(list(out[:7]), [torch.ones_like(t) for t in out[:7]])

            def imagine():
                with torch.no_grad():
                    tm(s4, a4, h4)

            for key, fn, units in (("cfg2_observe_fwd", fwd, 256 * 50), ("cfg2_observe_fwd_bwd", fwd_bwd, 256 * 50),
                                   ("cfg4_imagine", imagine, Bi * H)):
                ms = timed(fn)
                ours.setdefault(key, {})[mode] = dict(ms=ms, seq_steps_per_s=units / (ms * 1e-3))
    finally:
        ops.set_bf16_mode(was)
        model.model_optimizer.zero_grad()
    return {"ours": ours, "reference_cuda_stock_pytorch": same_box_reference(),
            "what": "config 2: observe rollout B=256 T=50 (49 steps); config 4: imagination B=4096 H=100; fp32 = exact CUDA-core "
                    "kernels, bf16 = tcgen05 rollout; CUDA events, median of 5"}


def run_ours(args):
    from mrssm_b200 import _lib as L
    from mrssm_b200.config import hot_path_config
    from mrssm_b200.dist import DataParallel, init_from_env
    from algos.MRSSM.MRSSM.algo import build_RSSM

    rank, local, world = init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path (use --impl reference)"
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    cfg = build_cfg(args, device)
    torch.manual_seed(0)
    model = build_RSSM(cfg, torch.device(device))
    if world > 1:
        DataParallel(model)
    D = SyntheticReplay(cfg, device, seed=1234 + rank)

    for _ in range(args.warmup):
        model.optimize(D)
    sampler = ClockSampler(local)
    k0 = L.kernel_launches
    if rank == 0:
        sampler.start()
    ms, _ = timed_steps(model, D, args.steps, sync_loss=False)
    clocks = sampler.stop() if rank == 0 else None
    launches = (L.kernel_launches - k0)
    ms_step = ms / args.steps
    value = world * args.batch * args.chunk / (ms_step * 1e-3)

    # end to end through the public API: model.optimize(D) with D = the pinned-host chunk source.  Every step copies that
    # step's chunk host -> device (uint8 frames + fp32 vectors, as the reference's replay buffer stores them), normalises
    # the frames on the device, runs the step, and reads the loss back; the copy of step k+1 overlaps step k.
    from mrssm_b200.data import PinnedChunkSource
    Dh = PinnedChunkSource(host_chunks(cfg, seed=4321 + rank), device, bit_depth=5, seed=rank)
    model.optimize(Dh)
    ms_e2e, loss = timed_steps(model, Dh, args.steps, sync_loss=True)
    e2e_value = world * args.batch * args.chunk / (ms_e2e / args.steps * 1e-3)
    h2d_bytes = Dh.h2d_bytes
    del Dh
    strong = None
    if world > 1 and args.batch % world == 0:
        # secondary: strong scaling — the same GLOBAL batch (args.batch sequences) split over the ranks
        cfg_s = build_cfg(args, device, batch=args.batch // world)
        torch.manual_seed(0)
        model_s = build_RSSM(cfg_s, torch.device(device))
        DataParallel(model_s)
        Ds = SyntheticReplay(cfg_s, device, seed=99 + rank)
        for _ in range(3):
            model_s.optimize(Ds)
        ms_s, _ = timed_steps(model_s, Ds, args.steps, sync_loss=False)
        strong = {"global_batch": args.batch, "per_gpu_batch": args.batch // world, "ms_per_step": ms_s / args.steps,
                  "value": args.batch * args.chunk / (ms_s / args.steps * 1e-3), "unit": UNIT}
        del model_s, Ds
    if world > 1:
        # collectives are over: the per-kernel profile below runs on rank 0 alone (no gradient all-reduce)
        import torch.distributed as dist
        model.dp = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()

    if rank != 0:
        return
    pk = peaks()
    agg = profile_one_step(model, D)
    total_ms = sum(d["ms"] for d in agg.values())
    dom_key = max(agg, key=lambda k: agg[k]["ms"])
    roof = roofline_of(dom_key, agg[dom_key], pk, total_ms)
    rollout_roof = {k: roofline_of(k, agg[k], pk, total_ms) for k in ("rollout_fwd:observe", "rollout_bwd:observe", "rollout_tc_fwd:observe", "rollout_tc_bwd:observe") if k in agg}
    # The recurrence is serial: T-1 dependent steps, each a fixed chain of tcgen05.mma instructions issued by one thread.  The floor
    # below is that chain at the tensor pipe's measured issue rate and nothing else (profiles/micro/umma_rate_tight.txt: 56 cycles
    # per M=64, N=112 MMA; 263 MMAs per forward step, 175 per BPTT step for the 4-head model) - what the kernel would
    # take if epilogues, shared-memory traffic and the weight stream were free.
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    for k, n_mma in (("rollout_tc_fwd:observe", 263), ("rollout_tc_bwd:observe", 175)):
        if k in rollout_roof and model_of(args)["belief"] == 200 and args.fusion == "MoPoE":
            steps_t = args.chunk - 1
            floor = n_mma * 56.0 / (sm_hz * 1e6) * steps_t * 1e3
            rollout_roof[k]["serial"] = {"dependent_steps": steps_t, "us_per_step": rollout_roof[k]["launch_ms"] / steps_t * 1e3,
                                         "mma_per_step": n_mma, "mma_issue_floor_ms": floor,
                                         "frac_of_floor": floor / rollout_roof[k]["launch_ms"]}
    step_keys = [k for k in agg if k.startswith(("rstep_", "rollout_steps_", "tc_conv_down:step_", "tc_conv_up:step_")) or k == "add2"]
    if step_keys:       # large-model rollout (config 5): per-step launches, launch-latency-bound by construction
        ms_r = sum(agg[k]["ms"] for k in step_keys)
        n_r = 2 * (args.chunk - 1) * 7       # kernels of the two C-issued loops: 7 per time step and direction (one chunk of heads)
        rollout_roof["rollout_step(all launches)"] = {
            "bound": "latency", "ms_per_train_step": ms_r, "launches": n_r, "us_per_launch": ms_r / n_r * 1e3,
            "share_of_step": ms_r / total_ms, "dependent_steps": args.chunk - 1,
            "algorithmic_flops": sum(agg[k]["flops"] for k in step_keys),
            "achieved_tflops": sum(agg[k]["flops"] for k in step_keys) / (ms_r * 1e-3) / 1e12, "peak_tflops": pk["tf_sust"]}
    top = sorted(((k, d["ms"], d["n"]) for k, d in agg.items()), key=lambda x: -x[1])[:60]

    cpu = same_box = None
    if world == 1 and not args.no_cpu_baseline:
        sb, st = (16, 3) if args.config == 3 else (4, 1)
        rate, med, cores, kind = reference_rate(args.chunk, args.fusion, sample_B=sb, steps=st, warmup=1, model=model_of(args))
        what = "unmodified reference model.optimize(D)" if kind == "reference" else "fp32 oracle port"
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"B={sb} sequences x T={args.chunk}, full train step, {what}, median of {st} steps",
               "ms_per_step": med * 1e3}
        if args.config == 3:
            same_box = rollout_configs(model, device)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.mode == "fp32" else "bf16", "data": "synthetic",
        "config": workload_config(args, world),
        "model_steps_per_s": world * args.batch * (args.chunk - 1) / (ms_step * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": loss,
                "input": "uint8 frames + fp32 vectors from pinned host memory, normalised on device, copy of step k+1 overlapped with step k"},
        "gpu_launches": launches,
        "roofline": roof, "rollout_roofline": rollout_roof, "cpu_baseline": cpu, "same_box_rollout": same_box,
        "strong_scaling": strong, "kernel_order": list(agg.keys()),
        "top_kernels_ms": [{"kernel": k, "ms_per_step": round(m, 4), "launches": n} for k, m, n in top],
        "profiled_step_kernel_ms": total_ms,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(MODELS), help="BASELINE.json config: 3 (headline) or 5 (scaled model)")
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU (default: the config's)")
    ap.add_argument("--chunk", type=int, default=None)
    ap.add_argument("--fusion", default="MoPoE", choices=["MoPoE", "PoE", "NN", "single"])
    ap.add_argument("--mode", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.batch = args.batch or MODELS[args.config]["batch"]
    args.chunk = args.chunk or MODELS[args.config]["chunk"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
