"""Train-step time of the UNMODIFIED shipped YAML (image + sound, BatchNorm, MoPoE, deter = hidden = 1024, stoch = 128, B = 50 x T = 50,
bf16 mode) on one B200, with the per-kernel profile of one step.  python profiles/time_shipped_yaml.py [B] [T]"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-rssm_b200"))
import bench
from mrssm_b200 import _lib as L
from mrssm_b200.config import load_config
from algos.MRSSM.MRSSM.algo import build_RSSM
ENTRY = os.path.join(ROOT, "multimodal-rssm_b200", "train", "COBOTTA", "SingleHoleDrilling", "MRSSM", "MRSSM")
ov = ["main.wandb=False", "main.device=cuda:0"]
if len(sys.argv) > 1: ov.append("train.batch_size=%d" % int(sys.argv[1]))
if len(sys.argv) > 2: ov.append("train.chunk_size=%d" % int(sys.argv[2]))
cfg = load_config(os.path.join(ENTRY, "config"), ov)
torch.manual_seed(0)
model = build_RSSM(cfg, torch.device("cuda:0"))
D = bench.SyntheticReplay(cfg, "cuda:0", seed=1)
for _ in range(2): model.optimize(D)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): model.optimize(D)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
agg = bench.profile_one_step(model, D)
top = sorted(((k, d["ms"], d["n"], d["flops"]) for k, d in agg.items()), key=lambda x: -x[1])[:25]
B, T = cfg.train.batch_size, cfg.train.chunk_size
print(json.dumps(dict(workload="shipped YAML, B=%d T=%d, bf16 mode" % (B, T), ms_per_step=ms, seq_steps_per_s=B * T / (ms * 1e-3),
                      loss=float(model.model_loss),
                      top=[dict(kernel=k, ms=round(m, 3), launches=n, tflops=round(f / (m * 1e-3) / 1e12, 2) if m > 0 else 0) for k, m, n, f in top])))
