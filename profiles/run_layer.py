"""Run one production plane layer a few times (target of `ncu -k regex:plane_fwd`): python profiles/run_layer.py E2_fwd [n] [BI TH NA bres NB]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch
from sweep_plans import L, build, timed

name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50176
if len(sys.argv) > 7:
    L.call_host("mrssm_pl_set_plan_override", *[int(v) for v in sys.argv[3:8]])
f = build(name, n)
print(name, n, "%.3f ms" % timed(f, reps=3))

# per-tile clock64 timeline of CTA 0 (slots: see PROF() in csrc/conv_plane.cu)
prof = torch.zeros(148 * 16 * 8, device="cuda:0", dtype=torch.int64)
L.call_host("mrssm_pl_set_profile_buffer", prof.data_ptr())
f()
torch.cuda.synchronize()
L.call_host("mrssm_pl_set_profile_buffer", None)
p = prof.cpu().reshape(148, 16, 8)
names = ["P:start", "P:loop", "M:afull", "M:done", "E:accfull", "E:end", "M:accfree", "M:mmas"]
t0 = int(p[0, 0, 0])
for it in range(2, 8):
    print("  tile", it, " ".join(f"{names[s]}={int(p[0, it, s]) - t0 if p[0, it, s] else -1:>8d}" for s in (1, 2, 6, 7, 3, 4, 5)))
