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
