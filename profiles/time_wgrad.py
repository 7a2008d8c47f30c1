"""Time the plane weight-gradient kernel on every production layer of the 64x64 stacks (n frames, production layouts, fused bias
gradient): python profiles/time_wgrad.py [n] [layer ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import torch
from sweep_plans import DEV, L, ops, timed

# name: (Hl, Cl, Hs, Cs, k, s2d, dbias_from)
WG = {"E4": (6, 128, 2, 256, 4, 0, 1), "E1": (64, 3, 31, 32, 4, 1, 1), "E2": (31, 32, 14, 64, 4, 0, 1), "E3": (14, 64, 6, 128, 4, 0, 1),
      "D2": (13, 64, 5, 128, 5, 0, 2), "D3": (30, 32, 13, 64, 6, 0, 2), "D4": (64, 3, 30, 32, 6, 1, 2)}


def build(name, n):
    Hl, Cl, Hs, Cs, k, s2d, frm = WG[name]
    Clp, Csp = (16 if s2d else ops.pad8(Cl)), ops.pad16(Cs)
    if s2d:
        lg = ops.new_act(n, (Hl + 1) // 2, (Hl + 1) // 2, 16, L.PLANAR, DEV)
    else:
        lg = ops.new_act(n, Hl, Hl, Clp, L.PARITY, DEV)
    sm = ops.new_act(n, Hs, Hs, Csp, L.PLANAR, DEV)
    lg[0].normal_()
    sm[0].normal_()
    dw = torch.zeros(Cs, Cl, k, k, device=DEV)
    db = torch.zeros(Cs if frm == 1 else Cl, device=DEV)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    flops = 2.0 * n * Hs * Hs * Cs * Cl * k * k
    return (lambda: ops.pl_conv_wgrad(gp, lg[1], sm[1], L.ptr(dw), Cl * k * k, k * k, Cs, Cl, s2d_cq=Cl if s2d else 0, dbias=db, dbias_from=frm)), flops


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50176
    tot = 0.0
    for name in (sys.argv[2:] or list(WG)):
        f, flops = build(name, n)
        ms = timed(f, reps=3)
        tot += ms
        print(f"{name} wgrad {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
    print(f"total {tot:.3f} ms")
