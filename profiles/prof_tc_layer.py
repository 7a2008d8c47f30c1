"""Run one tensor-core conv-family op on a realistic number of frames (for ncu captures and quick timing)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-rssm_b200"))
import torch
from mrssm_b200 import _lib as L, ops

def main(op="up", n=2048, Hl=64, Cl=3, Hs=30, Cs=32, k=6, reps=3):
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    large = torch.randn(n, Hl, Hl, Cl, device=dev, generator=g)
    small = torch.randn(n, Hs, Hs, Cs, device=dev, generator=g)
    w = torch.randn(Cs, Cl, k, k, device=dev, generator=g) / (Cl * k * k) ** 0.5
    Clp, Csp = ops.pad8(Cl), ops.pad8(Cs)
    lb = ops.tc_to_bf16(L.nhwc(large, Hl, Hl, Cl), n, Hl, Hl, Cl, dev)
    sb = ops.tc_to_bf16(L.nhwc(small, Hs, Hs, Cs), n, Hs, Hs, Cs, dev)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    bias_s, bias_l = torch.zeros(Cs, device=dev), torch.zeros(Cl, device=dev)
    if op == "up":
        wp = ops.tc_pack_weight(w, 1, Csp, Clp)
        out = torch.zeros(n, Hl, Hl, ops.pad16(Cl), device=dev, dtype=torch.bfloat16)
        f = lambda: ops.tc_conv_up(gp, L.nhwc(out, Hl, Hl, out.shape[-1]), L.nhwc(sb, Hs, Hs, Csp), wp, bias_l, Cl, act=1)
    elif op == "down":
        wp = ops.tc_pack_weight(w, 0, Csp, Clp)
        out = torch.zeros(n, Hs, Hs, ops.pad16(Cs), device=dev, dtype=torch.bfloat16)
        f = lambda: ops.tc_conv_down(gp, L.nhwc(lb, Hl, Hl, Clp), L.nhwc(out, Hs, Hs, out.shape[-1]), wp, bias_s, Cs, act=1)
    else:
        dw = torch.zeros_like(w)
        f = lambda: ops.tc_conv_wgrad(gp, L.nhwc(lb, Hl, Hl, Clp), L.nhwc(sb, Hs, Hs, Csp), L.ptr(dw), Cl * k * k, k * k, Cs, Cl)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * n * Hs * Hs * Cs * Cl * k * k
    print(f"{op} n={n} [{Hl}x{Hl}x{Cl}<->{Hs}x{Hs}x{Cs} k{k}]: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")

if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], *[int(x) for x in a[1:]])
