"""Per-tile clock64 timeline of the plane fwd-type kernel for one layer (tuning aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-rssm_b200"))
import torch
from mrssm_b200 import _lib as L, ops

def main(op="down", n=4096, Hl=64, Cl=3, Hs=31, Cs=32, k=4, f32=0, mask=0):
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    Clp, Csp = ops.pad8(Cl), ops.pad16(Cs)
    lbf = torch.randn(n, Hl, Hl, Clp, device=dev, generator=g)
    lb = ops.pl_import(L.nhwc(lbf, Hl, Hl, Clp), n, Hl, Hl, Clp, Clp, L.PARITY, dev)
    sbf = torch.randn(n, Hs, Hs, Csp, device=dev, generator=g)
    sb = ops.pl_import(L.nhwc(sbf, Hs, Hs, Csp), n, Hs, Hs, Csp, Csp, L.PLANAR, dev)
    w = torch.randn(Cs, Cl, k, k, device=dev, generator=g) / (Cl * k * k) ** 0.5
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    prof = torch.zeros(148 * 16 * 8, device=dev, dtype=torch.int64)
    if op == "down":
        wp = ops.pl_pack_weight(w, ops.DOWN, Csp, Clp)
        bias = torch.zeros(Cs, device=dev)
        out = ops.new_act(n, Hs, Hs, Csp, L.PARITY, dev)
        mk = ops.new_act(n, Hs, Hs, Csp, L.PLANAR, dev)
        mk[0].fill_(1.0)
        f = lambda: ops.pl_conv_down(gp, lb[1], out[1], wp, bias, Cs, Csp, act=0 if mask else 1, mask=mk[1] if mask else None, mask_mode=1 if mask else 0)
    else:
        wp = ops.pl_pack_weight(w, ops.UP, Csp, Clp)
        bias = torch.zeros(Cl, device=dev)
        if f32:
            out = torch.zeros(n, Cl, Hl, Hl, device=dev)
            f = lambda: ops.pl_conv_up(gp, L.nchw(out, Hl, Hl, Cl), sb[1], wp, bias, Cl, Clp)
        else:
            out = ops.new_act(n, Hl, Hl, Clp, L.PLANAR, dev)
            mk = ops.new_act(n, Hl, Hl, Clp, L.PARITY, dev)
            mk[0].fill_(1.0)
            f = lambda: ops.pl_conv_up(gp, out[1], sb[1], wp, bias, Cl, Clp, act=0 if mask else 1, mask=mk[1] if mask else None, mask_mode=1 if mask else 0)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    print(f"{op} n={n} [{Hl}x{Hl}x{Cl}<->{Hs}x{Hs}x{Cs} k{k}] f32={f32}: {e0.elapsed_time(e1):.3f} ms")
    L.call_host("mrssm_pl_set_profile_buffer", prof.data_ptr())
    f(); torch.cuda.synchronize()
    L.call_host("mrssm_pl_set_profile_buffer", None)
    p = prof.cpu().reshape(148, 16, 8)
    names = ["P:start", "P:loop", "M:afull", "M:done", "E:accfull", "E:end", "M:accfree", "M:mmas"]
    for cta in (0,):
        t0 = int(p[cta, 0, 0])
        print(f"CTA {cta} (cycles since start)")
        for it in range(0, 4):
            print("  tile", it, " ".join(f"{names[s]}={int(p[cta, it, s]) - t0 if p[cta, it, s] else -1:>8d}" for s in (1, 2, 6, 7, 3, 4, 5)))

if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0], *[int(x) for x in a[1:]])
