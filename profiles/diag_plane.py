"""Bring-up diagnostic for csrc/conv_plane.cu on a B200: every geometry x {down, up, dgrad+mask, wgrad} against the
exact fp32 CUDA-core kernels, for both LBO/SBO variants of the un-swizzled descriptors.  Prints max errors."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-rssm_b200"))
import torch
from mrssm_b200 import _lib as L, ops

DEV = "cuda:0"
GEOMS = [(3, 64, 3, 31, 32, 4), (3, 31, 32, 14, 64, 4), (5, 14, 64, 6, 128, 4), (30, 6, 128, 2, 256, 4),
         (5, 13, 64, 5, 128, 5), (2, 30, 32, 13, 64, 6), (2, 64, 3, 30, 32, 6), (2, 128, 3, 63, 16, 4), (1, 128, 3, 62, 32, 6)]
p8 = lambda c: (c + 7) // 8 * 8
p16 = lambda c: (c + 15) // 16 * 16
rnd = lambda t: t.to(torch.bfloat16).to(torch.float32)


def run(which, g):
    n, Hl, Cl, Hs, Cs, k = g
    gen = torch.Generator(device=DEV).manual_seed(1)
    large = rnd(torch.randn(n, Hl, Hl, Cl, device=DEV, generator=gen))
    small = rnd(torch.randn(n, Hs, Hs, Cs, device=DEV, generator=gen))
    w = rnd(torch.randn(Cs, Cl, k, k, device=DEV, generator=gen) / (Cl * k * k) ** 0.5)
    geom = (n, Hl, Hl, Cl, Hs, Hs, Cs, k)
    Clp, Csp = p8(Cl), p16(Cs)
    lb = ops.tc_to_bf16(L.nhwc(large, Hl, Hl, Cl), n, Hl, Hl, Cl, DEV, Cpad=Clp)
    sb = ops.tc_to_bf16(L.nhwc(small, Hs, Hs, Cs), n, Hs, Hs, Cs, DEV, Cpad=Csp)
    gp = (n, Hl, Hl, Clp, Hs, Hs, Csp, k)
    if which == "down":
        bias = torch.randn(Cs, device=DEV)
        ref = torch.empty_like(small)
        ops._conv("mrssm_conv_down", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(ref, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k, L.ptr(bias), ops.RELU)
        wp = ops.pl_pack_weight(w, ops.DOWN, Csp, Clp)
        out = torch.full((n, Hs, Hs, Csp), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.pl_conv_down(gp, L.nhwc(lb, Hl, Hl, Clp), L.nhwc(out, Hs, Hs, Csp), wp, bias, Cs, Csp, act=ops.RELU)
        torch.cuda.synchronize()
        return float((out[..., :Cs].float() - ref).abs().max()), float(ref.abs().max())
    if which in ("up", "dgrad"):
        if which == "up" and Hl != 2 * (Hs - 1) + k:
            return None
        bias = torch.randn(Cl, device=DEV) if which == "up" else None
        ref = torch.empty_like(large)
        ops._conv("mrssm_conv_up", geom, L.nhwc(ref, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(w), Cl * k * k, k * k,
                  L.ptr(bias) if bias is not None else None, ops.RELU if which == "up" else 0,
                  L.ptr(large) if which == "dgrad" else None, ops.RELU if which == "dgrad" else 0)
        wp = ops.pl_pack_weight(w, ops.UP, Csp, Clp)
        out = torch.full((n, Hl, Hl, Clp), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.pl_conv_up(gp, L.nhwc(out, Hl, Hl, Clp), L.nhwc(sb, Hs, Hs, Csp), wp, bias, Cl, Clp,
                       act=ops.RELU if which == "up" else 0,
                       mask=L.nhwc(lb, Hl, Hl, Clp) if which == "dgrad" else None, mask_mode=ops.RELU if which == "dgrad" else 0)
        torch.cuda.synchronize()
        return float((out[..., :Cl].float() - ref).abs().max()), float(ref.abs().max())
    ref = torch.zeros_like(w)
    ops._conv("mrssm_conv_wgrad", geom, L.nhwc(large, Hl, Hl, Cl), L.nhwc(small, Hs, Hs, Cs), L.ptr(ref), Cl * k * k, k * k)
    out = torch.zeros_like(w)
    ops.pl_conv_wgrad(gp, L.nhwc(lb, Hl, Hl, Clp), L.nhwc(sb, Hs, Hs, Csp), L.ptr(out), Cl * k * k, k * k, Cs, Cl)
    torch.cuda.synchronize()
    return float((out - ref).abs().max()), float(ref.abs().max())


def main():
    if len(sys.argv) > 1:                       # child: one (variant, which) over all geometries
        variant, which = int(sys.argv[1]), sys.argv[2]
        L.call_host("mrssm_pl_set_debug", 0, variant)
        L.call_host("mrssm_pl_set_debug", 1, variant)
        for g in GEOMS:
            try:
                r = run(which, g)
            except Exception as e:               # noqa
                print(f"variant {variant} {which:6s} {g}: EXC {str(e)[:300]}", flush=True)
                if "CUDA" in str(e) or "illegal" in str(e) or "launch" in str(e):
                    return
                continue
            if r is not None:
                print(f"variant {variant} {which:6s} {g}: max_err {r[0]:.4g} (ref max {r[1]:.4g}) {'OK' if r[0] <= 2e-2 * max(1.0, r[1]) else 'BAD'}", flush=True)
        return
    for which in ("down", "up", "dgrad", "wgrad"):
        for variant in (0, 1):
            try:
                out = subprocess.run([sys.executable, __file__, str(variant), which], capture_output=True, text=True, timeout=240)
                print(out.stdout, end="")
                if out.returncode != 0:
                    print(f"variant {variant} {which}: exit {out.returncode}: {out.stderr[-800:]}")
            except subprocess.TimeoutExpired:
                print(f"variant {variant} {which}: TIMEOUT")


if __name__ == "__main__":
    main()
