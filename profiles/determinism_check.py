"""Run-to-run determinism of two optimisation steps in bf16 mode (same weights, batch, noise): max parameter difference between two
fresh models, and which parameter tensor carries it.  python profiles/determinism_check.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multimodal-rssm_b200"))
from oracle import mrssm_oracle as O
from tests import parity_util as U
dev = "cuda:0"
oc = U.oracle_cfg("MoPoE")
B, T = 8, 6
res, grads = [], []
for run in range(2):
    model, P = U.build_product(oc, B, T, dev, bf16=True)
    named = U.named_params(model, oc)
    for step in range(2):
        batch, noise = O.synthetic_batch(oc, B, T, seed=50 + step)
        U.product_step(model, oc, batch, noise, dev)
        if step == 0:
            grads.append({k: v.grad.detach().clone() for k, v in named.items()})
    res.append({k: v.detach().clone() for k, v in named.items()})
worst = sorted(((float((res[0][k] - res[1][k]).abs().max()), k) for k in res[0]), reverse=True)[:5]
print("side_wgrad", os.environ.get("MRSSM_SIDE_WGRAD", "1"), "worst param diffs:", worst)
gw = sorted(((float((grads[0][k] - grads[1][k]).abs().max() / (grads[0][k].abs().max() + 1e-30)), k) for k in grads[0]), reverse=True)[:5]
print("  worst step-1 gradient diffs (relative to the tensor's max):", gw)
