import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/multimodal-rssm_b200")
from tests.test_gpu_rollout_step import CASES, _setup
for case in CASES[:5]:
    ops, spec, observe, det, ins, embs, params, rn, E = _setup(case, 11, True, act=__import__("mrssm_b200.ops").ops.ELU)
    leaves = ins[:3] + embs + params
    gouts, results = None, []
    for bf16 in (False, True):
        for t in leaves: t.grad = None
        ops.set_bf16_mode(bf16)
        try:
            outs = ops.RolloutFn.apply(spec, observe, det, *ins, *embs, *params)
            if gouts is None: gouts = [rn(*o.shape) / o.shape[-1] ** 0.5 for o in outs]
            torch.autograd.backward(outs, gouts)
        finally:
            ops.set_bf16_mode(False)
        results.append([t.grad.clone() for t in leaves])
    names = ["g_prev_state", "g_actions", "g_prev_belief"] + [f"g_emb{i}" for i in range(len(embs))] + [f"g_param{i}" for i in range(len(params))]
    print(case)
    print("  ", " ".join("%s=%.3g" % (n, float((o - r).norm() / (r.norm() + 1e-12))) for n, r, o in zip(names, *results)))
