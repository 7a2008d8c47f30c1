"""In-kernel timeline of one time step of the tcgen05 rollout (clock64 stamps of CTA 0: producer, MMA issuer, first
epilogue warp).  Usage on the GPU box: python profiles/timeline_rollout_tc.py [B] [T] > gpurun_out/rollout_tc_timeline.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "multimodal-rssm_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from mrssm_b200 import _lib as L, ops          # noqa: E402
from test_gpu_rollout_tc import _params          # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 49
D, S, H, A = 200, 30, 200, 3
emb_sizes = (0, 1024, 128)
gen = torch.Generator(device=DEV).manual_seed(0)
table = ops.FusionTable(3, S, "MoPoE")
spec = ops.RolloutSpec(D, S, H, A, ops.RELU, 0.1, table, [e > 0 for e in emb_sizes])
params = [p_.requires_grad_(True) for p_ in _params(gen, D, S, H, A, emb_sizes)]
rn = lambda *s: torch.randn(*s, device=DEV, generator=gen)
ins = [rn(B, S), rn(T, B, A), (rn(B, D) * 0.5).requires_grad_(True), torch.ones(T, B, device=DEV), rn(T, B, S), rn(T, B, S)]
embs = [rn(T, B, e) for e in emb_sizes if e > 0]
ops.set_bf16_mode(True)
prof = torch.zeros(4 * 512, device=DEV, dtype=torch.int64)
for it in range(3):
    if it == 2:
        L.call_host("mrssm_rollout_tc_set_profile_buffer", prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.profile = []
    outs = ops.RolloutFn.apply(spec, True, False, *ins, *embs, *params)
    torch.cuda.synchronize()
    for name, tag, work, a, b in L.profile:
        print(f"iter {it} {name}:{tag} {a.elapsed_time(b):.3f} ms")
    L.profile = None
L.call_host("mrssm_rollout_tc_set_profile_buffer", None)
p = prof.cpu().view(4, 512)
# backward: same buffer, epilogue warp 0 of CTA 0
prof2 = torch.zeros(4 * 512, device=DEV, dtype=torch.int64)
gouts = [torch.randn_like(o) for o in outs]
for it in range(3):
    for t_ in [ins[2]] + params:
        t_.grad = None
    outs = ops.RolloutFn.apply(spec, True, False, *ins, *embs, *params)
    if it == 2:
        L.call_host("mrssm_rollout_tc_set_profile_buffer", prof2.data_ptr())
    L.profile = []
    torch.autograd.backward(outs, gouts)
    torch.cuda.synchronize()
    for name, tag, work, a, b in L.profile:
        if "rollout" in name:
            print(f"bwd iter {it} {name}:{tag} {a.elapsed_time(b):.3f} ms")
    L.profile = None
L.call_host("mrssm_rollout_tc_set_profile_buffer", None)
v = [int(x) for x in prof2.cpu().view(4, 512)[2] if int(x) > 0]
print("--- BWD epilogue warp 0: step start; per half-head (after wait, after signal) x 2NH; Ed (after wait, after signal); Ee (after wait, after signal); Ef (after wait, before step_a, after signal)")
print(" ".join(str(x - v[0]) for x in v))
print("deltas:", " ".join(str(b - a) for a, b in zip(v, v[1:])))
w = [int(x) for x in prof2.cpu().view(4, 512)[3] if int(x) > 0]
print("--- BWD gate backward (Ed), per chunk: loop top, after TMEM load, after math (before global stores), after global stores")
print("deltas:", " ".join(str(b - a) for a, b in zip(w, w[1:])))
t0 = int(p[p > 0].min())
for role, name in enumerate(("producer (stamp after empty-wait, per tile)", "mma (before full-wait, after full-wait, after issue+commit; per tile)",
                             "epilogue warp 0 (before wait / after wait / after signal per unit)",
                             "epilogue warp 0 inside E2 (per chunk: before last ld, after wait::ld, after math; per half: loop end, after fences, after arrive)")):
    v = [int(x) - t0 for x in p[role] if int(x) > 0]
    print(f"--- {name}: {len(v)} stamps, span {v[-1] - v[0] if v else 0} cycles")
    print(" ".join(str(x) for x in v))
