"""Is the per-step rollout (config 5 shapes) bound by the host's launch rate?  Times RolloutFn forward + backward eagerly and
as a replayed CUDA graph of the very same launches.  python profiles/step_graph_probe.py [B] [T]"""
import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/multimodal-rssm_b200")
from tests.test_gpu_rollout_step import _setup
from mrssm_b200 import ops, _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 63
case = (1024, 64, 1024, 3, "MoPoE", (0, 1024, 128), T, B, False)
_, spec, observe, det, ins, embs, params, rn, E = _setup(case, 3, True, act=ops.ELU)
ops.set_bf16_mode(True)
gouts = None
def step():
    global gouts
    outs = ops.RolloutFn.apply(spec, observe, det, *ins, *embs, *params)
    if gouts is None:
        gouts = [rn(*o.shape) / o.shape[-1] ** 0.5 for o in outs]
    torch.autograd.backward(outs, gouts)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
step(); step()
n0 = L.launches; step(); print("launches per fwd+bwd:", L.launches - n0)
print("eager  ms (device events, wall): %.2f %.2f" % timed(step))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
print("graph  ms (device events, wall): %.2f %.2f" % timed(g.replay))
