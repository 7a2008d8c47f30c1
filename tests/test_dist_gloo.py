"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel exchange step (mrssm_b200/dist.py) —
rank-0 parameter broadcast, SUM all-reduce of the flat gradient buffer, 1/world folded into the optimiser scale.
With equal per-rank batches the averaged rank gradients equal the global-batch gradient; that identity
is checked here on the oracle's loss with the batch split across the two ranks."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Opt:
    def __init__(self, n, rank):
        g = torch.Generator().manual_seed(100 + rank)
        self.flat_p = torch.randn(n, generator=g)
        self.flat_g = torch.randn(n, generator=g)
        self.grad_scale = 1.0


class _Model:
    def __init__(self, opt):
        self.model_optimizer = opt
        self.dp = None


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "multimodal-rssm_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mrssm_b200.dist import DataParallel, init_from_env
    r, l, w = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    opt = _Opt(1000, rank)
    g_local = opt.flat_g.clone()
    model = _Model(opt)
    dp = DataParallel(model)
    assert model.dp is dp and opt.grad_scale == 1.0 / world
    p_after = opt.flat_p.clone()
    dp.all_reduce_grads(opt)
    q.put((rank, p_after, g_local, opt.flat_g.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_broadcast_and_sum_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, p0, g0, s0), (_, p1, g1, s1) = out
    torch.testing.assert_close(p0, p1)                       # rank-0 weights everywhere
    torch.testing.assert_close(s0, g0 + g1)                  # SUM all-reduce
    torch.testing.assert_close(s0, s1)


def test_mean_of_rank_gradients_is_global_batch_gradient():
    """The identity the DP design rests on, on the oracle: loss means over (t,b) with per-(t,b) clamps."""
    for p in (ROOT, os.path.join(ROOT, "multimodal-rssm_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import mrssm_oracle as O
    oc = O.OracleConfig(fusion="MoPoE")
    P = O.make_params(oc, seed=0)
    batch, noise = O.synthetic_batch(oc, 4, 5, seed=7)

    def grads(sl):
        b = {"obs": {n: t[:, sl] for n, t in batch["obs"].items()}, "actions": batch["actions"][:, sl],
             "rewards": batch["rewards"][:, sl], "nonterminals": batch["nonterminals"][:, sl]}
        nz = {k: v[:, sl] for k, v in noise.items()}
        return O.train_step({k: v.clone() for k, v in P.items()}, {}, oc, b, nz)["grads"]

    g_all, g_a, g_b = grads(slice(0, 4)), grads(slice(0, 2)), grads(slice(2, 4))
    for k in g_all:
        scale = float(g_all[k].abs().max()) + 1e-12
        assert float((0.5 * (g_a[k] + g_b[k]) - g_all[k]).abs().max()) <= 1e-4 * scale + 1e-7, k


class _Mod:
    def __init__(self, params):
        self._p = params

    def get_model_params(self):
        return self._p


def _bucket_worker(rank, world, port, q, late_decoder=False):
    for p in (ROOT, os.path.join(ROOT, "multimodal-rssm_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mrssm_b200.dist import DataParallel, init_from_env
    init_from_env(backend="gloo")
    # a flat gradient buffer laid out like FusedClipAdam's: transition | decoder | reward (no bucket) | encoder, with padding
    sizes = dict(transition=37, decoder=53, reward=11, encoder=29)
    total = sum((n + 3) // 4 * 4 for n in sizes.values())
    opt = _Opt(total, rank)
    opt.flat_g.zero_()
    params, off = {}, 0
    g = torch.Generator().manual_seed(5)                                    # same weights on every rank
    for name, n in sizes.items():
        p = torch.nn.Parameter(torch.randn(n, generator=g))
        p.grad = opt.flat_g[off:off + n]
        params[name] = p
        off += (n + 3) // 4 * 4
    model = _Model(opt)
    model.observation_model = _Mod([params["decoder"]])
    model.transition_model = _Mod([params["transition"]])
    model.encoder = _Mod([params["encoder"]])
    dp = DataParallel(model)
    assert dp._buckets is not None and set(dp._buckets) == {"decoder", "transition", "encoder"}
    # forward: x -> encoder -> emb -> transition -> z -> decoder -> loss (rank-dependent data); reward head gets a gradient too.
    # Like the product's autograd Functions, a layer writes its weight gradient into p.grad INSIDE backward and returns None for
    # the parameter, so the gradient exists before the hook of the layer's input fires.
    class Lin(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, p, n_out):
            ctx.save_for_backward(x)
            ctx.p = p
            return (x * p.detach()).sum() * torch.ones(n_out)

        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            if late_decoder and ctx.p is params["decoder"]:
                # like the product in bf16 mode: the decoder's weight gradient is only QUEUED here (ops.side_defer) and lands
                # when the side stream is joined, after backward
                deferred.append((ctx.p, g.sum() * x))
            else:
                ctx.p.grad += g.sum() * x
            return g.sum() * ctx.p.detach(), None, None

    deferred = []

    def run(dp_):
        opt.flat_g.zero_()
        x = torch.full((29,), float(rank + 1), requires_grad=True)
        emb = Lin.apply(x, params["encoder"], 37)
        if dp_ is not None:
            dp_.watch("transition", [emb])
        z = Lin.apply(emb, params["transition"], 53)
        if dp_ is not None:
            dp_.watch("decoder", [z])
        loss = Lin.apply(z, params["decoder"], 1).sum()
        loss.backward()
        params["reward"].grad += float(rank + 1)

    run(None)
    for p_, g_ in deferred:
        p_.grad += g_
    deferred.clear()
    local = opt.flat_g.clone()                          # this rank's gradient, no exchange
    order = []
    orig = dp._launch
    dp._launch = lambda name, final=False: (order.append(name), orig(name, final=final))[1]
    if late_decoder:
        from mrssm_b200 import ops
        ops.side_pending = lambda: bool(deferred)            # what ops reports while weight gradients are queued / in flight
    run(dp)
    launched_in_backward = list(order)
    for p_, g_ in deferred:                                  # ops.side_wgrad_scope's exit: the queued gradients have landed
        p_.grad += g_
    deferred.clear()
    dp.all_reduce_grads(opt)
    q.put((rank, local, opt.flat_g.clone(), launched_in_backward, dp.last_order))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_launches_under_backward_gloo():
    """Buckets are exchanged in the order backward completes them (decoder, transition from hooks; encoder and the
    un-bucketed remainder at the end) and the result equals one SUM all-reduce of the whole buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, s0, in_bwd0, ord0), (_, l1, s1, in_bwd1, ord1) = out
    assert in_bwd0 == in_bwd1 == ["decoder", "transition"]          # launched by the hooks, before backward returned
    assert ord0 == ["decoder", "transition", "encoder"]
    # note: the hook of a bucket fires after that bucket's gradients were written, so the early launch sums final values
    torch.testing.assert_close(s0, s1)
    torch.testing.assert_close(s0, l0 + l1)


def _reattach_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "multimodal-rssm_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mrssm_b200.dist import DataParallel, init_from_env
    init_from_env(backend="gloo")
    model = _Model(_Opt(64, rank))
    dp = DataParallel(model)
    # load_model() rebuilds the optimiser (base/algo.py: _init_optimizer): fresh flat buffers, grad_scale back to 1.0,
    # rank-dependent weights (each rank read its own file / rank 0 only has the checkpoint)
    new = _Opt(64, 10 + rank)
    model.model_optimizer = new
    assert new.grad_scale == 1.0
    g_local = new.flat_g.clone()
    dp.all_reduce_grads(new)                     # notices the replaced optimiser: re-binds, re-broadcasts, sets the scale
    q.put((rank, new.flat_p.clone(), g_local, new.flat_g.clone(), new.grad_scale))
    # and the explicit path the algorithm layer takes
    newer = _Opt(64, 20 + rank)
    model.model_optimizer = newer
    dp.attach(newer)
    q.put((rank, newer.flat_p.clone(), None, None, newer.grad_scale))
    dist.barrier()
    dist.destroy_process_group()


def test_optimizer_rebuilt_after_attach_is_rebound_gloo():
    """load_model after DataParallel was attached must not apply world x the mean gradient nor leave the ranks
    with different weights."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reattach_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    first = sorted([o for o in out if o[2] is not None], key=lambda t: t[0])
    second = sorted([o for o in out if o[2] is None], key=lambda t: t[0])
    (_, p0, g0, s0, sc0), (_, p1, g1, s1, sc1) = first
    assert sc0 == sc1 == 0.5
    torch.testing.assert_close(p0, p1)
    torch.testing.assert_close(s0, g0 + g1)
    torch.testing.assert_close(s0, s1)
    torch.testing.assert_close(second[0][1], second[1][1])
    assert second[0][4] == second[1][4] == 0.5


def test_decoder_bucket_waits_for_side_stream_weight_gradients_gloo():
    """bf16 mode queues the decoder's weight gradients for a side stream (ops.side_defer) and joins it after backward: while they are
    pending the decoder bucket must NOT go out from its autograd hook (it would exchange a buffer its kernels have not written); it is
    exchanged from all_reduce_grads instead, and the result is still the SUM over ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, s0, in_bwd0, ord0), (_, l1, s1, in_bwd1, ord1) = out
    assert in_bwd0 == in_bwd1 == ["decoder", "transition"]          # both hooks fired ...
    assert ord0 == ord1 == ["transition", "decoder", "encoder"]     # ... but the decoder bucket was held back until the join
    torch.testing.assert_close(s0, s1)
    torch.testing.assert_close(s0, l0 + l1)
