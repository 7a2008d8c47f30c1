"""Data parallelism over sequence batches: one process per GPU, torch.distributed
(NCCL over NVLink) for the single exchange step of the path — a SUM all-reduce of the flat gradient
buffer, issued bucket by bucket under backward; the 1/world average is folded into the fused clip+Adam kernel
(``grad_scale``).

Every (b) is independent through encoders, rollout, decoders and the per-(t,b) free-nats clamp, so with
equal per-rank batches mean(rank gradients) == global-batch gradient exactly.  The reference has no
distributed code at all; this is the one strategy the new build adds.
"""
import os

import torch
import torch.distributed as dist


class DataParallel:
    """Gradient exchange of the data-parallel train step, overlapped with backward.

    The flat gradient buffer is cut into buckets that follow the order in which backward completes them — decoders first,
    then the transition model (the rollout's BPTT and its deferred weight-gradient GEMMs), then the encoders and whatever is
    left (reward head, alignment padding).  ``watch(bucket, tensors)`` hangs an autograd hook on a tensor whose gradient is
    produced right after the bucket's last weight-gradient kernel was enqueued (the decoders' input latent, the encoders'
    output embedding); the hook issues an asynchronous SUM all-reduce of that slice (NCCL orders it after the kernels already
    on the compute stream and runs it on its own stream, under the rest of backward).  ``all_reduce_grads`` launches what is
    left and makes the compute stream wait for all of it before the fused clip + Adam step."""

    def __init__(self, model, overlap=None):
        assert dist.is_initialized()
        if overlap is None:          # MRSSM_DP_OVERLAP=0: one all-reduce of the whole buffer after backward (A/B switch for measurements)
            overlap = os.environ.get("MRSSM_DP_OVERLAP", "1") != "0"
        self.world = dist.get_world_size()
        self.model = model
        self.overlap = overlap
        self._pending, self._launched = [], []
        model.dp = self
        self.attach(model.model_optimizer)

    def attach(self, opt):
        """(Re-)bind to the model's optimiser: called at construction and whenever the model rebuilds it (load_model ->
        _init_optimizer): the fresh optimiser has new flat buffers, so the weights are broadcast again (rank 0's checkpoint
        wins), the buckets are re-planned over the new gradient buffer and the 1/world scale is set."""
        opt.grad_scale = 1.0 / self.world
        dist.broadcast(opt.flat_p, src=0)                 # identical weights everywhere
        self._buckets = self._plan_buckets(self.model, opt) if self.overlap else None
        self._pending, self._launched = [], []
        self._opt = opt

    @staticmethod
    def _plan_buckets(model, opt):
        """{name: (lo, hi)} element ranges of flat_g, or None when the model does not expose the reference's module objects."""
        total = opt.flat_g.numel()
        base = opt.flat_g.data_ptr()
        getters = (("decoder", lambda: model.observation_model.get_model_params()),
                   ("transition", lambda: model.transition_model.get_model_params()),
                   ("encoder", lambda: model.encoder.get_model_params()))
        ranges = {}
        try:
            for name, get in getters:
                ps = [p for p in get() if p.grad is not None]
                if not ps:
                    return None
                lo = min((p.grad.data_ptr() - base) // 4 for p in ps)
                hi = max((p.grad.data_ptr() - base) // 4 + p.numel() for p in ps)
                if lo < 0 or hi > total:
                    return None
                ranges[name] = (int(lo), int(hi))
        except AttributeError:
            return None
        spans = sorted(ranges.values())
        if any(a[1] > b[0] for a, b in zip(spans, spans[1:])):
            return None                                   # interleaved modules: one all-reduce at the end
        return ranges

    def watch(self, name, tensors):
        """Launch bucket ``name`` as soon as the gradient of (the first differentiable one of) ``tensors`` is computed."""
        if self.world == 1 or self._buckets is None or name not in self._buckets:
            return
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.requires_grad:
                def hook(grad, name=name):
                    self._launch(name)
                    return None
                t.register_hook(hook)
                return

    def _launch(self, name, final=False):
        if name in self._launched:
            return
        if not final:
            from . import ops
            if name == "decoder" and ops.side_pending():
                return      # its weight gradients are queued for / running on the side stream: exchanged from all_reduce_grads
        lo, hi = self._buckets[name]
        self._launched.append(name)
        self._pending.append(dist.all_reduce(self.model.model_optimizer.flat_g[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def all_reduce_grads(self, opt):
        if opt is not getattr(self, "_opt", None):        # optimiser replaced behind our back: re-bind before exchanging
            self.attach(opt)
        opt.grad_scale = 1.0 / self.world                 # read at step time: the SUM below is turned into the mean by clip+Adam
        if self.world == 1:
            return
        if self._buckets is None:
            dist.all_reduce(opt.flat_g, op=dist.ReduceOp.SUM)
            return
        for name in self._buckets:
            self._launch(name, final=True)       # (the backward scope has joined the side stream by now)
        # the gaps between the buckets (reward head, padding)
        pos = 0
        for lo, hi in sorted(self._buckets.values()) + [(opt.flat_g.numel(), opt.flat_g.numel())]:
            if lo > pos:
                self._pending.append(dist.all_reduce(opt.flat_g[pos:lo], op=dist.ReduceOp.SUM, async_op=True))
            pos = max(pos, hi)
        for w in self._pending:
            w.wait()
        self.last_order = list(self._launched)          # for tests / logging
        self._pending, self._launched = [], []


def init_from_env(backend=None):
    """Read RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (torchrun) and initialise the process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world
