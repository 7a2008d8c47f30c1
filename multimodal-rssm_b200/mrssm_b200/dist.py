"""Data parallelism over sequence batches: one process per GPU, torch.distributed
(NCCL over NVLink) for the single exchange step of the path — a SUM all-reduce of the flat gradient
buffer; the 1/world average is folded into the fused clip+Adam kernel (``grad_scale``).

Every (b) is independent through encoders, rollout, decoders and the per-(t,b) free-nats clamp, so with
equal per-rank batches mean(rank gradients) == global-batch gradient exactly.  The reference has no
distributed code at all; this is the one strategy the new build adds.
"""
import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, model, bucket_bytes=None):
        assert dist.is_initialized()
        self.world = dist.get_world_size()
        self.model = model
        opt = model.model_optimizer
        opt.grad_scale = 1.0 / self.world
        # identical weights everywhere: broadcast rank 0's flat parameter buffer once
        dist.broadcast(opt.flat_p, src=0)
        model.dp = self

    def all_reduce_grads(self, opt):
        if self.world > 1:
            dist.all_reduce(opt.flat_g, op=dist.ReduceOp.SUM)


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
