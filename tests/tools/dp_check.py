"""Multi-GPU check of the overlapped gradient exchange (run under torchrun on a GPU box):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/dp_check.py
Every rank trains the same model on the SAME batch, so the averaged gradient equals the local one and two DP steps must
leave the parameters equal (up to summation order of the all-reduce) to those of a plain single-process model."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "multimodal-rssm_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from mrssm_b200.dist import DataParallel, init_from_env          # noqa: E402
from oracle import mrssm_oracle as O                             # noqa: E402
from tests import parity_util as U                               # noqa: E402


def main():
    rank, local, world = init_from_env()
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    oc = U.oracle_cfg("MoPoE")
    B, T = 8, 6
    results, grads = [], []
    for use_dp in (False, True):
        model, P = U.build_product(oc, B, T, dev, bf16=True)
        if use_dp:
            dp = DataParallel(model)
            assert dp._buckets is not None, "bucket plan failed"
        for step in range(2):
            batch, noise = O.synthetic_batch(oc, B, T, seed=50 + step)
            U.product_step(model, oc, batch, noise, dev)
            if step == 0:       # the flat gradient buffer as the optimiser saw it (DP: the SUM over ranks; 1/world is applied in clip+Adam)
                opt = model.model_optimizer
                grads.append(opt.flat_g.detach().clone() * float(opt.grad_scale))
        if use_dp:
            # the transition bucket goes out under backward; the decoder bucket too unless its weight gradients were queued for the
            # side stream (bf16 mode: ops.side_wgrad_scope), in which case it is exchanged right after the join
            assert dp.last_order in (["decoder", "transition", "encoder"], ["transition", "decoder", "encoder"]), dp.last_order
        results.append(model.model_optimizer.flat_p.detach().clone())
    gdiff = float((grads[0] - grads[1]).abs().max()) / float(grads[0].abs().max())
    diff = float((results[0] - results[1]).abs().max())
    scale = float(results[0].abs().max())
    print(f"rank {rank}: step-1 gradient: max |g_single - mean_ranks g_dp| / max|g| = {gdiff:.3e}; after 2 steps max |p_single - p_dp| = {diff:.3e} "
          f"(scale {scale:.3f}), bucket order {dp.last_order}")
    # identical batches on every rank: the averaged gradient is the local one up to the summation order of fp32 atomics (two
    # single-process runs differ by the same ~4e-7, profiles/determinism_check.py)
    assert gdiff <= 1e-5, gdiff
    # Adam's first updates are lr * g / (|g| + 1e-7): an element whose gradient is summation noise moves by +-lr either way, so the
    # parameters of two runs agree to a few lr, not to rounding (single-process runs: 1.4e-3 at lr 1e-3)
    assert diff <= 3e-3 * max(1.0, scale), diff
    # different batches per rank: the exchanged gradient must be the MEAN of the per-batch gradients (this is what catches an
    # all-reduce that reads a bucket before its last weight-gradient kernel — side stream included — has written it)
    per_batch = []
    for r in range(world):
        m, _ = U.build_product(oc, B, T, dev, bf16=True)
        batch, noise = O.synthetic_batch(oc, B, T, seed=70 + r)
        U.product_step(m, oc, batch, noise, dev)
        per_batch.append(m.model_optimizer.flat_g.detach().clone())
        del m
    want = sum(per_batch) / world
    m, _ = U.build_product(oc, B, T, dev, bf16=True)
    DataParallel(m)
    batch, noise = O.synthetic_batch(oc, B, T, seed=70 + rank)
    U.product_step(m, oc, batch, noise, dev)
    got = m.model_optimizer.flat_g.detach().clone() * float(m.model_optimizer.grad_scale)
    mdiff = float((got - want).abs().max()) / float(want.abs().max())
    print(f"rank {rank}: per-rank batches: max |mean_ranks g - g_dp| / max|g| = {mdiff:.3e}")
    assert mdiff <= 1e-5, mdiff
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
