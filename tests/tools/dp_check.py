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
    results = []
    for use_dp in (False, True):
        model, P = U.build_product(oc, B, T, dev, bf16=True)
        if use_dp:
            dp = DataParallel(model)
            assert dp._buckets is not None, "bucket plan failed"
        for step in range(2):
            batch, noise = O.synthetic_batch(oc, B, T, seed=50 + step)
            U.product_step(model, oc, batch, noise, dev)
        if use_dp:
            assert dp.last_order[:2] == ["decoder", "transition"], dp.last_order
        results.append(model.model_optimizer.flat_p.detach().clone())
    diff = float((results[0] - results[1]).abs().max())
    scale = float(results[0].abs().max())
    print(f"rank {rank}: max |p_single - p_dp| = {diff:.3e} (scale {scale:.3f}), bucket order {dp.last_order}")
    assert diff <= 1e-5 * max(1.0, scale), diff
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
