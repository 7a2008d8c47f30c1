"""Entry point — the B200 build's counterpart of the reference's
train/COBOTTA/SingleHoleDrilling/MRSSM/MRSSM/main.py (:37-49): compose the YAML tree under ./config, apply
`group.key=value` overrides from the command line, name the experiment, fix the seed, and hand over to
`algos.MRSSM.MRSSM.train.run`.  hydra / omegaconf are not available in this image, so the composition is done by
mrssm_b200.config.load_config (same `defaults` list, same override syntax).

    python main.py [--config-dir DIR] [group.key=value ...]
    torchrun --nproc-per-node N main.py ...        # data parallel: one process per GPU, per-rank batch = train.batch_size
"""
import argparse
import copy
import datetime
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PACKAGE_ROOT = os.path.normpath(os.path.join(HERE, "..", "..", "..", "..", ".."))      # where algos/ and utils/ live
if PACKAGE_ROOT not in sys.path:
    sys.path.insert(0, PACKAGE_ROOT)

import yaml  # noqa: E402

from mrssm_b200.config import load_config  # noqa: E402


def results_folder(cwd, experiment_name):
    """<cwd>/results/<experiment>/<date>/run_<k> with the first free k (the reference's naming, utils/logger.py:14-28)."""
    day = datetime.date.today()
    k = 0
    while os.path.exists(os.path.join(cwd, "results", experiment_name, str(day), f"run_{k}")):
        k += 1
    path = os.path.join(cwd, "results", experiment_name, str(day), f"run_{k}")
    os.makedirs(path, exist_ok=True)
    return path


def prepare(cfg_raw, experiment_name, tags, seed):
    cfg = copy.deepcopy(cfg_raw)
    cfg.main.experiment_name = f"{experiment_name}-seed_{seed}"
    cfg.main.tags = list(tags)
    cfg.main.seed = seed
    cfg.rssm.overshooting_distance = min(cfg.train.chunk_size, cfg.rssm.overshooting_distance)   # utils/logger.py:42
    return cfg


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config-dir", default=os.path.join(HERE, "config"))
    ap.add_argument("--cwd", default=HERE, help="directory the relative data paths of train.yaml are resolved against")
    ap.add_argument("overrides", nargs="*", help="group.key=value")
    args = ap.parse_args(argv)
    cfg_raw = load_config(args.config_dir, args.overrides)

    import torch
    from algos.MRSSM.MRSSM.train import run
    dp = None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        from mrssm_b200.dist import DataParallel, init_from_env
        _, local, _ = init_from_env()
        cfg_raw.main.device = f"cuda:{local}"       # the algorithm layer allocates on cfg.main.device (base/algo.py estimate_state)
        dp = DataParallel
    if cfg_raw.main.disable_cuda or not torch.cuda.is_available():
        raise SystemExit("main.py: the B200 build has no CPU path (cfg.main.disable_cuda / no CUDA device)")
    torch.cuda.set_device(torch.device(cfg_raw.main.device))

    models = []
    for seed in range(0, 1):
        cfg = prepare(cfg_raw, "RSSM", ["RSSM"], seed)
        rank0 = int(os.environ.get("RANK", "0")) == 0
        results_dir = results_folder(args.cwd, cfg.main.experiment_name) if rank0 else None
        if world > 1:
            import torch.distributed as dist
            box = [results_dir]
            dist.broadcast_object_list(box, src=0)
            results_dir = box[0]
        cfg.main.log_dir = results_dir
        if rank0:
            with open(os.path.join(results_dir, "hydra_config.yaml"), "w") as f:       # same file name the reference saves
                yaml.safe_dump(_plain(cfg), f)
        if cfg.main.wandb:
            try:
                import wandb
                if rank0:
                    wandb.init(name=cfg.main.experiment_name, project=cfg.env.env_config.env_name, config=_plain(cfg), tags=cfg.main.tags)
            except Exception as e:                    # offline image: keep training, losses stay on the device
                print(f"wandb unavailable ({e!r}); continuing with main.wandb=False")
                cfg.main.wandb = False
        models.append(run(cfg, cwd=args.cwd, results_dir=results_dir, device=torch.device(cfg.main.device), dp=dp))
        if cfg.main.wandb:
            import wandb
            wandb.finish()
    return models


def _plain(d):
    if isinstance(d, dict):
        return {k: _plain(v) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return [_plain(v) for v in d]
    return d


if __name__ == "__main__":
    main()
