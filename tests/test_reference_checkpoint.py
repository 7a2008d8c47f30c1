"""CPU: a checkpoint WRITTEN BY THE UNMODIFIED REFERENCE (its own save_model after two optimisation steps, base/algo.py:55-58 —
nested module state dicts plus torch.optim.Adam's state_dict) loads into the product: parameters, Adam moments and step
count (`FusedClipAdam.load_state_dict`), and `load_model` (which, like the reference :51-54, rebuilds the optimiser).
The reference runs in its own process from the staged copy oracle/_ref (see __graft_entry__.stage_reference)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_arm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")
@pytest.mark.parametrize("fusion", ["MoPoE", "single"])
def test_reference_written_checkpoint_loads(tmp_path, fusion):
    from algos.MRSSM.MRSSM.algo import build_RSSM
    from mrssm_b200.config import hot_path_config
    out = subprocess.run([sys.executable, "-m", "oracle.ref_arm", "ckpt", "--out", str(tmp_path), "--fusion", fusion], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    ckpt = torch.load(info["file"], map_location="cpu")
    model = build_RSSM(hot_path_config(fusion=fusion, batch_size=2, chunk_size=4, device="cpu"), torch.device("cpu"))
    assert sum(p.numel() for p in model.param_list) == info["n_params"]
    if fusion == "single":
        # algos/MRSSM/RSSM keeps nn.Module's flat state_dict and saves no optimiser state (RSSM/algo.py get_state_dict)
        assert "model_optimizer" not in ckpt
        model.load_model(info["file"])
        mine = model.state_dict()
        assert set(mine) == set(ckpt)
        for k in ckpt:
            assert torch.equal(mine[k].cpu(), ckpt[k]), k
        return
    assert "model_optimizer" in ckpt and ckpt["model_optimizer"]["state"], "the reference saves Adam's state"
    model.load_state_dict(ckpt)
    # parameters: same nested keys, same tensors
    mine = model.get_state_dict()
    mine.pop("model_optimizer")
    theirs = {k: v for k, v in ckpt.items() if k != "model_optimizer"}

    def flat(d, pre=""):
        for k, v in d.items():
            if isinstance(v, dict):
                yield from flat(v, pre + k + "/")
            else:
                yield pre + k, v
    a, b = dict(flat(mine)), dict(flat(theirs))
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k].cpu(), b[k]), k
    # Adam state: param_list order is the reference's (transition, observation, reward, encoder), index i <-> parameter i
    st = ckpt["model_optimizer"]["state"]
    n_loaded = 0
    for i, p in enumerate(model.param_list):
        mst = model.model_optimizer.state[p]
        if i in st:
            assert torch.equal(mst["exp_avg"].cpu(), st[i]["exp_avg"]) and torch.equal(mst["exp_avg_sq"].cpu(), st[i]["exp_avg_sq"])
            n_loaded += 1
        else:                                    # reward head: the reference never stepped it (grad None)
            assert float(mst["exp_avg"].abs().max()) == 0.0
    assert n_loaded == len(st) > 0
    assert model.model_optimizer.step_count == 2
    assert model.model_optimizer.param_groups[0]["lr"] == ckpt["model_optimizer"]["param_groups"][0]["lr"]
    # load_model: same weights, fresh optimiser (the reference rebuilds Adam after loading)
    model2 = build_RSSM(hot_path_config(fusion=fusion, batch_size=2, chunk_size=4, device="cpu"), torch.device("cpu"))
    model2.load_model(info["file"])
    for p, q in zip(model.param_list, model2.param_list):
        assert torch.equal(p, q)
    assert model2.model_optimizer.step_count == 0 and float(model2.model_optimizer.flat_m.abs().max()) == 0.0


@pytest.mark.skipif(not ref_arm.available(), reason="oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")
def test_expert_dist_q_st_ot_is_unreachable_in_the_reference_and_rejected_here():
    """SURVEY §8f rank 4 lists expert_dist="q(st|ot)" as a variant.  The unmodified reference cannot construct it (its PoE / MoPoE
    factories pass `observation_names_enc=` to MultimodalStochasticEncoder, whose parameter is `observation_names_rec`: TypeError), so
    there is no behaviour to mirror; the product says so with NotImplementedError instead of inventing one."""
    from algos.MRSSM.MRSSM.algo import build_RSSM
    from mrssm_b200.config import hot_path_config
    out = subprocess.run([sys.executable, "-m", "oracle.ref_arm", "expert_dist"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert info["constructed"] is False and info["error"] == "TypeError" and "observation_names_enc" in info["message"], info
    ok = subprocess.run([sys.executable, "-m", "oracle.ref_arm", "expert_dist", "--expert-dist", "q(st|ht,ot)"], cwd=ROOT,
                        capture_output=True, text=True, timeout=600)
    assert json.loads([l for l in ok.stdout.splitlines() if l.startswith("{")][-1])["constructed"] is True
    cfg = hot_path_config(fusion="MoPoE", batch_size=2, chunk_size=4, device="cpu")
    cfg.rssm.multimodal_params.expert_dist = "q(st|ot)"
    with pytest.raises(NotImplementedError):
        build_RSSM(cfg, torch.device("cpu"))
