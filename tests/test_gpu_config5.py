"""-m gpu: BASELINE config 5 — the scaled model (deter = hidden = 1024, stoch = 64, 128x128 images: `ImageEncoder_128`
encoder.py:415-509, `ImageDecoder_128` observation_model.py:162-229, the rollout of transition_model.py:200-285 at D = 1024) —
at a reduced batch, one full train step against the CPU oracle: exact fp32 mode at the north star's rtol 1e-3, bf16
tensor-core mode within the stated bf16 tolerances.  (At D = H = 1024 the rollout runs on the L2-streaming fp32 kernels in both
modes: the tcgen05 rollout holds D, H <= 208; the conv stacks, dense layers and MLPs are on the tensor cores in bf16 mode.)"""
import pytest

from tests import parity_util as U

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG5 = dict(belief_size=1024, hidden_size=1024, state_size=64,
            names_enc=("image_horizon_128", "pose_quat_v2"), names_rec=("image_horizon_128", "pose_quat_v2"),
            observation_shapes={"image_horizon_128": [3, 128, 128], "pose_quat_v2": [3]})


def test_config5_train_step_fp32_matches_oracle():
    out = U.run_train_parity("MoPoE", B=2, T=4, steps=2, device=DEV, rtol=1e-3, **CFG5)
    assert out["worst_grad_err"] < 2e-3, out


def test_config5_train_step_bf16_within_stated_tolerance():
    rep = U.run_train_parity_bf16("MoPoE", B=3, T=5, steps=2, device=DEV, **CFG5)
    assert rep["state_err"] < 2e-2 and rep["loss_rel"] < 2e-2 and rep["gnorm_rel"] < 3e-2 and rep["grad_rel_fro"] < 5e-2, rep


def test_config5_single_modal_rssm_fp32():
    kw = dict(CFG5, names_enc=("image_horizon_128",), names_rec=("image_horizon_128",))
    out = U.run_train_parity("single", B=2, T=4, steps=1, device=DEV, rtol=1e-3, **kw)
    assert out["worst_grad_err"] < 2e-3, out
