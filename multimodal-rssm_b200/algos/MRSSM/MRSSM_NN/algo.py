"""MRSSM "NN" variant (drop-in for the reference's ``algos/MRSSM/MRSSM_NN/algo.py``).  Q9: it is the
PoE rollout with the base loss — balanced KL and the rollout's own posterior sample to the decoders."""
from algos.MRSSM.base.algo import MRSSM_base
from algos.MRSSM.base.builders import build_multimodal_models


class MRSSM_NN(MRSSM_base):
    _kl_mode = 0
    _refuse = False

    def __init__(self, cfg, device):
        super().__init__(cfg, device)
        print("Multimodal RSSM (NN)")

    def _init_models(self, device):
        build_multimodal_models(self, device)
