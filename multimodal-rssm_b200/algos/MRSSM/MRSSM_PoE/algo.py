"""MRSSM with Product-of-Experts fusion (drop-in for the reference's ``algos/MRSSM/MRSSM_PoE/algo.py``).
Balanced KL on the fused posterior; the decoder latent is a fresh sample of the re-fused experts (Q4)."""
from algos.MRSSM.base.algo import MRSSM_base
from algos.MRSSM.base.builders import build_multimodal_models


class MRSSM_PoE(MRSSM_base):
    _kl_mode = 0
    _refuse = True

    def __init__(self, cfg, device):
        super().__init__(cfg, device)
        print("Multimodal RSSM (PoE)")

    def _init_models(self, device):
        build_multimodal_models(self, device)
