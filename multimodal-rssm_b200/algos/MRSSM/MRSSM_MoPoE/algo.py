"""MRSSM with Mixture-of-Products-of-Experts fusion (drop-in for the reference's
``algos/MRSSM/MRSSM_MoPoE/algo.py``).  Q4: the decoder latent is re-sampled from the re-fused experts;
Q6: the KL is the subset average with free nats per (t,b) and ignores kl_balancing_alpha."""
from algos.MRSSM.base.algo import MRSSM_base
from algos.MRSSM.base.builders import build_multimodal_models


class MRSSM_MoPoE(MRSSM_base):
    _kl_mode = 1
    _refuse = True

    def __init__(self, cfg, device):
        super().__init__(cfg, device)
        print("Multimodal RSSM (MoPoE)")

    def _init_models(self, device):
        build_multimodal_models(self, device)
