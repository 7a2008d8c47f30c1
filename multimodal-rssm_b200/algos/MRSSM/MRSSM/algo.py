"""Factory (drop-in for the reference's ``algos/MRSSM/MRSSM/algo.py:6-18``)."""
from algos.MRSSM.MRSSM_MoPoE.algo import MRSSM_MoPoE
from algos.MRSSM.MRSSM_NN.algo import MRSSM_NN
from algos.MRSSM.MRSSM_PoE.algo import MRSSM_PoE
from algos.MRSSM.RSSM.algo import RSSM

_MULTIMODAL = {"NN": MRSSM_NN, "PoE": MRSSM_PoE, "MoPoE": MRSSM_MoPoE}


def build_RSSM(cfg, device):
    if not cfg.rssm.multimodal:
        return RSSM(cfg, device)
    try:
        cls = _MULTIMODAL[cfg.rssm.multimodal_params.fusion_method]
    except KeyError:
        raise NotImplementedError(cfg.rssm.multimodal_params.fusion_method)
    return cls(cfg, device)
