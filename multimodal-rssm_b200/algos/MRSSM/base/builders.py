"""Model construction shared by the three multimodal algorithm classes (the reference repeats this
block in MRSSM_{NN,PoE,MoPoE}/algo.py:_init_models)."""
from utils.models.encoder import MultimodalEncoder
from utils.models.observation_model import MultimodalObservationModel
from utils.models.reward_model import RewardModel
from utils.models.transition_model import MultimodalTransitionModel


def build_multimodal_models(algo, device):
    cfg = algo.cfg
    r = cfg.rssm
    if r.multimodal_params.expert_dist != "q(st|ht,ot)":
        raise NotImplementedError("expert_dist %r is outside the B200 hot path" % r.multimodal_params.expert_dist)
    emb = dict(r.embedding_size)
    acts = dict(r.activation_function)
    algo.transition_model = MultimodalTransitionModel(
        belief_size=r.belief_size, state_size=r.state_size, action_size=cfg.env.action_size,
        hidden_size=r.hidden_size, observation_names_enc=r.observation_names_enc, embedding_size=emb,
        device=device, fusion_method=r.multimodal_params.fusion_method,
        expert_dist=r.multimodal_params.expert_dist).to(device=device)
    algo.reward_model = RewardModel(h_size=r.belief_size, s_size=r.state_size, hidden_size=r.hidden_size,
                                    activation=r.activation_function.dense).to(device=device)
    algo.observation_model = MultimodalObservationModel(
        observation_names_rec=r.observation_names_rec, observation_shapes=cfg.env.observation_shapes,
        embedding_size=emb, belief_size=r.belief_size, state_size=r.state_size, hidden_size=r.hidden_size,
        activation_function=acts, normalization=r.normalization, device=device)
    algo.encoder = MultimodalEncoder(
        observation_names_enc=r.observation_names_enc, observation_shapes=cfg.env.observation_shapes,
        embedding_size=emb, activation_function=acts, normalization=r.normalization, device=device)
