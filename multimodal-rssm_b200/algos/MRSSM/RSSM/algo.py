"""Single-modality RSSM (drop-in for the reference's ``algos/MRSSM/RSSM/algo.py``): one encoder, one
decoder, TransitionModel with the configured dense activation, one nn.Module state_dict."""
import torch

from algos.MRSSM.base.algo import RSSM_base
from utils.models.encoder import bottle_tupele, build_Encoder
from utils.models.observation_model import build_ObservationModel
from utils.models.reward_model import RewardModel
from utils.models.transition_model import TransitionModel


class RSSM(RSSM_base):
    _kl_mode = 0
    _refuse = False

    def __init__(self, cfg, device):
        super().__init__(cfg, device)
        print("RSSM")

    def _init_models(self, device):
        cfg, r = self.cfg, self.cfg.rssm
        emb, acts = dict(r.embedding_size), dict(r.activation_function)
        self.transition_model = TransitionModel(r.belief_size, r.state_size, cfg.env.action_size, r.hidden_size,
                                                emb, r.activation_function.dense).to(device=device)
        self.reward_model = RewardModel(h_size=r.belief_size, s_size=r.state_size, hidden_size=r.hidden_size,
                                        activation=r.activation_function.dense).to(device=device)
        self.observation_model = build_ObservationModel(
            name=r.observation_names_rec[0], observation_shapes=cfg.env.observation_shapes, embedding_size=emb,
            belief_size=r.belief_size, state_size=r.state_size, hidden_size=r.hidden_size,
            activation_function=acts, normalization=r.normalization).to(device=device)
        self.encoder = build_Encoder(name=r.observation_names_enc[0], observation_shapes=cfg.env.observation_shapes,
                                     embedding_size=emb, activation_function=acts,
                                     normalization=r.normalization).to(device=device)

    def _init_param_list(self):
        self.param_list = list(self.parameters())

    def get_state_dict(self):
        return self.state_dict()

    def estimate_state(self, observations, actions, rewards, nonterminals, batch_size=None, det=False):
        if batch_size is None:
            batch_size = actions.shape[1]
        r = self.cfg.rssm
        init_belief = torch.zeros(batch_size, r.belief_size, device=self.cfg.main.device)
        init_state = torch.zeros(batch_size, r.state_size, device=self.cfg.main.device)
        obs_emb = bottle_tupele(self.encoder, observations)
        if self.dp is not None:
            self.dp.watch("transition", [obs_emb])
        out = self.transition_model(init_state, actions, init_belief, obs_emb, nonterminals, det=det)
        keys = ("beliefs", "prior_states", "prior_means", "prior_std_devs", "posterior_states", "posterior_means",
                "posterior_std_devs", "expert_means", "expert_std_devs")
        return dict(zip(keys, out))

    def _calc_observations_loss(self, observations_target, beliefs, posterior_states):
        o = observations_target[self.observation_name]
        mse = self.observation_model.mse_loss(beliefs, posterior_states, o)
        return {self.observation_name: self._obs_loss_from_mse(mse, o[0, 0].numel())}
