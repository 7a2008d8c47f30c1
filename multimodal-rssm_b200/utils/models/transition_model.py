"""RSSM transition models — B200 host side.

Drop-in for the reference's ``utils/models/transition_model.py``: ``TransitionModel`` (single
modality) and ``MultimodalTransitionModel`` (K modalities, PoE / MoPoE fusion) keep the constructor
arguments, the ``forward`` signature, the 9-element (observe) / 4-element (imagine) return list, the
``.modules`` list, ``get_model_params`` order and the ``{'main', 'obs_encoder'}`` state-dict layout.

``forward`` is ONE kernel launch for all T steps (mrssm_rollout_fwd) instead of the reference's
Python loop over ~184 library ops and 7 host syncs per step; its backward is one BPTT
launch plus time-parallel weight-gradient GEMMs.  Noise comes from ``mrssm_b200.noise``.
"""
from typing import List, Optional

import torch
from torch import nn

from mrssm_b200 import noise, ops
from utils.models.encoder import MultimodalObsEncoder, ObsEncoder, StochasticStateModel, act_code


class _TransitionBase(nn.Module):
    __constants__ = ["min_std_dev"]

    def _build_core(self, belief_size, state_size, action_size, hidden_size, activation_function, min_std_dev):
        self.activation_function = activation_function
        self.act_fn = getattr(nn.functional, activation_function)
        self.min_std_dev = min_std_dev
        self.belief_size, self.state_size, self.action_size, self.hidden_size = (
            belief_size, state_size, action_size, hidden_size)
        self.fc_embed_state_action = nn.Linear(state_size + action_size, belief_size)
        self.rnn = nn.GRUCell(belief_size, belief_size)
        self.stochastic_state_model = StochasticStateModel(
            h_size=belief_size, s_size=state_size, hidden_size=hidden_size, activation=self.act_fn,
            min_std_dev=min_std_dev)

    # ordered list of expert heads: [(name, module, has_embedding)]
    def _experts(self):
        raise NotImplementedError

    def _fusion(self):
        raise NotImplementedError

    def _spec(self):
        experts = self._experts()
        key = (len(experts), self._fusion())
        if getattr(self, "_spec_cache", (None,))[0] != key:
            table = ops.FusionTable(len(experts), self.state_size, self._fusion())
            spec = ops.RolloutSpec(self.belief_size, self.state_size, self.hidden_size, self.action_size,
                                   act_code(self.activation_function), self.min_std_dev, table,
                                   [has for _, _, has in experts])
            self._spec_cache = (key, spec)
        return self._spec_cache[1]

    def _core_params(self):
        return [self.fc_embed_state_action.weight, self.fc_embed_state_action.bias, self.rnn.weight_ih,
                self.rnn.weight_hh, self.rnn.bias_ih, self.rnn.bias_hh]

    @staticmethod
    def _head_params(m):
        return [m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]

    def _rollout(self, prev_state, actions, prev_belief, embs, nonterminals, det):
        """embs: None (imagine) or list of per-expert embedding tensors [T,B,E] (None for experts
        without one).  Returns the flat tuple of RolloutFn."""
        observe = embs is not None
        spec = self._spec()
        T, B = actions.shape[0], actions.shape[1]
        dev = actions.device
        if nonterminals is not None:
            nonterminals = nonterminals.reshape(T, B)
        eps_prior = eps_post = None
        if not det:
            shape = (T, B, self.state_size)
            eps_prior = noise.draw("prior", shape, dev)
            if observe:
                eps_post = noise.draw("post", shape, dev)
        params = self._core_params() + self._head_params(self.stochastic_state_model)
        emb_args = []
        if observe:
            for (name, mod, has), e in zip(self._experts(), embs):
                params += self._head_params(mod)
                if has:
                    emb_args.append(e)
        return ops.RolloutFn.apply(spec, observe, bool(det), prev_state, actions, prev_belief, nonterminals,
                                   eps_prior, eps_post, *emb_args, *params)

    def get_state_dict(self):
        return {"main": self.state_dict(), "obs_encoder": self.obs_encoder.get_state_dict()}

    def _load_state_dict(self, state_dict):
        self.load_state_dict(state_dict["main"])
        self.obs_encoder._load_state_dict(state_dict["obs_encoder"])

    def get_model_params(self):
        return list(self.parameters()) + self.obs_encoder.get_model_params()

    def _train(self):
        self.train()
        self.obs_encoder.train()

    def _eval(self):
        self.eval()
        self.obs_encoder.eval()


class TransitionModel(_TransitionBase):
    """Single-modality RSSM transition (reference transition_model.py:10-136).  The posterior is the
    one ObsEncoder head; the last two returned entries are None."""

    def __init__(self, belief_size, state_size, action_size, hidden_size, embedding_size,
                 activation_function="relu", min_std_dev=0.1):
        super().__init__()
        self._build_core(belief_size, state_size, action_size, hidden_size, activation_function, min_std_dev)
        self.obs_encoder = ObsEncoder(h_size=belief_size, s_size=state_size, activation=self.act_fn,
                                      embedding_size=embedding_size["fusion"], hidden_size=hidden_size,
                                      min_std_dev=min_std_dev)
        self.modules = [self.fc_embed_state_action, self.stochastic_state_model, self.obs_encoder, self.rnn]

    def _experts(self):
        return [("obs", self.obs_encoder, True)]

    def _fusion(self):
        return "single"

    def forward(self, prev_state, actions, prev_belief, obs_emb: Optional[torch.Tensor] = None,
                nonterminals: Optional[torch.Tensor] = None, det=False) -> List[torch.Tensor]:
        out = self._rollout(prev_state, actions, prev_belief, None if obs_emb is None else [obs_emb], nonterminals, det)
        if obs_emb is None:
            return list(out[:4])
        return list(out[:7]) + [None, None]


class MultimodalTransitionModel(_TransitionBase):
    """K-modality transition with PoE / MoPoE posterior fusion (reference transition_model.py:139-307).
    Q1: the activation stays at its 'relu' default in every reference caller."""

    def __init__(self, belief_size, state_size, action_size, hidden_size, observation_names_enc, embedding_size,
                 activation_function="relu", min_std_dev=0.1, device=torch.device("cpu"), fusion_method="MoPoE",
                 expert_dist="q(st|ht,ot)"):
        super().__init__()
        self._build_core(belief_size, state_size, action_size, hidden_size, activation_function, min_std_dev)
        self.observation_names_enc = observation_names_enc
        self.modules = [self.fc_embed_state_action, self.stochastic_state_model, self.rnn]
        sizes = {}
        for name in observation_names_enc:
            kind = "image" if "image" in name else ("sound" if "sound" in name else "other")
            sizes[name] = embedding_size[kind]
        # a plain object held in a plain attribute: its heads are NOT sub-modules of this nn.Module,
        # as in the reference (they are moved / collected through obs_encoder's own methods)
        self.obs_encoder = MultimodalObsEncoder(
            expert_dist=expert_dist, h_size=belief_size, s_size=state_size, activation=self.act_fn,
            embedding_sizes=sizes, hidden_size=hidden_size, min_std_dev=min_std_dev, device=device)
        self.modules += self.obs_encoder.modules
        self.fusion_method = fusion_method

    def _experts(self):
        return [(n, m, n != "prior_expert") for n, m in self.obs_encoder.obs_encoder.items()]

    def _fusion(self):
        return "MoPoE" if self.fusion_method == "MoPoE" else "PoE"      # "NN" == PoE rollout (Q9)

    def forward(self, prev_state, actions, prev_belief, observations=None,
                nonterminals: Optional[torch.Tensor] = None, det=False) -> List[torch.Tensor]:
        if observations is None:
            return list(self._rollout(prev_state, actions, prev_belief, None, nonterminals, det))
        experts = self._experts()
        names = [n for n, _, _ in experts]
        missing = [n for n in names[1:] if n not in observations]
        if missing:
            raise KeyError(f"observations lack embeddings for {missing}")
        embs = [observations[n] if has else None for n, _, has in experts]
        out = self._rollout(prev_state, actions, prev_belief, embs, nonterminals, det)
        E = len(experts)
        means = dict(zip(names, out[7:7 + E]))
        stds = dict(zip(names, out[7 + E:7 + 2 * E]))
        return list(out[:7]) + [means, stds]
