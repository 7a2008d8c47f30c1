"""Algorithm layer of the MRSSM hot path (B200 host side).

Drop-in for the reference's ``algos/MRSSM/base/algo.py``: ``RSSM_base`` / ``MRSSM_base`` keep the
public surface the driver and the evaluation tools use — ``optimize(D)``, ``validation(D)``,
``estimate_state(...)``, ``save_model`` / ``load_model``, ``eval()`` / ``train()``, ``_clip_obs``, and
the attributes ``param_list``, ``model_optimizer``, ``transition_model``, ``encoder``,
``observation_model``, ``reward_model``, ``cfg``, ``itr_optim`` — and the checkpoint layout.

What differs is how a step executes (reference call stack: SURVEY §3.1):
  * encoders / rollout / decoders / ELBO are a few dozen launches of libmrssm_b200.so kernels, none of
    them per time-step; no host sync happens inside a step (losses stay on the device; ``loss_info``
    holds 0-dim tensors, converted to floats only when wandb logging is on);
  * gradients land in one flat buffer; global-norm clip + Adam is one fused call;
  * under data parallelism (``mrssm_b200.dist``) the flat gradient is all-reduced over NCCL.
"""
import os

import torch
from torch import nn

from mrssm_b200 import noise, ops
from mrssm_b200.optim import FusedClipAdam
from utils.models.encoder import bottle_tupele_multimodal

try:                                   # optional, exactly as unused as in the reference when main.wandb is False
    import wandb
except Exception:                      # pragma: no cover
    wandb = None

_LOG_SQRT_2PI = 0.9189385332046727


class RSSM_base(nn.Module):
    # how the latent half of the ELBO is evaluated; subclasses override
    _kl_mode = 0          # 0: balanced KL on the rollout posterior, 1: MoPoE subset-averaged KL
    _refuse = False       # re-fuse experts and draw a fresh decoder latent (PoE / MoPoE)

    def __init__(self, cfg, device):
        super().__init__()
        self.cfg = cfg
        self.device = device
        self.observation_name = cfg.rssm.observation_names_enc[0]
        self._init_models(device)
        self._init_param_list()
        self._init_optimizer()
        self.free_nats = torch.full((1,), float(cfg.rssm.free_nats), device=device)
        self.model_modules = (self.transition_model.modules + self.encoder.modules
                              + self.observation_model.modules + self.reward_model.modules)
        self.itr_optim = 0
        self.dp = None                 # mrssm_b200.dist.DataParallel, attached by the launcher
        self.loss_info = {}

    # ---- construction hooks ---------------------------------------------------------------------
    def _init_models(self, device):
        raise NotImplementedError

    def _init_param_list(self):
        raise NotImplementedError

    def _init_optimizer(self):
        r = self.cfg.rssm
        lr0 = 0 if r.learning_rate_schedule != 0 else r.model_learning_rate
        self.model_optimizer = FusedClipAdam(self.param_list, lr=lr0, eps=r.adam_epsilon, max_grad_norm=r.grad_clip_norm)
        ops.bump_weight_version()      # new or re-loaded parameters: no cached bf16 packing may match them by address
        if getattr(self, "dp", None) is not None:     # rebuilt after load_model under data parallelism: re-bind the exchange
            self.dp.attach(self.model_optimizer)

    def get_state_dict(self):
        raise NotImplementedError

    # ---- checkpoints (reference base/algo.py:47-58) ------------------------------------------------
    def _load_model_dicts(self, model_path):
        print("load model_dicts from {}".format(model_path))
        return torch.load(model_path, map_location=torch.device(self.device))

    def load_model(self, model_path):
        self.load_state_dict(self._load_model_dicts(model_path))
        self._init_optimizer()         # the reference rebuilds Adam after loading, dropping the moments

    def save_model(self, results_dir, itr):
        """Every rank holds the same weights under data parallelism: rank 0 writes, everyone waits for the file."""
        import torch.distributed as dist
        multi = self.dp is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if not multi or dist.get_rank() == 0:
            torch.save(self.get_state_dict(), os.path.join(results_dir, "models_%d.pth" % itr))
        if multi:
            dist.barrier()

    def _clip_obs(self, observations, idx_start=0, idx_end=None):
        return {k: v[idx_start:idx_end] for k, v in observations.items()}

    def estimate_state(self, observations, actions, rewards, nonterminals, batch_size=None, det=False):
        raise NotImplementedError

    # ---- ELBO --------------------------------------------------------------------------------------
    def _latent_spec(self):
        tm = self.transition_model
        table = tm._spec().table
        key = (self._kl_mode, self._refuse)
        if getattr(self, "_lspec", (None,))[0] != key:
            r = self.cfg.rssm
            self._lspec = (key, ops.LatentSpec(r.state_size, table, self._kl_mode, self._refuse, r.free_nats,
                                               r.kl_balancing_alpha))
        return self._lspec[1]

    def _latent_terms(self, states):
        """-> (decoder latent, q mean, q std, kl_loss, global KL) in one fused launch."""
        spec = self._latent_spec()
        pm, ps = states["prior_means"], states["prior_std_devs"]
        qm, qs = states["posterior_means"], states["posterior_std_devs"]
        experts = []
        if self._refuse or self._kl_mode == 1:
            names = list(states["expert_means"].keys())
            experts = [states["expert_means"][n] for n in names] + [states["expert_std_devs"][n] for n in names]
        if self._refuse:
            eps = noise.draw("dec", pm.shape, pm.device)
            z, fm, fs, sums = ops.LatentFn.apply(spec, pm, ps, qm, qs, eps, *experts)
        else:
            (sums,) = ops.LatentFn.apply(spec, pm, ps, qm, qs, None, *experts)
            z, fm, fs = states["posterior_states"], qm, qs
        return z, fm, fs, sums[0], sums[1]

    def _get_posterior_states(self, states):
        z, fm, fs, _, _ = self._latent_terms(states)
        return z, fm, fs

    def _calc_kl(self, states):
        return self._latent_terms(states)[3]

    def _overshooting_targets(self, states):
        """-> list of (n_experts, expert mask, target tensors): one open-loop run per entry.  RSSM / NN: the rollout's
        posterior (reference base/algo.py:118); PoE: the product of all experts (MRSSM_PoE _get_posterior_states);
        MoPoE: one run per modality subset (MRSSM_MoPoE/algo.py:78-80)."""
        if self._refuse or self._kl_mode == 1:
            table = self.transition_model._spec().table
            names = list(states["expert_means"].keys())
            experts = [states["expert_means"][n] for n in names] + [states["expert_std_devs"][n] for n in names]
            masks = table.masks if self._kl_mode == 1 else [(1 << table.n_experts) - 1]
            return [(table.n_experts, m, experts) for m in masks]
        return [(0, 0, [states["posterior_means"], states["posterior_std_devs"]])]

    def _latent_overshooting(self, actions, rewards, nonterminals, states):
        """Reference base/algo.py:111-148 (MoPoE: MRSSM_MoPoE/algo.py:69-108).  actions / rewards / nonterminals are the
        full chunk.  Every start step's open-loop run is one column block of a single imagination rollout."""
        r = self.cfg.rssm
        T, B, OD = actions.shape[0], actions.shape[1], r.overshooting_distance
        beliefs, prior_states = states["beliefs"], states["prior_states"]
        targets = self._overshooting_targets(states)
        want_reward = r.overshooting_reward_scale != 0 and r.predict_reward    # zeroed otherwise (reference :200-201)
        gspec = ops.OvershootSpec(T, B, r.state_size, actions.shape[2], OD, r.free_nats)
        act_o, nt_o, rw_o, mask_o = ops.overshoot_gather(gspec, actions, nonterminals, rewards if want_reward else None,
                                                         want_mask=want_reward)
        h0 = beliefs[:T - 2].reshape(gspec.N, -1)            # beliefs[t-1] of every start t = 1..T-2, side by side
        s0 = prior_states[:T - 2].reshape(gspec.N, -1)
        kl_sum = None
        for n_experts, mask, tg in targets:
            o_beliefs, o_states, o_means, o_stds = self.transition_model(s0, act_o, h0, None, nt_o)[:4]
            spec = ops.OvershootSpec(T, B, r.state_size, actions.shape[2], OD, r.free_nats,
                                     scale=r.overshooting_kl_beta / len(targets), n_experts=n_experts, mask=mask)
            kl = ops.OvershootKlFn.apply(spec, o_means, o_stds, *tg)
            kl_sum = kl if kl_sum is None else kl_sum + kl
        reward_loss = torch.zeros((), device=kl_sum.device)
        if want_reward:                                       # on the last run, as the reference does
            pred = self.reward_model(h_t=o_beliefs, s_t=o_states)["loc"] * mask_o
            rows = pred.numel()
            mse = ops.MseLossFn.apply(pred.reshape(rows, 1), rw_o.reshape(rows, 1), rows)
            reward_loss = mse * ((1.0 / OD) * r.overshooting_reward_scale * (T - 1))
        return kl_sum, reward_loss

    def _calc_observations_loss(self, observations_target, beliefs, posterior_states):
        raise NotImplementedError

    def _obs_loss_from_mse(self, mse_sum, n_features):
        if self.cfg.rssm.worldmodel_LogProbLoss:          # -log N(o; loc, 1) summed over features
            return 0.5 * mse_sum + n_features * _LOG_SQRT_2PI
        return mse_sum

    def _calc_loss(self, observations_target, actions, rewards, nonterminals, states):
        z, _, _, kl_loss, kl_global = self._latent_terms(states)
        if self.dp is not None:        # decoder gradients are complete once the gradient of the decoders' input latent exists
            self.dp.watch("decoder", [z, states["beliefs"]])
        observations_loss = self._calc_observations_loss(observations_target, states["beliefs"], z)
        kl_loss_sum = kl_loss
        if self.cfg.rssm.global_kl_beta != 0:
            kl_loss_sum = kl_loss + self.cfg.rssm.global_kl_beta * kl_global
        if self.cfg.rssm.predict_reward:                        # reference _calc_reward_loss :96-109, :175
            r = self.reward_model(h_t=states["beliefs"], s_t=z)["loc"]
            rows = r.numel()
            mse = ops.MseLossFn.apply(r.reshape(rows, 1), rewards[:-1].reshape(rows, 1).contiguous(), rows)
            reward_loss = self._obs_loss_from_mse(mse, 1)       # MSE, or -log N(r; loc, 1) with worldmodel_LogProbLoss
        else:                                                   # the reference zeroes it (:200-201): forward skipped
            reward_loss = torch.zeros((), device=kl_loss.device)
        if self.cfg.rssm.overshooting_kl_beta != 0:             # reference :190-193
            kl_over, reward_over = self._latent_overshooting(actions, rewards, nonterminals, states)
            kl_loss_sum = kl_loss_sum + kl_over
            reward_loss = reward_loss + reward_over
        return observations_loss, reward_loss, kl_loss_sum, kl_loss

    def _get_model_loss(self, observations_target, actions, rewards, nonterminals, states):
        observations_loss, reward_loss, kl_loss_sum, kl_loss = self._calc_loss(
            observations_target, actions, rewards, nonterminals, states)
        observations_loss_sum = sum(observations_loss.values())
        model_loss = observations_loss_sum + reward_loss + self.cfg.rssm.kl_beta * kl_loss_sum
        info = {"observations_loss_sum": observations_loss_sum.detach()}
        for name, v in observations_loss.items():
            info["observation_{}_loss".format(name)] = v.detach()
        info["reward_loss"] = reward_loss.detach()
        info["kl_loss_sum"] = kl_loss_sum.detach()
        info["kl_loss"] = kl_loss.detach()
        return model_loss, info

    # ---- one optimisation step (reference base/algo.py:234-276) ----------------------------------------
    def _sample_data(self, D):
        observations, actions, rewards, nonterminals = D.sample(self.cfg.train.batch_size, self.cfg.train.chunk_size)
        return self._clip_obs(observations, idx_start=1), actions, rewards, nonterminals

    def _ramp_lr(self):
        r = self.cfg.rssm
        if r.learning_rate_schedule != 0:
            for group in self.model_optimizer.param_groups:
                group["lr"] = min(group["lr"] + r.model_learning_rate / r.learning_rate_schedule, r.model_learning_rate)

    def _log(self, info, suffix, step):
        if self.cfg.main.wandb and wandb is not None:
            vals = torch.stack([torch.as_tensor(v, device=self.device).float() for v in info.values()]).tolist()
            for name, v in zip(info.keys(), vals):
                wandb.log(data={"{}/{}".format(name, suffix): v}, step=step)

    def optimize_loss(self, observations_target, actions, rewards, nonterminals, states, itr_optim):
        self.model_optimizer.zero_grad()            # kernels accumulate straight into the flat grad buffer
        model_loss, info = self._get_model_loss(observations_target, actions, rewards, nonterminals, states)
        with ops.side_wgrad_scope():                # decoder weight gradients may run beside the rollout BPTT; joined on exit
            model_loss.backward()
        if self.dp is not None:
            self.dp.all_reduce_grads(self.model_optimizer)
        self._ramp_lr()
        self.model_optimizer.step()                 # fused global-norm clip + Adam
        self.loss_info = info
        self.model_loss = model_loss.detach()
        self._log(info, "train", itr_optim)
        if self.cfg.main.wandb and wandb is not None:
            wandb.log(data={"frame": itr_optim * self.cfg.train.batch_size * self.cfg.train.chunk_size}, step=itr_optim)

    def optimize(self, D):
        self.itr_optim += 1
        ops.set_bf16_mode(bool(self.cfg.train.use_amp))      # use_amp selects the bf16 tensor-core kernels (Q11)
        ops.forget_s2d()
        observations_target, actions, rewards, nonterminals = self._sample_data(D)
        # gradients are zeroed inside optimize_loss *before* any backward work; the forward below only
        # builds the graph
        states = self.estimate_state(observations_target, actions[:-1], rewards, nonterminals[:-1])
        self.optimize_loss(observations_target, actions, rewards, nonterminals, states, self.itr_optim)

    def validation(self, D):
        self.eval()
        ops.set_bf16_mode(bool(self.cfg.train.use_amp))
        ops.forget_s2d()
        with torch.no_grad():
            observations_target, actions, rewards, nonterminals = self._sample_data(D)
            states = self.estimate_state(observations_target, actions[:-1], rewards, nonterminals[:-1])
            _, info = self._get_model_loss(observations_target, actions, rewards, nonterminals, states)
        self._ramp_lr()        # the reference advances its LR ramp inside _calc_loss (base/algo.py:195-198), validation included
        self.validation_info = info
        self._log(info, "validation", self.itr_optim)
        self.train()


class MRSSM_base(RSSM_base):
    """Multimodal flavour: plain-object containers for encoder / decoder (reference base/algo.py:295-385)."""

    def eval(self):
        self.transition_model._eval()
        self.observation_model.eval()
        self.reward_model.eval()
        self.encoder.eval()

    def train(self, mode=True):
        self.transition_model._train()
        self.observation_model.train()
        self.reward_model.train()
        self.encoder.train()

    def load_state_dict(self, model_dicts):
        self.observation_model._load_state_dict(model_dicts["observation_model"])
        self.encoder._load_state_dict(model_dicts["encoder"])
        self.transition_model._load_state_dict(model_dicts["transition_model"])
        self.reward_model.load_state_dict(model_dicts["reward_model"])
        if "model_optimizer" in model_dicts:
            self.model_optimizer.load_state_dict(model_dicts["model_optimizer"])

    def _init_param_list(self):
        self.param_list = (self.transition_model.get_model_params() + self.observation_model.get_model_params()
                           + list(self.reward_model.parameters()) + self.encoder.get_model_params())

    def get_state_dict(self):
        return {"transition_model": self.transition_model.get_state_dict(),
                "observation_model": self.observation_model.get_state_dict(),
                "reward_model": self.reward_model.state_dict(),
                "encoder": self.encoder.get_state_dict(),
                "model_optimizer": self.model_optimizer.state_dict()}

    def estimate_state(self, observations, actions, rewards, nonterminals, batch_size=None, det=False):
        if batch_size is None:
            batch_size = actions.shape[1]
        r = self.cfg.rssm
        init_belief = torch.zeros(batch_size, r.belief_size, device=self.cfg.main.device)
        init_state = torch.zeros(batch_size, r.state_size, device=self.cfg.main.device)
        obs_emb = bottle_tupele_multimodal(self.encoder, observations)
        if self.dp is not None:        # transition gradients are complete once the gradient of the embeddings exists
            self.dp.watch("transition", list(obs_emb.values()))
        out = self.transition_model(init_state, actions, init_belief, obs_emb, nonterminals, det=det)
        keys = ("beliefs", "prior_states", "prior_means", "prior_std_devs", "posterior_states", "posterior_means",
                "posterior_std_devs", "expert_means", "expert_std_devs")
        return dict(zip(keys, out))

    def _calc_observations_loss(self, observations_target, beliefs, posterior_states):
        mse = self.observation_model.mse_loss(h_t=beliefs, s_t=posterior_states, o_t=observations_target)
        return {name: self._obs_loss_from_mse(v, observations_target[name][0, 0].numel()) for name, v in mse.items()}
