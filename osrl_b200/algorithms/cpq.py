"""CPQ / CPQTrainer with the reference's signatures (osrl/algorithms/cpq.py:38-313)."""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from ..common.net import VAE, EnsembleQCritic, SquashedGaussianMLPActor
from ._base import EngineModel, EngineTrainer


class CPQ(EngineModel):
    algo = "cpq"

    def __init__(self, state_dim: int, action_dim: int, max_action: float, a_hidden_sizes: list = [128, 128],
                 c_hidden_sizes: list = [128, 128], vae_hidden_sizes: int = 64, sample_action_num: int = 10,
                 gamma: float = 0.99, tau: float = 0.005, beta: float = 1.5, num_q: int = 1, num_qc: int = 1,
                 qc_scalar: float = 1.5, cost_limit: int = 10, episode_len: int = 300, device: str = "cuda:0"):
        super().__init__()
        self.a_hidden_sizes, self.c_hidden_sizes, self.vae_hidden_sizes = a_hidden_sizes, c_hidden_sizes, vae_hidden_sizes
        self.gamma, self.tau, self.beta, self.cost_limit = gamma, tau, beta, cost_limit
        self.num_q, self.num_qc, self.qc_scalar, self.sample_action_num = num_q, num_qc, qc_scalar, sample_action_num
        self.state_dim, self.action_dim, self.latent_dim = state_dim, action_dim, action_dim * 2
        self.episode_len, self.max_action, self.device = episode_len, max_action, device
        # construction order of the reference (cpq.py:77-92)
        self.actor = SquashedGaussianMLPActor(state_dim, action_dim, a_hidden_sizes, nn.ReLU)
        self.critic = EnsembleQCritic(state_dim, action_dim, c_hidden_sizes, nn.ReLU, num_q=num_q)
        self.vae = VAE(state_dim, action_dim, vae_hidden_sizes, self.latent_dim, max_action, device)
        self.cost_critic = EnsembleQCritic(state_dim, action_dim, c_hidden_sizes, nn.ReLU, num_q=num_qc)
        self.actor_old = deepcopy(self.actor)
        self.critic_old = deepcopy(self.critic)
        self.cost_critic_old = deepcopy(self.cost_critic)
        self.q_thres = cost_limit * (1 - gamma**episode_len) / (1 - gamma) / episode_len
        self.qc_thres = qc_scalar * self.q_thres

    @property
    def log_alpha(self):
        """cpq.py:93 -- lives on the device inside the engine."""
        e = self.engine
        return torch.tensor(e.scalars()["log_alpha"] if e else 0.0)

    def _hyper(self):
        return dict(state_dim=self.state_dim, action_dim=self.action_dim, max_action=self.max_action,
                    a_hidden_sizes=self.a_hidden_sizes, c_hidden_sizes=self.c_hidden_sizes,
                    vae_hidden_sizes=self.vae_hidden_sizes, sample_action_num=self.sample_action_num,
                    gamma=self.gamma, tau=self.tau, beta=self.beta, num_q=self.num_q, num_qc=self.num_qc,
                    qc_scalar=self.qc_scalar, cost_limit=self.cost_limit, episode_len=self.episode_len)

    def setup_optimizers(self, actor_lr, critic_lr, alpha_lr, vae_lr):
        self._lrs = dict(actor_lr=actor_lr, critic_lr=critic_lr, alpha_lr=alpha_lr, vae_lr=vae_lr)

    def act(self, obs, deterministic=False, with_logprob=False):
        dev = self.vae.d1.weight.device
        obs = torch.tensor(obs[None, ...], dtype=torch.float32, device=dev)
        a, logp = self.actor(obs, deterministic, with_logprob)
        logp = np.squeeze(logp.data.cpu().numpy()) if logp is not None else None
        return np.squeeze((a * self.max_action).data.cpu().numpy(), axis=0), logp


class CPQTrainer(EngineTrainer):
    act_args = (True, True)   # deterministic roll-outs (cpq.py rollout)

    def __init__(self, model: CPQ, env=None, logger=None, actor_lr: float = 1e-4, critic_lr: float = 1e-4,
                 alpha_lr: float = 1e-4, vae_lr: float = 1e-4, reward_scale: float = 1.0, cost_scale: float = 1.0,
                 device="cuda:0", **kw):
        super().__init__(model, env, logger, reward_scale, cost_scale, device, **kw)
        self.model.setup_optimizers(actor_lr, critic_lr, alpha_lr, vae_lr)
        self._lrs = self.model._lrs

    def _torch_noise(self, eng):
        m = self.model
        B, S, L, a = eng.batch_size, m.sample_action_num, m.latent_dim, m.action_dim
        nz = {"vae_eps": torch.randn(B, L), "pi_critic": torch.randn(B, a), "pi_cost": torch.randn(B, a)}
        torch.randn(B, a)                       # cpq.py:164: rsample consumed, result unused
        nz["ood_sample"] = torch.randn(S, B, a)
        torch.randn(S * B, L)                   # cpq.py:178: randn_like inside VAE.forward, unused
        nz["pi_actor"] = torch.randn(B, a)
        return nz

    def train_one_step(self, observations, next_observations, actions, rewards, costs, done):
        self._step({"observations": observations, "next_observations": next_observations, "actions": actions,
                    "rewards": rewards, "costs": costs, "done": done})
