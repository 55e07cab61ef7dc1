"""BCQL / BCQLTrainer with the reference's signatures (osrl/algorithms/bcql.py:44-306)."""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from ..common.net import VAE, EnsembleDoubleQCritic, LagrangianPIDController, MLPGaussianPerturbationActor
from ._base import EngineModel, EngineTrainer


class BCQL(EngineModel):
    algo = "bcql"

    def __init__(self, state_dim: int, action_dim: int, max_action: float, a_hidden_sizes: list = [128, 128],
                 c_hidden_sizes: list = [128, 128], vae_hidden_sizes: int = 64, sample_action_num: int = 10,
                 gamma: float = 0.99, tau: float = 0.005, phi: float = 0.05, lmbda: float = 0.75, beta: float = 0.5,
                 PID: list = [0.1, 0.003, 0.001], num_q: int = 1, num_qc: int = 1, cost_limit: int = 10,
                 episode_len: int = 300, device: str = "cuda:0"):
        super().__init__()
        self.state_dim, self.action_dim, self.max_action = state_dim, action_dim, max_action
        self.latent_dim = action_dim * 2
        self.a_hidden_sizes, self.c_hidden_sizes, self.vae_hidden_sizes = a_hidden_sizes, c_hidden_sizes, vae_hidden_sizes
        self.sample_action_num, self.gamma, self.tau, self.phi = sample_action_num, gamma, tau, phi
        self.lmbda, self.beta = lmbda, beta
        self.KP, self.KI, self.KD = PID
        self.num_q, self.num_qc, self.cost_limit, self.episode_len, self.device = num_q, num_qc, cost_limit, episode_len, device
        # same construction order as the reference (bcql.py:85-98) => same init under the same seed
        self.actor = MLPGaussianPerturbationActor(state_dim, action_dim, a_hidden_sizes, nn.Tanh, phi, max_action)
        self.critic = EnsembleDoubleQCritic(state_dim, action_dim, c_hidden_sizes, nn.ReLU, num_q=num_q)
        self.cost_critic = EnsembleDoubleQCritic(state_dim, action_dim, c_hidden_sizes, nn.ReLU, num_q=num_qc)
        self.vae = VAE(state_dim, action_dim, vae_hidden_sizes, self.latent_dim, max_action, device)
        self.actor_old = deepcopy(self.actor)
        self.critic_old = deepcopy(self.critic)
        self.cost_critic_old = deepcopy(self.cost_critic)
        self.qc_thres = cost_limit * (1 - gamma**episode_len) / (1 - gamma) / episode_len
        self.controller = LagrangianPIDController(self.KP, self.KI, self.KD, self.qc_thres)

    def _hyper(self):
        return dict(state_dim=self.state_dim, action_dim=self.action_dim, max_action=self.max_action,
                    a_hidden_sizes=self.a_hidden_sizes, c_hidden_sizes=self.c_hidden_sizes,
                    vae_hidden_sizes=self.vae_hidden_sizes, sample_action_num=self.sample_action_num,
                    gamma=self.gamma, tau=self.tau, phi=self.phi, lmbda=self.lmbda, beta=self.beta,
                    PID=[self.KP, self.KI, self.KD], num_q=self.num_q, num_qc=self.num_qc,
                    cost_limit=self.cost_limit, episode_len=self.episode_len)

    def setup_optimizers(self, actor_lr, critic_lr, vae_lr):
        self._lrs = dict(actor_lr=actor_lr, critic_lr=critic_lr, vae_lr=vae_lr)

    def act(self, obs, deterministic=False, with_logprob=False):
        dev = self.vae.d1.weight.device
        obs = torch.tensor(obs[None, ...], dtype=torch.float32, device=dev)
        act = self.actor(obs, self.vae.decode(obs))
        return np.squeeze(act.data.cpu().numpy(), axis=0), None


class BCQLTrainer(EngineTrainer):
    def __init__(self, model: BCQL, env=None, logger=None, actor_lr: float = 1e-4, critic_lr: float = 1e-4,
                 vae_lr: float = 1e-4, reward_scale: float = 1.0, cost_scale: float = 1.0, device="cuda:0", **kw):
        super().__init__(model, env, logger, reward_scale, cost_scale, device, **kw)
        self.model.setup_optimizers(actor_lr, critic_lr, vae_lr)
        self._lrs = self.model._lrs

    def _torch_noise(self, eng):
        B, S, L = eng.batch_size, self.model.sample_action_num, self.model.latent_dim
        # draw order of the reference (net.py:327 then net.py:334 x3)
        return {"vae_eps": torch.randn(B, L), "z_critic": torch.randn(B * S, L), "z_cost": torch.randn(B * S, L),
                "z_actor": torch.randn(B, L)}

    def train_one_step(self, observations, next_observations, actions, rewards, costs, done):
        self._step({"observations": observations, "next_observations": next_observations, "actions": actions,
                    "rewards": rewards, "costs": costs, "done": done})
