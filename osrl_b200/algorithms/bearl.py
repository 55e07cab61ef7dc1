"""BEARL / BEARLTrainer with the reference's signatures (osrl/algorithms/bearl.py:46-412)."""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from ..common.net import VAE, EnsembleDoubleQCritic, LagrangianPIDController, SquashedGaussianMLPActor
from ._base import EngineModel, EngineTrainer


class BEARL(EngineModel):
    algo = "bearl"

    def __init__(self, state_dim: int, action_dim: int, max_action: float, a_hidden_sizes: list = [128, 128],
                 c_hidden_sizes: list = [128, 128], vae_hidden_sizes: int = 64, sample_action_num: int = 10,
                 gamma: float = 0.99, tau: float = 0.005, beta: float = 0.5, lmbda: float = 0.75, mmd_sigma: float = 50,
                 target_mmd_thresh: float = 0.05, num_samples_mmd_match: int = 10, PID: list = [0.1, 0.003, 0.001],
                 kernel: str = "gaussian", num_q: int = 1, num_qc: int = 1, cost_limit: int = 10,
                 episode_len: int = 300, start_update_policy_step: int = 20_000, device: str = "cuda:0"):
        super().__init__()
        self.state_dim, self.action_dim, self.latent_dim, self.max_action = state_dim, action_dim, action_dim * 2, max_action
        self.a_hidden_sizes, self.c_hidden_sizes, self.vae_hidden_sizes = a_hidden_sizes, c_hidden_sizes, vae_hidden_sizes
        self.sample_action_num, self.gamma, self.tau, self.beta, self.lmbda = sample_action_num, gamma, tau, beta, lmbda
        self.mmd_sigma, self.target_mmd_thresh, self.num_samples_mmd_match = mmd_sigma, target_mmd_thresh, num_samples_mmd_match
        self.start_update_policy_step = start_update_policy_step
        self.KP, self.KI, self.KD = PID
        self.kernel, self.num_q, self.num_qc = kernel, num_q, num_qc
        self.cost_limit, self.episode_len, self.device = cost_limit, episode_len, device
        # construction order of the reference (bearl.py:96-111)
        self.actor = SquashedGaussianMLPActor(state_dim, action_dim, a_hidden_sizes, nn.ReLU)
        self.critic = EnsembleDoubleQCritic(state_dim, action_dim, c_hidden_sizes, nn.ReLU, num_q=num_q)
        self.cost_critic = EnsembleDoubleQCritic(state_dim, action_dim, c_hidden_sizes, nn.ReLU, num_q=num_qc)
        self.vae = VAE(state_dim, action_dim, vae_hidden_sizes, self.latent_dim, max_action, device)
        self.actor_old = deepcopy(self.actor)
        self.critic_old = deepcopy(self.critic)
        self.cost_critic_old = deepcopy(self.cost_critic)
        self.qc_thres = cost_limit * (1 - gamma**episode_len) / (1 - gamma) / episode_len
        self.controller = LagrangianPIDController(self.KP, self.KI, self.KD, self.qc_thres)

    @property
    def log_alpha(self):
        e = self.engine
        return torch.tensor(e.scalars()["log_alpha"] if e else 0.0)

    @property
    def n_train_steps(self):
        e = self.engine
        return int(e.scalars()["n_train_steps"]) if e else 0

    def _hyper(self):
        return dict(state_dim=self.state_dim, action_dim=self.action_dim, max_action=self.max_action,
                    a_hidden_sizes=self.a_hidden_sizes, c_hidden_sizes=self.c_hidden_sizes,
                    vae_hidden_sizes=self.vae_hidden_sizes, sample_action_num=self.sample_action_num,
                    gamma=self.gamma, tau=self.tau, beta=self.beta, lmbda=self.lmbda, mmd_sigma=self.mmd_sigma,
                    target_mmd_thresh=self.target_mmd_thresh, num_samples_mmd_match=self.num_samples_mmd_match,
                    PID=[self.KP, self.KI, self.KD], kernel=self.kernel, num_q=self.num_q, num_qc=self.num_qc,
                    cost_limit=self.cost_limit, episode_len=self.episode_len,
                    start_update_policy_step=self.start_update_policy_step)

    def setup_optimizers(self, actor_lr, critic_lr, vae_lr, alpha_lr):
        self._lrs = dict(actor_lr=actor_lr, critic_lr=critic_lr, vae_lr=vae_lr, alpha_lr=alpha_lr)

    def act(self, obs, deterministic=False, with_logprob=False):
        dev = self.vae.d1.weight.device
        obs = torch.tensor(obs[None, ...], dtype=torch.float32, device=dev)
        a, logp = self.actor(obs, deterministic, with_logprob)
        logp = np.squeeze(logp.data.cpu().numpy()) if logp is not None else None
        return np.squeeze((a * self.max_action).data.cpu().numpy(), axis=0), logp


class BEARLTrainer(EngineTrainer):
    act_args = (True, True)   # deterministic roll-outs (bearl.py rollout)

    def __init__(self, model: BEARL, env=None, logger=None, actor_lr: float = 1e-3, critic_lr: float = 1e-3,
                 alpha_lr: float = 1e-3, vae_lr: float = 1e-3, reward_scale: float = 1.0, cost_scale: float = 1.0,
                 device="cuda:0", **kw):
        super().__init__(model, env, logger, reward_scale, cost_scale, device, **kw)
        self.model.setup_optimizers(actor_lr, critic_lr, vae_lr, alpha_lr)
        self._lrs = self.model._lrs

    def _torch_noise(self, eng):
        m = self.model
        B, S, L, a, N = eng.batch_size, m.sample_action_num, m.latent_dim, m.action_dim, m.num_samples_mmd_match
        return {"vae_eps": torch.randn(B, L), "pi_critic": torch.randn(B * S, a), "pi_cost": torch.randn(B * S, a),
                "z_mmd": torch.randn(B, N, L), "pi_actor": torch.randn(B * N, a)}

    def train_one_step(self, observations, next_observations, actions, rewards, costs, done):
        self._step({"observations": observations, "next_observations": next_observations, "actions": actions,
                    "rewards": rewards, "costs": costs, "done": done})
