"""COptiDICE / COptiDICETrainer with the reference's signatures (osrl/algorithms/coptidice.py:41-287).

`update` (the whole train step: nu / chi / tau / lambda, then policy extraction) runs in the engine
(csrc/algo_coptidice.cu); `tau` and `lmbda` are not in the reference's state_dict (plain tensors, :104-105) and
live in the engine's device state -- read them through the properties below or `engine.scalars()`."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..common.net import EnsembleQCritic, SquashedGaussianMLPActor
from ._base import EngineModel, EngineTrainer


class COptiDICE(EngineModel):
    algo = "coptidice"

    def __init__(self, state_dim: int, action_dim: int, max_action: float, f_type: str, init_state_propotion: float,
                 observations_std: np.ndarray, actions_std: np.ndarray, a_hidden_sizes: list = [128, 128],
                 c_hidden_sizes: list = [128, 128], gamma: float = 0.99, alpha: float = 0.5,
                 cost_ub_epsilon: float = 0.01, num_nu: int = 1, num_chi: int = 1, cost_limit: int = 10,
                 episode_len: int = 300, device: str = "cuda:0"):
        super().__init__()
        self.state_dim, self.action_dim, self.max_action = state_dim, action_dim, max_action
        self.a_hidden_sizes, self.c_hidden_sizes = a_hidden_sizes, c_hidden_sizes
        self.gamma, self.alpha, self.cost_ub_epsilon = gamma, alpha, cost_ub_epsilon
        self.num_nu, self.num_chi, self.cost_limit, self.episode_len, self.device = num_nu, num_chi, cost_limit, episode_len, device
        self.f_type, self.init_state_propotion = f_type, float(init_state_propotion)
        self.qc_thres = cost_limit * (1 - gamma ** episode_len) / (1 - gamma) / episode_len
        # construction order of the reference (coptidice.py:106-119) => same init under the same seed
        self.actor = SquashedGaussianMLPActor(state_dim, action_dim, a_hidden_sizes, nn.ReLU)
        self.nu_network = EnsembleQCritic(state_dim, 0, c_hidden_sizes, nn.ReLU, num_q=num_nu)
        self.chi_network = EnsembleQCritic(state_dim, 0, c_hidden_sizes, nn.ReLU, num_q=num_chi)
        self.observations_std = np.asarray(observations_std, dtype=np.float32).reshape(-1)
        self.actions_std = np.asarray(actions_std, dtype=np.float32).reshape(-1)

    def _scalar(self, name):
        e = self.engine
        return torch.tensor([e.scalars()[name] if e else 1.0])

    @property
    def tau(self):
        return self._scalar("tau")

    @property
    def lmbda(self):
        return self._scalar("lmbda")

    def _hyper(self):
        return dict(state_dim=self.state_dim, action_dim=self.action_dim, max_action=self.max_action,
                    a_hidden_sizes=self.a_hidden_sizes, c_hidden_sizes=self.c_hidden_sizes, gamma=self.gamma,
                    f_type=self.f_type, init_state_propotion=self.init_state_propotion, alpha=self.alpha,
                    cost_ub_epsilon=self.cost_ub_epsilon, num_nu=self.num_nu, num_chi=self.num_chi,
                    cost_limit=self.cost_limit, episode_len=self.episode_len,
                    observations_std=self.observations_std, actions_std=self.actions_std)

    def setup_optimizers(self, actor_lr, critic_lr, scalar_lr):
        self._lrs = dict(actor_lr=actor_lr, critic_lr=critic_lr, scalar_lr=scalar_lr)

    def act(self, obs, deterministic=False, with_logprob=False):
        """coptidice.py:236-248: (tanh action, log-probability) of a single observation."""
        dev = self.actor.mu_layer.weight.device
        obs = torch.tensor(obs[None, ...], dtype=torch.float32, device=dev)
        a, logp = self.actor(obs, deterministic, with_logprob)
        logp = np.squeeze(logp.data.cpu().numpy()) if logp is not None else None
        return np.squeeze(a.data.cpu().numpy(), axis=0), logp


class COptiDICETrainer(EngineTrainer):
    batch_keys = ("observations", "next_observations", "actions", "rewards", "costs", "done", "is_init")

    def __init__(self, model: COptiDICE, env=None, logger=None, actor_lr: float = 1e-3, critic_lr: float = 1e-3,
                 scalar_lr: float = 1e-3, reward_scale: float = 1.0, cost_scale: float = 1.0, device="cuda:0", **kw):
        super().__init__(model, env, logger, reward_scale, cost_scale, device, **kw)
        self.model.setup_optimizers(actor_lr, critic_lr, scalar_lr)
        self._lrs = self.model._lrs

    def _torch_noise(self, eng):
        B, o, a = eng.batch_size, self.model.state_dim, self.model.action_dim
        nz = {"obs_eps": torch.randn(B, o), "act_eps": torch.randn(B, a)}   # coptidice.py:201-202
        torch.randn(B, a)                                                    # the actor's unused rsample (net.py:186)
        return nz

    def train_one_step(self, batch):
        """batch = (observations, next_observations, actions, rewards, costs, done, is_init) (coptidice.py:126-127)."""
        keys = ("observations", "next_observations", "actions", "rewards", "costs", "done", "is_init")
        self._step(dict(zip(keys, batch)))

    @torch.no_grad()
    def rollout(self):
        """coptidice.py:305-322: deterministic policy."""
        obs, info = self.env.reset()
        ret, cost, n = 0.0, 0.0, 0
        for _ in range(self.model.episode_len):
            act, _ = self.model.act(obs, True, True)
            obs, reward, terminated, truncated, info = self.env.step(act)
            ret += reward
            cost += info["cost"] * self.cost_scale
            n += 1
            if terminated or truncated:
                break
        return ret, n, cost
