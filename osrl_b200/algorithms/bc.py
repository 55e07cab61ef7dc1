"""BC / BCTrainer with the reference's signatures (osrl/algorithms/bc.py:26-109)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..common.net import MLPActor
from ._base import DummyLogger, EngineModel, EngineTrainer


class BC(EngineModel):
    algo = "bc"

    def __init__(self, state_dim: int, action_dim: int, max_action: float, a_hidden_sizes: list = [128, 128],
                 episode_len: int = 300, device: str = "cuda:0"):
        super().__init__()
        self.state_dim, self.action_dim, self.max_action = state_dim, action_dim, max_action
        self.a_hidden_sizes, self.episode_len, self.device = a_hidden_sizes, episode_len, device
        self.actor = MLPActor(state_dim, action_dim, a_hidden_sizes, nn.ReLU, max_action)

    def _hyper(self):
        return dict(state_dim=self.state_dim, action_dim=self.action_dim, max_action=self.max_action,
                    a_hidden_sizes=self.a_hidden_sizes)

    def setup_optimizers(self, actor_lr):
        self._lrs = dict(actor_lr=actor_lr)

    def act(self, obs):
        obs = torch.tensor(obs[None, ...], dtype=torch.float32, device=self.actor.pi[0].weight.device)
        return np.squeeze(self.actor(obs).data.cpu().numpy(), axis=0)


class BCTrainer(EngineTrainer):
    batch_keys = ("observations", "actions")

    def __init__(self, model: BC, env=None, logger=None, actor_lr: float = 1e-4, bc_mode: str = "all",
                 cost_limit: int = 10, device="cuda:0", **kw):
        super().__init__(model, env, logger, device=device, **kw)
        self.bc_mode, self.cost_limit = bc_mode, cost_limit
        self.model.setup_optimizers(actor_lr)
        self._lrs = self.model._lrs

    def set_target_cost(self, target_cost):
        self.cost_limit = target_cost

    def train_one_step(self, observations, actions):
        self._step({"observations": observations, "actions": actions})

    def evaluate(self, eval_episodes):
        """bc.py:111-124: unscaled episode means (BC has no reward / cost scale)."""
        self.model.eval()
        out = [self.rollout() for _ in range(eval_episodes)]
        self.model.train()
        return np.mean([o[0] for o in out]), np.mean([o[2] for o in out]), np.mean([o[1] for o in out])

    @torch.no_grad()
    def rollout(self):
        """bc.py:126-150; "multi-task" BC conditions the policy on the cost limit appended to the observation."""
        tag = (lambda o: np.append(o, self.cost_limit)) if self.bc_mode == "multi-task" else (lambda o: o)
        obs, info = self.env.reset()
        obs = tag(obs)
        ret, cost, n = 0.0, 0.0, 0
        for _ in range(self.model.episode_len):
            obs_next, reward, terminated, truncated, info = self.env.step(self.model.act(obs))
            obs = tag(obs_next)
            ret += reward
            n += 1
            cost += info["cost"]
            if terminated or truncated:
                break
        return ret, n, cost
