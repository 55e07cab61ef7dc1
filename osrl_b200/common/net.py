"""Parameter shells with the reference's class names, constructor signatures and state_dict
keys (osrl/common/net.py).  They construct ``nn.Linear`` modules in the reference's order,
so ``seed_all(s)`` yields the same initial values; once a trainer binds the model to an
engine, every parameter's storage is a view into the engine's HBM arena and all training
arithmetic happens in libosrl_b200.so.  The ``forward`` methods here are only used by
``act()`` at evaluation time (batch 1, outside the hot path).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def mlp(sizes, activation, output_activation=nn.Identity):
    """Same module layout as the reference's mlp(): Linear at even indices (net.py:12-30)."""
    layers = []
    for j in range(len(sizes) - 1):
        act = activation if j < len(sizes) - 2 else output_activation
        layers += [nn.Linear(sizes[j], sizes[j + 1]), act()]
    return nn.Sequential(*layers)


class MLPActor(nn.Module):
    """net.py:65-85"""

    def __init__(self, obs_dim, act_dim, hidden_sizes, activation, act_limit=1):
        super().__init__()
        self.pi = mlp([obs_dim] + list(hidden_sizes) + [act_dim], activation, nn.Tanh)
        self.act_limit = act_limit

    def forward(self, obs):
        return self.act_limit * self.pi(obs)


class MLPGaussianPerturbationActor(nn.Module):
    """net.py:33-62"""

    def __init__(self, obs_dim, act_dim, hidden_sizes, activation, phi=0.05, act_limit=1):
        super().__init__()
        self.pi = mlp([obs_dim + act_dim] + list(hidden_sizes) + [act_dim], activation, nn.Tanh)
        self.act_limit = act_limit
        self.phi = phi

    def forward(self, obs, act):
        a = self.phi * self.act_limit * self.pi(torch.cat([obs, act], 1))
        return (a + act).clamp(-self.act_limit, self.act_limit)


class SquashedGaussianMLPActor(nn.Module):
    """net.py:152-205 (evaluation path: deterministic or sampled action, no log-prob)."""

    def __init__(self, obs_dim, act_dim, hidden_sizes, activation):
        super().__init__()
        self.net = mlp([obs_dim] + list(hidden_sizes), activation, activation)
        self.mu_layer = nn.Linear(hidden_sizes[-1], act_dim)
        self.log_std_layer = nn.Linear(hidden_sizes[-1], act_dim)

    def forward(self, obs, deterministic=False, with_logprob=True, **_):
        """Evaluation path (net.py:170-205): tanh action and, on request, its log-density with the tanh correction."""
        h = self.net(obs)
        mu = self.mu_layer(h)
        std = torch.exp(torch.clamp(self.log_std_layer(h), -20, 2))
        u = mu if deterministic else mu + std * torch.randn_like(std)
        logp = None
        if with_logprob:
            logp = torch.distributions.Normal(mu, std).log_prob(u).sum(-1)
            logp = logp - (2 * (np.log(2) - u - F.softplus(-2 * u))).sum(1)
        return torch.tanh(u), logp


class EnsembleQCritic(nn.Module):
    """net.py:208-242"""

    def __init__(self, obs_dim, act_dim, hidden_sizes, activation, num_q=2):
        super().__init__()
        assert num_q >= 1, "num_q param should be greater than 1"
        self.q_nets = nn.ModuleList([mlp([obs_dim + act_dim] + list(hidden_sizes) + [1], nn.ReLU) for _ in range(num_q)])


class EnsembleDoubleQCritic(nn.Module):
    """net.py:245-287"""

    def __init__(self, obs_dim, act_dim, hidden_sizes, activation, num_q=2):
        super().__init__()
        assert num_q >= 1, "num_q param should be greater than 1"
        self.q1_nets = nn.ModuleList([mlp([obs_dim + act_dim] + list(hidden_sizes) + [1], nn.ReLU) for _ in range(num_q)])
        self.q2_nets = nn.ModuleList([mlp([obs_dim + act_dim] + list(hidden_sizes) + [1], nn.ReLU) for _ in range(num_q)])


class VAE(nn.Module):
    """net.py:290-353"""

    def __init__(self, obs_dim, act_dim, hidden_size, latent_dim, act_lim, device="cpu"):
        super().__init__()
        self.e1 = nn.Linear(obs_dim + act_dim, hidden_size)
        self.e2 = nn.Linear(hidden_size, hidden_size)
        self.mean = nn.Linear(hidden_size, latent_dim)
        self.log_std = nn.Linear(hidden_size, latent_dim)
        self.d1 = nn.Linear(obs_dim + latent_dim, hidden_size)
        self.d2 = nn.Linear(hidden_size, hidden_size)
        self.d3 = nn.Linear(hidden_size, act_dim)
        self.act_lim = act_lim
        self.latent_dim = latent_dim
        self.device = device

    def decode(self, obs, z=None):
        if z is None:
            z = torch.randn((obs.shape[0], self.latent_dim)).clamp(-0.5, 0.5).to(obs.device)
        a = F.relu(self.d1(torch.cat([obs, z], 1)))
        a = F.relu(self.d2(a))
        return self.act_lim * torch.tanh(self.d3(a))


class LagrangianPIDController:
    """Host-side view of the PID state (net.py:356-387); the state itself lives on the device
    inside the engine (error_old / error_integral) and is updated by the actor-loss kernel."""

    def __init__(self, KP, KI, KD, thres) -> None:
        self.KP, self.KI, self.KD, self.thres = KP, KI, KD, thres
        self.error_old = 0
        self.error_integral = 0
