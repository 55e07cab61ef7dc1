"""TEST INFRASTRUCTURE ONLY -- CPU restatement of COptiDICE.update (osrl/algorithms/coptidice.py:125-227) with explicit
Adam, pinned against the unmodified reference by oracle/make_golden.py (`python -m oracle.make_golden coptidice`).

COptiDICE is SURVEY.md section 8(f) rank 1: the next algorithm for the CUDA engine.  It is not built there yet; this
restatement and its fixture (tests/golden/coptidice_small.npz) are the parity checker that work will be held to.
Only tests/, __graft_entry__.smoke() and bench.py's CPU baselines may import oracle/.

Step semantics worth spelling out (all from the source):
  * nu / chi are EnsembleQCritic(state_dim, act_dim=0): MLPs on the observation only; `predict` = min over the
    ensemble (net.py:235-238).
  * lambda' = softplus(lambda) and tau' = softplus(tau) are computed ONCE at the top of the step; the actor phase at
    the end re-evaluates w(s,a) with the UPDATED nu network but the OLD lambda' (coptidice.py:204-206 reuses
    self._lmbda, which predates lmbda_optim.step()).
  * chi, tau, nu, lambda are stepped in that order, every gradient taken on the graph built before any of them moved.
  * softmax / log_softmax of the chi logits run over the BATCH dimension (:167-168).
  * policy extraction adds Gaussian noise scaled by 0.1 x the dataset std to observations and actions, and the actor's
    forward draws an (unused) rsample: three normal draws per step, in the order obs, act, pi (:201-203, net.py:186).
  * the actor loss is the w-weighted log-density of the (noisy) dataset action under the PRE-tanh Normal(mu, std),
    no tanh correction (:208).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import core as C


@dataclass
class COptiDICEConfig:
    state_dim: int
    action_dim: int
    max_action: float = 1.0
    f_type: str = "softchi"
    init_state_propotion: float = 1.0
    a_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    c_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    gamma: float = 0.99
    alpha: float = 0.5
    cost_ub_epsilon: float = 0.01
    num_nu: int = 1
    num_chi: int = 1
    cost_limit: float = 10
    episode_len: int = 300
    actor_lr: float = 1e-3
    critic_lr: float = 1e-3
    scalar_lr: float = 1e-3


def f_div(f_type: str):
    """get_f_div_fn (coptidice.py:15-38)."""
    if f_type == "chi2":
        return (lambda x: 0.5 * (x - 1) ** 2), (lambda x: x + 1)
    if f_type == "softchi":
        return (lambda x: torch.where(x < 1, x * (torch.log(x + 1e-10) - 1) + 1, 0.5 * (x - 1) ** 2),
                lambda x: torch.where(x < 0, torch.exp(x.clamp(max=0.0)), x + 1))
    if f_type == "kl":
        return (lambda x: x * torch.log(x + 1e-10)), (lambda x: torch.exp(x - 1))
    raise NotImplementedError(f_type)


class COptiDICEOracle:
    """COptiDICE.__init__ (coptidice.py:68-123), update (:125-227), setup_optimizers (:229-234)."""

    def __init__(self, cfg: COptiDICEConfig, observations_std, actions_std):
        self.cfg = cfg
        p = OrderedDict()
        # construction order of the reference: tau, lmbda (torch.ones, no RNG), actor, nu_network, chi_network
        C.init_squashed_actor(p, "actor", cfg.state_dim, cfg.action_dim, cfg.a_hidden_sizes)
        C.init_single_q(p, "nu_network", cfg.state_dim, cfg.c_hidden_sizes, cfg.num_nu)
        C.init_single_q(p, "chi_network", cfg.state_dim, cfg.c_hidden_sizes, cfg.num_chi)
        p["tau"] = torch.ones(1)
        p["lmbda"] = torch.ones(1)
        self.params = p
        self.g = {"actor": C.group(p, "actor"), "nu": C.group(p, "nu_network"), "chi": C.group(p, "chi_network"),
                  "tau": ["tau"], "lmbda": ["lmbda"]}
        self.opt = {"actor": C.AdamState(self.g["actor"], p, cfg.actor_lr),
                    "nu": C.AdamState(self.g["nu"], p, cfg.critic_lr),
                    "chi": C.AdamState(self.g["chi"], p, cfg.critic_lr),
                    "lmbda": C.AdamState(self.g["lmbda"], p, cfg.scalar_lr),
                    "tau": C.AdamState(self.g["tau"], p, cfg.scalar_lr)}
        self.qc_thres = C.qc_threshold(cfg.cost_limit, cfg.gamma, cfg.episode_len)
        self.f_fn, self.f_prime_inv = f_div(cfg.f_type)
        self.obs_std = torch.as_tensor(observations_std, dtype=torch.float32)
        self.act_std = torch.as_tensor(actions_std, dtype=torch.float32)
        self.nl_a, self.nl_c = len(cfg.a_hidden_sizes), len(cfg.c_hidden_sizes) + 1
        self.last_noise: Dict[str, torch.Tensor] = {}

    def _v(self, prefix, num, obs):   # EnsembleQCritic.predict(obs, None): min over the ensemble
        lst = C.q_list(self.params, prefix + ".q_nets", num, self.nl_c, obs)
        return torch.min(torch.vstack(lst), dim=0).values

    def _optimal_w(self, obs, nobs, rew, cost, done, lm):
        cfg = self.cfg
        nu_s = self._v("nu_network", cfg.num_nu, obs)
        nu_next = self._v("nu_network", cfg.num_nu, nobs)
        e = rew - lm.detach() * cost
        e = e + cfg.gamma * (1.0 - done) * nu_next - nu_s
        w = F.relu(self.f_prime_inv(e / cfg.alpha))
        return nu_s, nu_next, e, w

    def step(self, observations, next_observations, actions, rewards, costs, done, is_init,
             noise: Optional[Dict[str, torch.Tensor]] = None):
        p, cfg = self.params, self.cfg
        names = [k for g in self.g.values() for k in g]
        C.require_grad(p, names)
        B = observations.shape[0]
        lm = F.softplus(p["lmbda"])                                         # :128
        nu_s, _, e, w = self._optimal_w(observations, next_observations, rewards, costs, done, lm)
        nu_init = nu_s * is_init / cfg.init_state_propotion
        w_ng = w.detach()
        Df = self.f_fn(w_ng).mean()
        tau = F.softplus(p["tau"])
        grads: Dict[str, torch.Tensor] = {}
        if cfg.cost_ub_epsilon == 0:
            weighted_c = (w_ng * costs).mean()
            chi_loss = tau_loss = D_kl = torch.zeros(1)
        else:
            chi_s = self._v("chi_network", cfg.num_chi, observations)
            chi_next = self._v("chi_network", cfg.num_chi, next_observations)
            chi_init = chi_s * is_init / cfg.init_state_propotion
            ell = (1 - cfg.gamma) * chi_init + w_ng * (costs + cfg.gamma * (1 - done) * chi_next - chi_s)
            logits = ell / tau.detach()
            weights = torch.softmax(logits, dim=0) * B                      # over the batch (:167)
            log_weights = torch.log_softmax(logits, dim=0) + np.log(B)
            D_kl = (weights * log_weights - weights + 1).mean()
            weighted_c = (weights * w_ng * costs).mean()
            chi_loss = (weights * ell).mean()
            grads.update(C.grads_of(chi_loss, p, self.g["chi"]))            # backward(retain_graph=True) (:176)
            tau_loss = tau * (cfg.cost_ub_epsilon - D_kl.detach())
            grads.update(C.grads_of(tau_loss.sum(), p, self.g["tau"]))
        nu_loss = (1 - cfg.gamma) * nu_init.mean() + (w * e - cfg.alpha * self.f_fn(w)).mean()
        td_error = e.pow(2).mean()
        grads.update(C.grads_of(nu_loss, p, self.g["nu"]))
        lmbda_loss = lm * (self.qc_thres - weighted_c.detach())
        grads.update(C.grads_of(lmbda_loss.sum(), p, self.g["lmbda"]))
        C.require_grad(p, names, False)
        # optimiser steps in the reference's order; every gradient above was taken before any parameter moved
        if cfg.cost_ub_epsilon != 0:
            self.opt["chi"].step(p, grads)
            self.opt["tau"].step(p, grads)
        self.opt["nu"].step(p, grads)
        self.opt["lmbda"].step(p, grads)

        # ---- 2. policy extraction (:200-212)
        used: Dict[str, torch.Tensor] = {}

        def draw(name, like):
            x = torch.randn(like.shape) if noise is None else torch.as_tensor(noise[name]).reshape(like.shape).float()
            used[name] = x
            return x

        obs_eps = draw("obs_eps", observations) * self.obs_std * 0.1
        act_eps = draw("act_eps", actions) * self.act_std * 0.1
        draw("pi", actions)                                                 # rsample inside actor.forward, unused
        C.require_grad(p, self.g["actor"])
        h = C.mlp_forward(p, "actor.net", observations + obs_eps, self.nl_a, F.relu, F.relu)
        mu = F.linear(h, p["actor.mu_layer.weight"], p["actor.mu_layer.bias"])
        log_std = torch.clamp(F.linear(h, p["actor.log_std_layer.weight"], p["actor.log_std_layer.bias"]), -20, 2)
        std = torch.exp(log_std)
        with torch.no_grad():   # updated nu, the step's original lambda' (:204-206)
            _, _, _, w2 = self._optimal_w(observations, next_observations, rewards, costs, done, lm.detach())
        x = actions + act_eps
        logp = (-((x - mu) ** 2) / (2 * std ** 2) - log_std - math.log(math.sqrt(2 * math.pi))).sum(-1)
        actor_loss = -(w2 * logp).mean()
        ga = C.grads_of(actor_loss, p, self.g["actor"])
        C.require_grad(p, self.g["actor"], False)
        self.opt["actor"].step(p, ga)
        grads.update(ga)
        self.last_grads, self.last_noise = grads, used
        f = lambda t: float(t.detach().reshape(-1)[0])   # noqa: E731
        return {"loss/chi_loss": f(chi_loss), "loss/tau_loss": f(tau_loss), "loss/D_kl": f(D_kl), "loss/Df": f(Df),
                "loss/td_error": f(td_error), "loss/nu_loss": f(nu_loss), "loss/lmbda_loss": f(lmbda_loss),
                "loss/actor_loss": f(actor_loss), "loss/tau": f(tau), "loss/lmbda": f(lm)}
