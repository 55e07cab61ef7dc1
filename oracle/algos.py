"""CPU oracle: one ``train_one_step`` of BC / BCQ-Lag / CPQ / BEAR-Lag, restated in
plain PyTorch fp32 (test infrastructure only -- see oracle/__init__.py).

Every ``step`` takes the minibatch and, optionally, the *raw standard-normal* noise
tensors it should consume (noise replay).  When ``noise`` is
``None`` the oracle draws them from torch's global CPU generator in exactly the
order and shapes the reference does, so seeding both identically keeps them in
lock-step; whatever was consumed is left in ``self.last_noise``.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import core as C


_DTYPE = [torch.float32]


class precision:
    """Context manager: run the oracle arithmetic in another dtype (float64 conditioning probe in tests)."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        _DTYPE.append(self.dtype)

    def __exit__(self, *a):
        _DTYPE.pop()


def _draw(noise: Optional[Dict[str, torch.Tensor]], used: Dict[str, torch.Tensor], name: str, shape):
    x = torch.randn(shape) if noise is None else noise[name].reshape(shape)
    x = x.to(_DTYPE[-1])
    used[name] = x
    return x


# =========================================================================== BC
@dataclass
class BCConfig:
    state_dim: int
    action_dim: int
    max_action: float = 1.0
    a_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    actor_lr: float = 1e-4


class BCOracle:
    """osrl/algorithms/bc.py: BC.__init__ 26-43, actor_loss 45-52, BCTrainer.train_one_step 103-109."""

    def __init__(self, cfg: BCConfig):
        self.cfg = cfg
        p = OrderedDict()
        C.init_mlp(p, "actor.pi", [cfg.state_dim, *cfg.a_hidden_sizes, cfg.action_dim])  # net.py:78-81
        self.params = p
        self.nl = len(cfg.a_hidden_sizes) + 1
        self.g_actor = C.group(p, "actor")
        self.opt = C.AdamState(self.g_actor, p, cfg.actor_lr)
        self.last_noise: Dict[str, torch.Tensor] = {}

    def step(self, observations, actions, noise=None):
        p, cfg = self.params, self.cfg
        C.require_grad(p, self.g_actor)
        pred = cfg.max_action * C.mlp_forward(p, "actor.pi", observations, self.nl, F.relu, torch.tanh)  # net.py:83-85
        loss = F.mse_loss(pred, actions)  # bc.py:47
        g = C.grads_of(loss, p, self.g_actor)
        C.require_grad(p, self.g_actor, False)
        if getattr(self, "dp", None) is not None:
            g = self.dp.reduce_grads(g)
        self.last_grads = dict(g)
        self.opt.step(p, g)
        return {"loss/actor_loss": loss.item()}


# =========================================================================== BCQ-Lag
@dataclass
class BCQLConfig:
    state_dim: int
    action_dim: int
    max_action: float = 1.0
    a_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    c_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    vae_hidden_sizes: int = 64
    sample_action_num: int = 10
    gamma: float = 0.99
    tau: float = 0.005
    phi: float = 0.05
    lmbda: float = 0.75
    beta: float = 0.5
    PID: List[float] = field(default_factory=lambda: [0.1, 0.003, 0.001])
    num_q: int = 1
    num_qc: int = 1
    cost_limit: float = 10
    episode_len: int = 300
    actor_lr: float = 1e-4
    critic_lr: float = 1e-4
    vae_lr: float = 1e-4


class BCQLOracle:
    """osrl/algorithms/bcql.py: BCQL.__init__ 44-112, losses 122-216, sync_weight 228-234,
    BCQLTrainer.train_one_step 283-306."""

    def __init__(self, cfg: BCQLConfig):
        self.cfg = cfg
        o, a = cfg.state_dim, cfg.action_dim
        self.L = 2 * a
        p = OrderedDict()
        C.init_mlp(p, "actor.pi", [o + a, *cfg.a_hidden_sizes, a])          # net.py:54-55
        C.init_double_q(p, "critic", o + a, cfg.c_hidden_sizes, cfg.num_q)
        C.init_double_q(p, "cost_critic", o + a, cfg.c_hidden_sizes, cfg.num_qc)
        C.init_vae(p, "vae", o, a, cfg.vae_hidden_sizes, self.L)
        C.clone_as(p, "actor", "actor_old")
        C.clone_as(p, "critic", "critic_old")
        C.clone_as(p, "cost_critic", "cost_critic_old")
        self.params = p
        self.na = len(cfg.a_hidden_sizes) + 1
        self.nc = len(cfg.c_hidden_sizes) + 1
        self.qc_thres = C.qc_threshold(cfg.cost_limit, cfg.gamma, cfg.episode_len)
        self.pid = C.PID(*cfg.PID, self.qc_thres)
        self.g = {k: C.group(p, k) for k in ("actor", "critic", "cost_critic", "vae")}
        self.opt = {
            "actor": C.AdamState(self.g["actor"], p, cfg.actor_lr),
            "critic": C.AdamState(self.g["critic"], p, cfg.critic_lr),
            "cost_critic": C.AdamState(self.g["cost_critic"], p, cfg.critic_lr),
            "vae": C.AdamState(self.g["vae"], p, cfg.vae_lr),
        }
        self.last_noise: Dict[str, torch.Tensor] = {}

    # -- pieces
    def _perturb(self, prefix, obs, act):
        """MLPGaussianPerturbationActor.forward (net.py:59-62); hidden act is Tanh (bcql.py:86-88)."""
        cfg = self.cfg
        out = C.mlp_forward(self.params, prefix + ".pi", torch.cat([obs, act], 1), self.na, torch.tanh, torch.tanh)
        return (cfg.phi * cfg.max_action * out + act).clamp(-cfg.max_action, cfg.max_action)

    def _decode(self, obs, raw_z):
        """VAE.decode with z ~ N(0,1).clamp(-0.5, 0.5) (net.py:332-339)."""
        return C.vae_decode(self.params, "vae", obs, raw_z.clamp(-0.5, 0.5), self.cfg.max_action)

    def _q_target(self, critic_old: str, num: int, next_obs, raw_z):
        cfg = self.cfg
        B = next_obs.shape[0]
        obs_next = torch.repeat_interleave(next_obs, cfg.sample_action_num, 0)      # bcql.py:138
        act_next = self._perturb("actor_old", obs_next, self._decode(obs_next, raw_z))  # :141
        q1, q2, _, _ = C.double_q_predict(self.params, critic_old, num, self.nc, obs_next, act_next)
        q = cfg.lmbda * torch.min(q1, q2) + (1.0 - cfg.lmbda) * torch.max(q1, q2)     # :144-145
        return q.reshape(B, -1).max(1)[0]                                            # :146

    def _update(self, name: str, loss: torch.Tensor) -> None:
        g = C.grads_of(loss, self.params, self.g[name])
        C.require_grad(self.params, self.g[name], False)
        if getattr(self, "dp", None) is not None:
            g = self.dp.reduce_grads(g)
        self.last_grads = getattr(self, "last_grads", {})
        self.last_grads.update(g)
        self.opt[name].step(self.params, g)

    def step(self, observations, next_observations, actions, rewards, costs, done, noise=None):
        p, cfg = self.params, self.cfg
        B, S, L = observations.shape[0], cfg.sample_action_num, self.L
        used: Dict[str, torch.Tensor] = {}
        stats = {}

        # ---- VAE (bcql.py:122-132)
        C.require_grad(p, self.g["vae"])
        mean, std = C.vae_encode(p, "vae", observations, actions)
        z = mean + std * _draw(noise, used, "vae_eps", (B, L))                       # net.py:327
        recon = C.vae_decode(p, "vae", observations, z, cfg.max_action)
        loss_vae = F.mse_loss(recon, actions) + cfg.beta * C.vae_kl(mean, std)
        stats["loss/loss_vae"] = loss_vae.item()
        self._update("vae", loss_vae)

        # ---- reward critic (bcql.py:134-155)
        C.require_grad(p, self.g["critic"])
        _, _, q1l, q2l = C.double_q_predict(p, "critic", cfg.num_q, self.nc, observations, actions)
        with torch.no_grad():
            q_t = self._q_target("critic_old", cfg.num_q, next_observations, _draw(noise, used, "z_critic", (B * S, L)))
            backup = rewards + cfg.gamma * (1 - done) * q_t
        loss_c = C.ensemble_mse(backup, q1l) + C.ensemble_mse(backup, q2l)
        stats["loss/critic_loss"] = loss_c.item()
        self._update("critic", loss_c)

        # ---- cost critic (bcql.py:157-179): no (1-done) factor
        C.require_grad(p, self.g["cost_critic"])
        _, _, q1l, q2l = C.double_q_predict(p, "cost_critic", cfg.num_qc, self.nc, observations, actions)
        with torch.no_grad():
            q_t = self._q_target("cost_critic_old", cfg.num_qc, next_observations,
                                 _draw(noise, used, "z_cost", (B * S, L)))
            backup = costs + cfg.gamma * q_t
        loss_cc = C.ensemble_mse(backup, q1l) + C.ensemble_mse(backup, q2l)
        stats["loss/cost_critic_loss"] = loss_cc.item()
        self._update("cost_critic", loss_cc)

        # ---- actor (bcql.py:181-216)
        C.require_grad(p, self.g["actor"])
        act_pi = self._perturb("actor", observations, self._decode(observations, _draw(noise, used, "z_actor", (B, L))))
        q1, q2, _, _ = C.double_q_predict(p, "critic", cfg.num_q, self.nc, observations, act_pi)
        qc1, qc2, _, _ = C.double_q_predict(p, "cost_critic", cfg.num_qc, self.nc, observations, act_pi)
        qc_pi, q_pi = torch.min(qc1, qc2), torch.min(q1, q2)
        with torch.no_grad():
            mult = self.pid.control(qc_pi, getattr(self, "dp", None) and self.dp.mean)
        qc_pen = ((qc_pi - self.qc_thres) * mult).mean()
        loss_a = -q_pi.mean() + qc_pen
        stats["loss/actor_loss"] = loss_a.item()
        stats["loss/qc_penalty"] = qc_pen.item()
        stats["loss/lagrangian"] = mult.item()
        self._update("actor", loss_a)

        # ---- targets (bcql.py:228-234)
        C.polyak(p, "critic_old", "critic", cfg.tau)
        C.polyak(p, "cost_critic_old", "cost_critic", cfg.tau)
        C.polyak(p, "actor_old", "actor", cfg.tau)
        self.last_noise = used
        return stats


# =========================================================================== shared squashed-Gaussian actor
def squashed_actor(params, prefix, obs, eps, n_trunk):
    """SquashedGaussianMLPActor.forward (net.py:169-205) with rsample = mu + std*eps.
    Returns (tanh(u), u, mu, std)."""
    h = C.mlp_forward(params, prefix + ".net", obs, n_trunk, F.relu, F.relu)
    mu = F.linear(h, params[prefix + ".mu_layer.weight"], params[prefix + ".mu_layer.bias"])
    log_std = F.linear(h, params[prefix + ".log_std_layer.weight"], params[prefix + ".log_std_layer.bias"])
    std = torch.exp(torch.clamp(log_std, -20, 2))
    u = mu + std * eps
    return torch.tanh(u), u, mu, std


# =========================================================================== CPQ
@dataclass
class CPQConfig:
    state_dim: int
    action_dim: int
    max_action: float = 1.0
    a_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    c_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    vae_hidden_sizes: int = 64
    sample_action_num: int = 10
    gamma: float = 0.99
    tau: float = 0.005
    beta: float = 1.5
    num_q: int = 1
    num_qc: int = 1
    qc_scalar: float = 1.5
    cost_limit: float = 10
    episode_len: int = 300
    actor_lr: float = 1e-4
    critic_lr: float = 1e-4
    alpha_lr: float = 1e-4
    vae_lr: float = 1e-4


class CPQOracle:
    """osrl/algorithms/cpq.py: CPQ.__init__ 38-105, losses 125-222, sync_weight 224-230,
    CPQTrainer.train_one_step 294-313."""

    def __init__(self, cfg: CPQConfig):
        self.cfg = cfg
        o, a = cfg.state_dim, cfg.action_dim
        self.L = 2 * a
        p = OrderedDict()
        C.init_squashed_actor(p, "actor", o, a, cfg.a_hidden_sizes)
        C.init_single_q(p, "critic", o + a, cfg.c_hidden_sizes, cfg.num_q)
        C.init_vae(p, "vae", o, a, cfg.vae_hidden_sizes, self.L)
        C.init_single_q(p, "cost_critic", o + a, cfg.c_hidden_sizes, cfg.num_qc)
        C.clone_as(p, "actor", "actor_old")
        C.clone_as(p, "critic", "critic_old")
        C.clone_as(p, "cost_critic", "cost_critic_old")
        self.params = p
        self.nt = len(cfg.a_hidden_sizes)
        self.nc = len(cfg.c_hidden_sizes) + 1
        self.q_thres = C.qc_threshold(cfg.cost_limit, cfg.gamma, cfg.episode_len)   # cpq.py:103-105
        self.qc_thres = cfg.qc_scalar * self.q_thres
        self.log_alpha = torch.tensor(0.0)
        self.g = {k: C.group(p, k) for k in ("actor", "critic", "cost_critic", "vae")}
        self.opt = {
            "actor": C.AdamState(self.g["actor"], p, cfg.actor_lr),
            "critic": C.AdamState(self.g["critic"], p, cfg.critic_lr),
            "cost_critic": C.AdamState(self.g["cost_critic"], p, cfg.critic_lr),
            "vae": C.AdamState(self.g["vae"], p, cfg.vae_lr),
        }
        self.last_noise: Dict[str, torch.Tensor] = {}

    def _update(self, name, loss):
        g = C.grads_of(loss, self.params, self.g[name])
        C.require_grad(self.params, self.g[name], False)
        if getattr(self, "dp", None) is not None:
            g = self.dp.reduce_grads(g)
        self.last_grads = getattr(self, "last_grads", {})
        self.last_grads.update(g)
        self.opt[name].step(self.params, g)

    def _pi(self, obs, eps):
        a, _, mu, std = squashed_actor(self.params, "actor", obs, eps, self.nt)
        return a * self.cfg.max_action, mu, std                                      # cpq.py:115-123

    def step(self, observations, next_observations, actions, rewards, costs, done, noise=None):
        p, cfg = self.params, self.cfg
        B, S, L, a = observations.shape[0], cfg.sample_action_num, self.L, cfg.action_dim
        used: Dict[str, torch.Tensor] = {}
        stats = {}

        # ---- VAE (cpq.py:125-135)
        C.require_grad(p, self.g["vae"])
        mean, std = C.vae_encode(p, "vae", observations, actions)
        z = mean + std * _draw(noise, used, "vae_eps", (B, L))
        recon = C.vae_decode(p, "vae", observations, z, cfg.max_action)
        loss_vae = F.mse_loss(recon, actions) + cfg.beta * C.vae_kl(mean, std)
        stats["loss/loss_vae"] = loss_vae.item()
        self._update("vae", loss_vae)

        # ---- reward critic (cpq.py:137-153): next action from the CURRENT actor
        C.require_grad(p, self.g["critic"])
        _, ql = C.single_q_predict(p, "critic", cfg.num_q, self.nc, observations, actions)
        with torch.no_grad():
            na, _, _ = self._pi(next_observations, _draw(noise, used, "pi_critic", (B, a)))
            q_t, _ = C.single_q_predict(p, "critic_old", cfg.num_q, self.nc, next_observations, na)
            qc_t, _ = C.single_q_predict(p, "cost_critic_old", cfg.num_qc, self.nc, next_observations, na)
            backup = rewards + cfg.gamma * (1 - done) * (qc_t <= self.q_thres) * q_t
        loss_c = C.ensemble_mse(backup, ql)
        stats["loss/critic_loss"] = loss_c.item()
        self._update("critic", loss_c)

        # ---- cost critic (cpq.py:155-201)
        C.require_grad(p, self.g["cost_critic"])
        _, qcl = C.single_q_predict(p, "cost_critic", cfg.num_qc, self.nc, observations, actions)
        with torch.no_grad():
            na, _, _ = self._pi(next_observations, _draw(noise, used, "pi_cost", (B, a)))
            qc_t, _ = C.single_q_predict(p, "cost_critic_old", cfg.num_qc, self.nc, next_observations, na)
            backup = costs + cfg.gamma * qc_t
            # :164 -- the rsample inside this call consumes RNG; only the distribution is kept
            _, mu, sd = self._pi(observations, _draw(noise, used, "pi_dist", (B, a)))
            # :166 Normal(mu, std).sample([S]) -- pre-tanh, unscaled, S-major
            sampled = (mu.unsqueeze(0) + sd.unsqueeze(0) * _draw(noise, used, "ood_sample", (S, B, a))).reshape(S * B, a)
            stacked = torch.tile(observations[None], (S, 1, 1)).reshape(S * B, cfg.state_dim)
            qc_s, _ = C.single_q_predict(p, "cost_critic_old", cfg.num_qc, self.nc, stacked, sampled)
            qc_s = qc_s.reshape(S, B)
            m2, s2 = C.vae_encode(p, "vae", stacked, sampled)                        # :178 (decoder output unused)
            _draw(noise, used, "ood_vae_eps", (S * B, L))                            # RNG consumed by randn_like
            m2, s2 = m2.reshape(S, B, L), s2.reshape(S, B, L)
            kl = -0.5 * (1 + torch.log(s2.pow(2)) - m2.pow(2) - s2.pow(2)).mean(2)
            quant = torch.quantile(kl, 0.75)
            qc_ood = ((kl >= quant) * qc_s).mean(0)
        loss_cc = C.ensemble_mse(backup, qcl) - self.log_alpha.exp() * (qc_ood.mean() - self.qc_thres)
        stats["loss/cost_critic_loss"] = loss_cc.item()
        self._update("cost_critic", loss_cc)
        self.log_alpha = self.log_alpha + cfg.alpha_lr * self.log_alpha.exp() * (self.qc_thres - qc_ood.mean())
        self.log_alpha = self.log_alpha.clamp(-5.0, 5.0)
        stats["loss/alpha_value"] = self.log_alpha.exp().item()

        # ---- actor (cpq.py:203-222)
        C.require_grad(p, self.g["actor"])
        act_pi, _, _ = self._pi(observations, _draw(noise, used, "pi_actor", (B, a)))
        q_pi, _ = C.single_q_predict(p, "critic", cfg.num_q, self.nc, observations, act_pi)
        qc_pi, _ = C.single_q_predict(p, "cost_critic", cfg.num_qc, self.nc, observations, act_pi)
        loss_a = -((qc_pi <= self.q_thres) * q_pi).mean()
        stats["loss/actor_loss"] = loss_a.item()
        self._update("actor", loss_a)

        C.polyak(p, "critic_old", "critic", cfg.tau)
        C.polyak(p, "cost_critic_old", "cost_critic", cfg.tau)
        C.polyak(p, "actor_old", "actor", cfg.tau)
        self.last_noise = used
        return stats


# =========================================================================== BEAR-Lag
@dataclass
class BEARLConfig:
    state_dim: int
    action_dim: int
    max_action: float = 1.0
    a_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    c_hidden_sizes: List[int] = field(default_factory=lambda: [128, 128])
    vae_hidden_sizes: int = 64
    sample_action_num: int = 10
    gamma: float = 0.99
    tau: float = 0.005
    beta: float = 0.5
    lmbda: float = 0.75
    mmd_sigma: float = 50
    target_mmd_thresh: float = 0.05
    num_samples_mmd_match: int = 10
    PID: List[float] = field(default_factory=lambda: [0.1, 0.003, 0.001])
    kernel: str = "gaussian"
    num_q: int = 1
    num_qc: int = 1
    cost_limit: float = 10
    episode_len: int = 300
    start_update_policy_step: int = 20_000
    actor_lr: float = 1e-4
    critic_lr: float = 1e-4
    vae_lr: float = 1e-4
    alpha_lr: float = 1e-3


def mmd(x, y, sigma, kernel):
    """mmd_loss_laplacian / mmd_loss_gaussian (bearl.py:283-318); x, y: [B, N, d] -> [B]."""
    def k(u, v):
        d = u.unsqueeze(2) - v.unsqueeze(1)
        d = d.abs().sum(-1) if kernel == "laplacian" else d.pow(2).sum(-1)
        return torch.mean((-d / (2.0 * sigma)).exp(), dim=(1, 2))
    return (k(x, x) + k(y, y) - 2.0 * k(x, y) + 1e-6).sqrt()


class BEARLOracle:
    """osrl/algorithms/bearl.py: BEARL.__init__ 46-124, losses 134-280, mmd 283-318,
    sync_weight 329-335, BEARLTrainer.train_one_step 389-412."""

    def __init__(self, cfg: BEARLConfig):
        self.cfg = cfg
        o, a = cfg.state_dim, cfg.action_dim
        self.L = 2 * a
        p = OrderedDict()
        C.init_squashed_actor(p, "actor", o, a, cfg.a_hidden_sizes)
        C.init_double_q(p, "critic", o + a, cfg.c_hidden_sizes, cfg.num_q)
        C.init_double_q(p, "cost_critic", o + a, cfg.c_hidden_sizes, cfg.num_qc)
        C.init_vae(p, "vae", o, a, cfg.vae_hidden_sizes, self.L)
        C.clone_as(p, "actor", "actor_old")
        C.clone_as(p, "critic", "critic_old")
        C.clone_as(p, "cost_critic", "cost_critic_old")
        self.params = p
        self.nt = len(cfg.a_hidden_sizes)
        self.nc = len(cfg.c_hidden_sizes) + 1
        self.qc_thres = C.qc_threshold(cfg.cost_limit, cfg.gamma, cfg.episode_len)
        self.pid = C.PID(*cfg.PID, self.qc_thres)
        self.log_alpha = torch.tensor(0.0)
        self.n_train_steps = 0
        self.g = {k: C.group(p, k) for k in ("actor", "critic", "cost_critic", "vae")}
        self.opt = {
            "actor": C.AdamState(self.g["actor"], p, cfg.actor_lr),
            "critic": C.AdamState(self.g["critic"], p, cfg.critic_lr),
            "cost_critic": C.AdamState(self.g["cost_critic"], p, cfg.critic_lr),
            "vae": C.AdamState(self.g["vae"], p, cfg.vae_lr),
        }
        self.last_noise: Dict[str, torch.Tensor] = {}

    def _update(self, name, loss):
        g = C.grads_of(loss, self.params, self.g[name])
        C.require_grad(self.params, self.g[name], False)
        if getattr(self, "dp", None) is not None:
            g = self.dp.reduce_grads(g)
        self.last_grads = getattr(self, "last_grads", {})
        self.last_grads.update(g)
        self.opt[name].step(self.params, g)

    def _q_target(self, critic_old, num, next_obs, eps):
        cfg = self.cfg
        B = next_obs.shape[0]
        obs_next = torch.repeat_interleave(next_obs, cfg.sample_action_num, 0)
        # bearl.py:163 -- actor_old(...) output is tanh(u), NOT scaled by max_action
        act_next, _, _, _ = squashed_actor(self.params, "actor_old", obs_next, eps, self.nt)
        q1, q2, _, _ = C.double_q_predict(self.params, critic_old, num, self.nc, obs_next, act_next)
        q = cfg.lmbda * torch.min(q1, q2) + (1.0 - cfg.lmbda) * torch.max(q1, q2)
        return q.reshape(B, -1).max(1)[0]

    def step(self, observations, next_observations, actions, rewards, costs, done, noise=None):
        p, cfg = self.params, self.cfg
        B, S, L, a, N = observations.shape[0], cfg.sample_action_num, self.L, cfg.action_dim, cfg.num_samples_mmd_match
        used: Dict[str, torch.Tensor] = {}
        stats = {}

        # ---- VAE (bearl.py:144-154)
        C.require_grad(p, self.g["vae"])
        mean, std = C.vae_encode(p, "vae", observations, actions)
        z = mean + std * _draw(noise, used, "vae_eps", (B, L))
        recon = C.vae_decode(p, "vae", observations, z, cfg.max_action)
        loss_vae = F.mse_loss(recon, actions) + cfg.beta * C.vae_kl(mean, std)
        stats["loss/loss_vae"] = loss_vae.item()
        self._update("vae", loss_vae)

        # ---- critic (bearl.py:156-179)
        C.require_grad(p, self.g["critic"])
        _, _, q1l, q2l = C.double_q_predict(p, "critic", cfg.num_q, self.nc, observations, actions)
        with torch.no_grad():
            q_t = self._q_target("critic_old", cfg.num_q, next_observations, _draw(noise, used, "pi_critic", (B * S, a)))
            backup = rewards + cfg.gamma * (1 - done) * q_t
        loss_c = C.ensemble_mse(backup, q1l) + C.ensemble_mse(backup, q2l)
        stats["loss/critic_loss"] = loss_c.item()
        self._update("critic", loss_c)

        # ---- cost critic (bearl.py:181-206)
        C.require_grad(p, self.g["cost_critic"])
        _, _, q1l, q2l = C.double_q_predict(p, "cost_critic", cfg.num_qc, self.nc, observations, actions)
        with torch.no_grad():
            q_t = self._q_target("cost_critic_old", cfg.num_qc, next_observations,
                                 _draw(noise, used, "pi_cost", (B * S, a)))
            backup = costs + cfg.gamma * q_t
        loss_cc = C.ensemble_mse(backup, q1l) + C.ensemble_mse(backup, q2l)
        stats["loss/cost_critic_loss"] = loss_cc.item()
        self._update("cost_critic", loss_cc)

        # ---- actor (bearl.py:208-280)
        C.require_grad(p, self.g["actor"])
        zc = _draw(noise, used, "z_mmd", (B, N, L)).clamp(-0.5, 0.5)                   # net.py:343-345
        obs_rep = observations.unsqueeze(1).expand(B, N, cfg.state_dim)
        raw_vae = C.vae_decode_raw(p, "vae", obs_rep, zc)                              # [B,N,a] pre-tanh (net.py:353)
        stacked = torch.repeat_interleave(observations, N, 0)
        samp, raw, _, _ = squashed_actor(p, "actor", stacked, _draw(noise, used, "pi_actor", (B * N, a)), self.nt)
        samp, raw = samp.reshape(B, N, a), raw.reshape(B, N, a)
        mmd_loss = mmd(raw_vae, raw, cfg.mmd_sigma, cfg.kernel)
        q1, q2, _, _ = C.double_q_predict(p, "critic", cfg.num_q, self.nc, observations, samp[:, 0, :])
        qc1, qc2, _, _ = C.double_q_predict(p, "cost_critic", cfg.num_qc, self.nc, observations, samp[:, 0, :])
        qc_val, q_val = torch.min(qc1, qc2), torch.min(q1, q2)
        with torch.no_grad():
            mult = self.pid.control(qc_val, getattr(self, "dp", None) and self.dp.mean)
        qc_pen = ((qc_val - self.qc_thres) * mult).mean()
        if self.n_train_steps >= cfg.start_update_policy_step:
            loss_a = (-q_val + self.log_alpha.exp() * (mmd_loss - cfg.target_mmd_thresh)).mean()
        else:
            loss_a = (self.log_alpha.exp() * (mmd_loss - cfg.target_mmd_thresh)).mean()
        loss_a = loss_a + qc_pen
        stats["loss/actor_loss"] = loss_a.item()
        stats["loss/mmd_loss"] = mmd_loss.mean().item()
        stats["loss/qc_penalty"] = qc_pen.item()
        stats["loss/lagrangian"] = mult.item()
        self._update("actor", loss_a)
        mmd_mean = (mmd_loss - cfg.target_mmd_thresh).mean().detach()
        if getattr(self, "dp", None) is not None:
            mmd_mean = self.dp.mean(mmd_mean)
        self.log_alpha = self.log_alpha + cfg.alpha_lr * self.log_alpha.exp() * mmd_mean
        self.log_alpha = self.log_alpha.clamp(-5.0, 5.0)
        self.n_train_steps += 1
        stats["loss/alpha_value"] = self.log_alpha.exp().item()

        C.polyak(p, "critic_old", "critic", cfg.tau)
        C.polyak(p, "cost_critic_old", "cost_critic", cfg.tau)
        C.polyak(p, "actor_old", "actor", cfg.tau)
        self.last_noise = used
        return stats


# =========================================================================== data-parallel restatement
class DataParallel:
    """What the N-rank engine does, restated with torch.distributed on any backend: every rank steps on
    its B/N rows, gradients are summed and divided by N before each optimiser update, and the two
    cross-batch scalars (PID error net.py:380, mean MMD bearl.py:261) are averaged over ranks."""

    def __init__(self, dist):
        self.dist = dist
        self.world = dist.get_world_size()

    def reduce_grads(self, g):
        out = {}
        for k, v in g.items():
            t = v.clone()
            self.dist.all_reduce(t)
            out[k] = t / self.world
        return out

    def mean(self, x):
        t = x.detach().clone().reshape(1)
        self.dist.all_reduce(t)
        return (t / self.world).reshape(())
