"""Shared pieces of the CPU oracle (test infrastructure only -- see oracle/__init__.py).

Parameters are kept in an ordered ``dict[str, Tensor]`` keyed by the reference's
``state_dict`` names, so a checkpoint of the reference loads into the oracle (and
into the CUDA engine) without renaming.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

Params = "OrderedDict[str, torch.Tensor]"


# --------------------------------------------------------------------------- init
def _new_linear(params, prefix: str, n_in: int, n_out: int) -> None:
    """One ``nn.Linear`` worth of parameters, drawn with the same torch RNG calls
    (kaiming_uniform weight, then uniform bias) the reference's constructor makes
    (osrl/common/net.py:28), so ``torch.manual_seed(s)`` gives the same init."""
    lin = nn.Linear(n_in, n_out)
    params[prefix + ".weight"] = lin.weight.detach().clone()
    params[prefix + ".bias"] = lin.bias.detach().clone()


def init_mlp(params, prefix: str, sizes: Sequence[int]) -> None:
    """``mlp(sizes, act, out_act)`` -> ``Sequential(Linear, act, Linear, act, ...)``:
    Linear modules sit at even indices 0, 2, 4 (osrl/common/net.py:24-30)."""
    for j in range(len(sizes) - 1):
        _new_linear(params, f"{prefix}.{2 * j}", sizes[j], sizes[j + 1])


def init_double_q(params, prefix: str, in_dim: int, hidden: Sequence[int], num_q: int) -> None:
    """EnsembleDoubleQCritic: q1_nets.* then q2_nets.* (net.py:258-267)."""
    for lst in ("q1_nets", "q2_nets"):
        for i in range(num_q):
            init_mlp(params, f"{prefix}.{lst}.{i}", [in_dim, *hidden, 1])


def init_single_q(params, prefix: str, in_dim: int, hidden: Sequence[int], num_q: int) -> None:
    """EnsembleQCritic: q_nets.* (net.py:223-226)."""
    for i in range(num_q):
        init_mlp(params, f"{prefix}.q_nets.{i}", [in_dim, *hidden, 1])


def init_vae(params, prefix: str, obs_dim: int, act_dim: int, hidden: int, latent: int) -> None:
    """VAE layer creation order e1,e2,mean,log_std,d1,d2,d3 (net.py:305-313)."""
    _new_linear(params, prefix + ".e1", obs_dim + act_dim, hidden)
    _new_linear(params, prefix + ".e2", hidden, hidden)
    _new_linear(params, prefix + ".mean", hidden, latent)
    _new_linear(params, prefix + ".log_std", hidden, latent)
    _new_linear(params, prefix + ".d1", obs_dim + latent, hidden)
    _new_linear(params, prefix + ".d2", hidden, hidden)
    _new_linear(params, prefix + ".d3", hidden, act_dim)


def init_squashed_actor(params, prefix: str, obs_dim: int, act_dim: int, hidden: Sequence[int]) -> None:
    """SquashedGaussianMLPActor: net (trunk), mu_layer, log_std_layer (net.py:165-167)."""
    init_mlp(params, prefix + ".net", [obs_dim, *hidden])
    _new_linear(params, prefix + ".mu_layer", hidden[-1], act_dim)
    _new_linear(params, prefix + ".log_std_layer", hidden[-1], act_dim)


def clone_as(params, src_prefix: str, dst_prefix: str) -> None:
    """``deepcopy(module)`` of the reference (bcql.py:100-105): same values, new names."""
    for k in [k for k in params if k.startswith(src_prefix + ".")]:
        params[dst_prefix + k[len(src_prefix):]] = params[k].clone()


def group(params, prefix: str) -> List[str]:
    return [k for k in params if k.startswith(prefix + ".")]


# --------------------------------------------------------------------------- forward pieces
def mlp_forward(params, prefix: str, x: torch.Tensor, n_layers: int, hidden_act, out_act=None):
    for j in range(n_layers):
        x = F.linear(x, params[f"{prefix}.{2 * j}.weight"], params[f"{prefix}.{2 * j}.bias"])
        if j < n_layers - 1:
            x = hidden_act(x)
        elif out_act is not None:
            x = out_act(x)
    return x


def q_list(params, prefix: str, num: int, n_layers: int, data: torch.Tensor):
    """One ensemble list: each net is a ReLU MLP ending in 1 unit, squeezed (net.py:228-233,274-276)."""
    return [mlp_forward(params, f"{prefix}.{i}", data, n_layers, F.relu).squeeze(-1) for i in range(num)]


def double_q_predict(params, prefix, num, n_layers, obs, act):
    """EnsembleDoubleQCritic.predict (net.py:278-283): per-list min over the ensemble."""
    data = torch.cat([obs, act], dim=-1)
    l1 = q_list(params, prefix + ".q1_nets", num, n_layers, data)
    l2 = q_list(params, prefix + ".q2_nets", num, n_layers, data)
    m1 = torch.min(torch.vstack(l1), dim=0).values
    m2 = torch.min(torch.vstack(l2), dim=0).values
    return m1, m2, l1, l2


def single_q_predict(params, prefix, num, n_layers, obs, act):
    """EnsembleQCritic.predict (net.py:235-238)."""
    data = torch.cat([obs, act], dim=-1)
    lst = q_list(params, prefix + ".q_nets", num, n_layers, data)
    return torch.min(torch.vstack(lst), dim=0).values, lst


def ensemble_mse(target, lst):
    """EnsembleQCritic.loss / EnsembleDoubleQCritic.loss (net.py:240-242, 285-287)."""
    return sum(((q - target) ** 2).mean() for q in lst)


def vae_encode(params, prefix, obs, act):
    h = F.relu(F.linear(torch.cat([obs, act], 1), params[prefix + ".e1.weight"], params[prefix + ".e1.bias"]))
    h = F.relu(F.linear(h, params[prefix + ".e2.weight"], params[prefix + ".e2.bias"]))
    mean = F.linear(h, params[prefix + ".mean.weight"], params[prefix + ".mean.bias"])
    log_std = F.linear(h, params[prefix + ".log_std.weight"], params[prefix + ".log_std.bias"]).clamp(-4, 15)
    return mean, torch.exp(log_std)  # net.py:320-326


def vae_decode_raw(params, prefix, obs, z):
    """Decoder up to the pre-tanh output of d3 (net.py:337-339, 347-353)."""
    a = F.relu(F.linear(torch.cat([obs, z], -1), params[prefix + ".d1.weight"], params[prefix + ".d1.bias"]))
    a = F.relu(F.linear(a, params[prefix + ".d2.weight"], params[prefix + ".d2.bias"]))
    return F.linear(a, params[prefix + ".d3.weight"], params[prefix + ".d3.bias"])


def vae_decode(params, prefix, obs, z, act_lim):
    return act_lim * torch.tanh(vae_decode_raw(params, prefix, obs, z))


def vae_kl(mean, std):
    """-0.5 * mean over ALL B*L elements (bcql.py:125)."""
    return -0.5 * (1 + torch.log(std.pow(2)) - mean.pow(2) - std.pow(2)).mean()


# --------------------------------------------------------------------------- optimiser / targets
class AdamState:
    """torch.optim.Adam, single-tensor CPU path, restated (bias-corrected, eps added
    after sqrt(v_hat); amsgrad/weight_decay off).  Used at bcql.py:222-226 etc."""

    def __init__(self, names: Iterable[str], params, lr: float, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay: float = 0.0, decoupled: bool = False):
        self.names = list(names)
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.wd, self.decoupled = weight_decay, decoupled
        self.t = 0
        self.m = {k: torch.zeros_like(params[k]) for k in self.names}
        self.v = {k: torch.zeros_like(params[k]) for k in self.names}

    def step(self, params, grads: Dict[str, torch.Tensor], lr: float | None = None) -> None:
        lr = self.lr if lr is None else lr
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        with torch.no_grad():
            for k in self.names:
                g = grads[k]
                p = params[k]
                if self.decoupled and self.wd != 0.0:  # AdamW: p *= 1 - lr*wd  (cdt.py:321-326)
                    p.mul_(1 - lr * self.wd)
                self.m[k].lerp_(g, 1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                denom = (self.v[k].sqrt() / bc2_sqrt).add_(self.eps)
                p.addcdiv_(self.m[k], denom, value=-step_size)


def polyak(params, tgt_prefix: str, src_prefix: str, tau: float) -> None:
    """_soft_update (bcql.py:114-120): tgt <- tau*src + (1-tau)*tgt."""
    with torch.no_grad():
        for k in group(params, src_prefix):
            t = params[tgt_prefix + k[len(src_prefix):]]
            t.copy_(tau * params[k] + (1 - tau) * t)


class PID:
    """LagrangianPIDController.control (net.py:376-387)."""

    def __init__(self, kp, ki, kd, thres):
        self.kp, self.ki, self.kd, self.thres = kp, ki, kd, thres
        self.e_old = torch.zeros(())
        self.e_int = torch.zeros(())

    def control(self, qc: torch.Tensor, reduce=None) -> torch.Tensor:
        e_new = torch.mean(qc - self.thres)
        if reduce is not None:  # data-parallel restatement: the mean runs over the global batch
            e_new = reduce(e_new)
        e_diff = F.relu(e_new - self.e_old)
        self.e_int = F.relu(self.e_int + e_new)
        self.e_old = e_new
        return F.relu(self.kp * F.relu(e_new) + self.ki * self.e_int + self.kd * e_diff)


def qc_threshold(cost_limit: float, gamma: float, episode_len: int) -> float:
    """bcql.py:109-110 (python double arithmetic)."""
    return cost_limit * (1 - gamma ** episode_len) / (1 - gamma) / episode_len


def grads_of(loss: torch.Tensor, params, names: List[str]) -> Dict[str, torch.Tensor]:
    gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in zip(names, gs)}


def require_grad(params, names: List[str], flag: bool = True) -> None:
    for k in names:
        params[k].requires_grad_(flag)
