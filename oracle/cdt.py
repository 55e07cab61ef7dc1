"""CPU oracle: CDT ``train_one_step`` and the SequenceDataset sampler, restated in plain PyTorch /
numpy (test infrastructure only -- see oracle/__init__.py).

Covers the configuration every reference task uses (examples/configs/cdt_configs.py:22-90):
time_emb, use_rew, use_cost, cost_transform (50 - ctg), stochastic DiagGaussian head,
action_head_layers=1, no cost prefix / cost features.  Attention is written out explicitly
(scaled QK^T, causal + key-padding mask, softmax, PV) instead of calling nn.MultiheadAttention.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import core as C


@dataclass
class CDTConfig:
    state_dim: int
    action_dim: int
    max_action: float = 1.0
    seq_len: int = 10
    episode_len: int = 1000
    embedding_dim: int = 128
    num_layers: int = 3
    num_heads: int = 8
    attention_dropout: float = 0.0
    residual_dropout: float = 0.0
    embedding_dropout: float = 0.0
    init_temperature: float = 0.1
    target_entropy: Optional[float] = None     # train_cdt.py:110: -action_dim
    learning_rate: float = 1e-4
    weight_decay: float = 1e-4
    betas: tuple = (0.9, 0.999)
    clip_grad: float = 0.25
    lr_warmup_steps: int = 500
    loss_cost_weight: float = 0.02
    loss_state_weight: float = 0.0


def _normal_init(t):  # CDT._init_weights (cdt.py:156-164)
    torch.nn.init.normal_(t, mean=0.0, std=0.02)


class CDTOracle:
    """osrl/algorithms/cdt.py: CDT.__init__ 45-148, forward 166-265, CDTTrainer 291-341, train_one_step
    343-418; TransformerBlock net.py:391-441; DiagGaussianActor net.py:509-533."""

    def __init__(self, cfg: CDTConfig):
        # dropout: the multipliers (0 or 1/(1-p)) are explicit inputs ("noise"), drawn here when not provided, so a
        # step can be replayed elsewhere on the same draws
        self.cfg = cfg
        E, o, a, T = cfg.embedding_dim, cfg.state_dim, cfg.action_dim, cfg.seq_len
        self.H, self.d, self.Lq = cfg.num_heads, E // cfg.num_heads, 4 * T
        # build with the same torch modules in the same order so torch.manual_seed gives the reference's init
        mods = OrderedDict()
        mods["emb_norm"] = nn.LayerNorm(E)
        mods["out_norm"] = nn.LayerNorm(E)
        mods["timestep_emb"] = nn.Embedding(cfg.episode_len + T, E)
        mods["state_emb"] = nn.Linear(o, E)
        mods["action_emb"] = nn.Linear(a, E)
        mods["cost_emb"] = nn.Linear(1, E)
        mods["return_emb"] = nn.Linear(1, E)
        blocks = []
        for _ in range(cfg.num_layers):
            b = OrderedDict()
            b["norm1"] = nn.LayerNorm(E)
            b["norm2"] = nn.LayerNorm(E)
            b["attention"] = nn.MultiheadAttention(E, cfg.num_heads, 0.0, batch_first=True)
            b["mlp.0"] = nn.Linear(E, 4 * E)
            b["mlp.2"] = nn.Linear(4 * E, E)
            blocks.append(b)
        head_mu, head_ls = nn.Linear(E, a), nn.Linear(E, a)
        for m in (head_mu, head_ls):           # DiagGaussianActor.weight_init (net.py:521-528): consumes RNG
            nn.init.orthogonal_(m.weight.data)
            m.bias.data.fill_(0.0)
        state_pred, cost_pred = nn.Linear(E, o), nn.Linear(E, 2)

        def reinit(m):  # self.apply(_init_weights): children first, registration order
            if isinstance(m, (nn.Linear, nn.Embedding)):
                _normal_init(m.weight)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.zeros_(m.bias)
                nn.init.ones_(m.weight)

        order = [mods["emb_norm"], mods["out_norm"], mods["timestep_emb"], mods["state_emb"], mods["action_emb"],
                 mods["cost_emb"], mods["return_emb"]]
        for b in blocks:  # nn.Module.apply visits children depth-first: norm1, norm2, drop, attention(out_proj), mlp
            order += [b["norm1"], b["norm2"], b["attention"].out_proj, b["mlp.0"], b["mlp.2"]]
        order += [head_mu, head_ls, state_pred, cost_pred]
        for m in order:
            reinit(m)

        p = OrderedDict()
        for name in ("emb_norm", "out_norm"):
            p[name + ".weight"], p[name + ".bias"] = mods[name].weight, mods[name].bias
        p["timestep_emb.weight"] = mods["timestep_emb"].weight
        for name in ("state_emb", "action_emb", "cost_emb", "return_emb"):
            p[name + ".weight"], p[name + ".bias"] = mods[name].weight, mods[name].bias
        for i, b in enumerate(blocks):
            pre = f"blocks.{i}."
            p[pre + "norm1.weight"], p[pre + "norm1.bias"] = b["norm1"].weight, b["norm1"].bias
            p[pre + "norm2.weight"], p[pre + "norm2.bias"] = b["norm2"].weight, b["norm2"].bias
            p[pre + "attention.in_proj_weight"] = b["attention"].in_proj_weight
            p[pre + "attention.in_proj_bias"] = b["attention"].in_proj_bias
            p[pre + "attention.out_proj.weight"] = b["attention"].out_proj.weight
            p[pre + "attention.out_proj.bias"] = b["attention"].out_proj.bias
            p[pre + "mlp.0.weight"], p[pre + "mlp.0.bias"] = b["mlp.0"].weight, b["mlp.0"].bias
            p[pre + "mlp.2.weight"], p[pre + "mlp.2.bias"] = b["mlp.2"].weight, b["mlp.2"].bias
        p["action_head.mu.weight"], p["action_head.mu.bias"] = head_mu.weight, head_mu.bias
        p["action_head.log_std.weight"], p["action_head.log_std.bias"] = head_ls.weight, head_ls.bias
        p["state_pred_head.weight"], p["state_pred_head.bias"] = state_pred.weight, state_pred.bias
        p["cost_pred_head.weight"], p["cost_pred_head.bias"] = cost_pred.weight, cost_pred.bias
        self.params = OrderedDict((k, v.detach().clone()) for k, v in p.items())
        self.names = list(self.params)
        self.opt = C.AdamState(self.names, self.params, cfg.learning_rate, cfg.betas, 1e-8, cfg.weight_decay, True)
        # temperature: CPU float64 leaf outside state_dict (cdt.py:144), Adam lr 1e-4 (cdt.py:332-337)
        self.log_temperature = torch.tensor(np.log(cfg.init_temperature))
        self.temp_m = torch.zeros((), dtype=torch.float64)
        self.temp_v = torch.zeros((), dtype=torch.float64)
        self.temp_t = 0
        self.target_entropy = -float(a) if cfg.target_entropy is None else cfg.target_entropy
        self.steps = 0
        self.last_noise: Dict[str, torch.Tensor] = {}
        self.dp = None      # tests: an object with all_reduce(tensor) / world (torch.distributed) -- restates the
                            # engine's data-parallel decomposition (global masked-mean denominators, summed gradients)

    # ---- dropout multipliers, one tensor per site, in the reference's call order (cdt.py:222; net.py:428-440:
    # attention weights inside nn.MultiheadAttention, self.drop(attention_out), the mlp's trailing nn.Dropout)
    def mask_shapes(self, B: int):
        cfg = self.cfg
        L, E, H = 4 * cfg.seq_len, cfg.embedding_dim, cfg.num_heads
        out = OrderedDict()
        if cfg.embedding_dropout > 0:
            out["drop_emb"] = ((B, L, E), cfg.embedding_dropout)
        for i in range(cfg.num_layers):
            if cfg.attention_dropout > 0:
                out[f"drop_attn{i}"] = ((B, H, L, L), cfg.attention_dropout)
            if cfg.residual_dropout > 0:
                out[f"drop_res{i}a"] = ((B, L, E), cfg.residual_dropout)
                out[f"drop_res{i}b"] = ((B, L, E), cfg.residual_dropout)
        return out

    def draw_masks(self, B: int, generator=None):
        return OrderedDict((k, (torch.rand(shape, generator=generator) >= pr).float() / (1.0 - pr))
                           for k, (shape, pr) in self.mask_shapes(B).items())

    # ---- model
    def forward(self, states, actions, returns, costs_return, time_steps, mask, drop=None):
        p, cfg = self.params, self.cfg
        drop = drop or {}
        B, T = states.shape[:2]
        E, H, d = cfg.embedding_dim, self.H, self.d
        te = p["timestep_emb.weight"][time_steps]                                   # cdt.py:180
        s = F.linear(states, p["state_emb.weight"], p["state_emb.bias"]) + te
        a_ = F.linear(actions, p["action_emb.weight"], p["action_emb.bias"]) + te
        c = F.linear((50 - costs_return).unsqueeze(-1), p["cost_emb.weight"], p["cost_emb.bias"]) + te   # :187-193
        r = F.linear(returns.unsqueeze(-1), p["return_emb.weight"], p["return_emb.bias"]) + te
        x = torch.stack([r, c, s, a_], dim=1).permute(0, 2, 1, 3).reshape(B, 4 * T, E)                  # :198-200
        pad = torch.stack([~mask.bool()] * 4, dim=1).permute(0, 2, 1).reshape(B, -1)                     # :203-205
        x = F.layer_norm(x, (E,), p["emb_norm.weight"], p["emb_norm.bias"])
        if "drop_emb" in drop:
            x = x * drop["drop_emb"]                                                                       # cdt.py:222
        L = 4 * T
        causal = ~torch.tril(torch.ones(L, L)).bool()                                                    # net.py:417-418
        for i in range(cfg.num_layers):
            pre = f"blocks.{i}."
            h = F.layer_norm(x, (E,), p[pre + "norm1.weight"], p[pre + "norm1.bias"])
            qkv = F.linear(h, p[pre + "attention.in_proj_weight"], p[pre + "attention.in_proj_bias"])
            q, k, v = (t.reshape(B, L, H, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
            sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
            sc = sc.masked_fill(causal[None, None] | pad[:, None, None, :], float("-inf"))
            pw = torch.softmax(sc, dim=-1)
            if f"drop_attn{i}" in drop:
                pw = pw * drop[f"drop_attn{i}"]
            o_ = (pw @ v).transpose(1, 2).reshape(B, L, E)
            ao = F.linear(o_, p[pre + "attention.out_proj.weight"], p[pre + "attention.out_proj.bias"])
            if f"drop_res{i}a" in drop:
                ao = ao * drop[f"drop_res{i}a"]
            x = x + ao
            h2 = F.layer_norm(x, (E,), p[pre + "norm2.weight"], p[pre + "norm2.bias"])
            m = F.gelu(F.linear(h2, p[pre + "mlp.0.weight"], p[pre + "mlp.0.bias"]))
            mo = F.linear(m, p[pre + "mlp.2.weight"], p[pre + "mlp.2.bias"])
            if f"drop_res{i}b" in drop:
                mo = mo * drop[f"drop_res{i}b"]
            x = x + mo
        out = F.layer_norm(x, (E,), p["out_norm.weight"], p["out_norm.bias"]).reshape(B, T, 4, E)
        state_feat, action_feat = out[:, :, 2], out[:, :, 3]                                             # :239-240
        mu = F.linear(state_feat, p["action_head.mu.weight"], p["action_head.mu.bias"])
        log_std = F.linear(state_feat, p["action_head.log_std.weight"], p["action_head.log_std.bias"])
        cost_preds = F.log_softmax(F.linear(action_feat, p["cost_pred_head.weight"], p["cost_pred_head.bias"]), -1)
        state_preds = F.linear(action_feat, p["state_pred_head.weight"], p["state_pred_head.bias"])
        return mu, log_std, cost_preds, state_preds

    def step(self, states, actions, returns, costs_return, time_steps, mask, episode_cost, costs, noise=None):
        p, cfg = self.params, self.cfg
        C.require_grad(p, self.names)
        drop = OrderedDict()
        shapes = self.mask_shapes(states.shape[0])
        if shapes:
            given = noise or {}
            fresh = self.draw_masks(states.shape[0])
            for k, (shape, _) in shapes.items():
                drop[k] = torch.as_tensor(given[k]).reshape(shape).float() if k in given else fresh[k]
        self.last_noise = dict(drop)
        mu, log_std, cost_preds, state_preds = self.forward(states, actions, returns, costs_return, time_steps, mask,
                                                            drop)
        std = log_std.exp()
        valid = mask > 0
        logp = -((actions - mu) ** 2) / (2 * std ** 2) - log_std - math.log(math.sqrt(2 * math.pi))
        ent = 0.5 + 0.5 * math.log(2 * math.pi) + log_std
        temp = self.log_temperature.exp().detach()
        cl = F.nll_loss(cost_preds.reshape(-1, 2), costs.flatten().long(), reduction="none")
        pred = cost_preds.reshape(-1, 2).max(dim=1)[1]
        sl = F.mse_loss(state_preds[:, :-1], states[:, 1:], reduction="none")
        if self.dp is None:
            ll = logp[valid].mean()                                                      # cdt.py:358
            entropy = ent[valid].mean()                                                  # :359
            cost_loss = (cl * mask.flatten()).mean()                                     # :378-380
            acc = (pred.eq(costs.flatten().long()) * mask.flatten()).sum() / mask.sum()
            state_loss = (sl * mask[:, :-1].unsqueeze(-1)).mean()                        # :388-392
        else:
            # data parallel (what k_cdt_loss does in two phases): the means are over the GLOBAL batch, so every rank
            # divides its partial sums by the all-reduced counts; the summed gradients are then the global gradient
            def gsum(x):
                t = x.detach().double().reshape(1).clone()
                self.dp.all_reduce(t)
                return float(t)
            n_valid = gsum(valid.sum()) * logp.shape[-1]
            n_bt = gsum(torch.tensor(float(mask.numel())))
            n_sl = gsum(torch.tensor(float(sl.numel())))
            ll_l, ent_l = logp[valid].sum() / n_valid, ent[valid].sum() / n_valid
            cost_l = (cl * mask.flatten()).sum() / n_bt
            state_l = (sl * mask[:, :-1].unsqueeze(-1)).sum() / n_sl
            # local contributions carry the gradient, the reported / temperature values are the global sums
            ll = ll_l + (gsum(ll_l) - ll_l.detach())
            entropy = ent_l + (gsum(ent_l) - ent_l.detach())
            cost_loss = cost_l + (gsum(cost_l) - cost_l.detach())
            state_loss = state_l + (gsum(state_l) - state_l.detach())
            acc = torch.tensor(gsum((pred.eq(costs.flatten().long()) * mask.flatten()).sum()) / gsum(mask.sum()))
        act_loss = -(ll + temp * entropy)                                            # :365
        loss = act_loss + cfg.loss_cost_weight * cost_loss + cfg.loss_state_weight * state_loss
        g = C.grads_of(loss, p, self.names)
        C.require_grad(p, self.names, False)
        g = {k: v.to(p[k].dtype) for k, v in g.items()}
        if self.dp is not None:                 # gradient all-reduce (sum), before the global-norm clip
            for v in g.values():
                self.dp.all_reduce(v)
        # clip_grad_norm_ (cdt.py:399)
        total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(v) for v in g.values()]))
        coef = torch.clamp(cfg.clip_grad / (total + 1e-6), max=1.0)
        self.last_grads = g                       # raw gradients (what the engine keeps in its G section)
        self.last_clip_coef = float(coef)
        g = {k: v * coef for k, v in g.items()}
        lr = cfg.learning_rate * min((self.steps + 1) / cfg.lr_warmup_steps, 1)       # LambdaLR, cdt.py:327-330
        self.opt.step(p, g, lr=lr)
        # temperature Adam (cdt.py:402-407), float64 scalar
        gT = (self.log_temperature.exp() * (entropy.detach().double() - self.target_entropy))
        self.temp_t += 1
        self.temp_m = self.temp_m + (1 - 0.9) * (gT - self.temp_m)
        self.temp_v = 0.999 * self.temp_v + (1 - 0.999) * gT * gT
        denom = self.temp_v.sqrt() / math.sqrt(1 - 0.999 ** self.temp_t) + 1e-8
        self.log_temperature = self.log_temperature - (1e-4 / (1 - 0.9 ** self.temp_t)) * self.temp_m / denom
        self.steps += 1
        next_lr = cfg.learning_rate * min((self.steps + 1) / cfg.lr_warmup_steps, 1)
        return {"nll": -ll.item(), "ent": entropy.item(), "ent_reg": temp.item(), "all_loss": loss.item(),
                "act_loss": act_loss.item(), "cost_loss": cost_loss.item(), "cost_acc": acc.item(),
                "state_loss": state_loss.item(), "train_lr": next_lr}


# =========================================================================== SequenceDataset sampler
def discounted_cumsum(x: np.ndarray, gamma: float) -> np.ndarray:
    """dataset.py:19-27"""
    out = np.zeros_like(x)
    out[-1] = x[-1]
    for t in reversed(range(x.shape[0] - 1)):
        out[t] = x[t] + gamma * out[t + 1]
    return out


def split_trajectories(dataset: dict, cost_reverse: bool = False):
    """process_sequence_dataset (dataset.py:137-183): list of per-episode dicts with returns / cost_returns
    (undiscounted suffix sums); cost_reverse: costs become 1.0 - cost (:164-165)."""
    trajs, start = [], 0
    n = dataset["rewards"].shape[0]
    for i in range(n):
        if dataset["terminals"][i] or dataset["timeouts"][i]:   # a trailing unfinished episode is dropped (:160)
            sl = slice(start, i + 1)
            ep = {k: np.asarray(dataset[k][sl], dtype=np.float32) for k in ("observations", "actions", "rewards", "costs")}
            if cost_reverse:
                ep["costs"] = np.array([1.0 - c for c in dataset["costs"][sl]], dtype=np.float32)
            ep["returns"] = discounted_cumsum(ep["rewards"], 1.0)
            ep["cost_returns"] = discounted_cumsum(ep["costs"], 1.0)
            trajs.append(ep)
            start = i + 1
    return trajs


def sequence_sample(trajs, traj_idx: int, start_idx: int, seq_len: int, reward_scale: float, cost_scale: float):
    """SequenceDataset.__prepare_sample (dataset.py:749-775): slice, scale, end-zero-pad, mask."""
    tr = trajs[traj_idx]
    sl = slice(start_idx, start_idx + seq_len)
    states, actions = tr["observations"][sl], tr["actions"][sl]
    returns, cost_returns = tr["returns"][sl] * reward_scale, tr["cost_returns"][sl] * cost_scale
    costs = tr["costs"][sl]
    time_steps = np.arange(start_idx, start_idx + seq_len)
    n = states.shape[0]
    mask = np.hstack([np.ones(n), np.zeros(seq_len - n)])
    pad = lambda x: np.concatenate([x, np.zeros((seq_len - n,) + x.shape[1:], dtype=x.dtype)], 0) if n < seq_len else x
    episode_cost = tr["cost_returns"][0] * cost_scale
    return pad(states), pad(actions), pad(returns), pad(cost_returns), time_steps, mask, episode_cost, pad(costs)
