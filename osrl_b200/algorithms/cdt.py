"""CDT / CDTTrainer with the reference's signatures (osrl/algorithms/cdt.py:45-560), for the mode every
reference task config uses: time_emb, use_rew, use_cost, cost_transform, stochastic head, 1-layer action
head, no cost prefix / cost features.  Training (forward, losses, backward, AdamW, the three dropouts) runs in the
engine; `forward` / `evaluate` / `rollout` below are the evaluation-time path (batch 1 against a CPU environment,
outside the hot path): plain torch on the parameter views, dropout inactive as in `model.eval()`."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from ._base import EngineModel, EngineTrainer


class _Block(nn.Module):
    """Parameter shell of TransformerBlock (net.py:391-441); forward = the evaluation-time block."""

    def __init__(self, seq_len, embedding_dim, num_heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.attention = nn.MultiheadAttention(embedding_dim, num_heads, 0.0, batch_first=True)
        self.mlp = nn.Sequential(nn.Linear(embedding_dim, 4 * embedding_dim), nn.GELU(),
                                 nn.Linear(4 * embedding_dim, embedding_dim), nn.Dropout(0.0))
        self.register_buffer("causal_mask", ~torch.tril(torch.ones(seq_len, seq_len)).to(bool))

    def forward(self, x, padding_mask=None):
        n = x.shape[1]
        h = self.norm1(x)
        x = x + self.attention(h, h, h, attn_mask=self.causal_mask[:n, :n], key_padding_mask=padding_mask,
                               need_weights=False)[0]
        return x + self.mlp(self.norm2(x))


class _DiagGaussianActor(nn.Module):
    def __init__(self, hidden_dim, act_dim):
        super().__init__()
        self.mu = nn.Linear(hidden_dim, act_dim)
        self.log_std = nn.Linear(hidden_dim, act_dim)
        for m in (self.mu, self.log_std):   # net.py:521-528 (consumes RNG like the reference)
            nn.init.orthogonal_(m.weight.data)
            m.bias.data.fill_(0.0)

    def forward(self, feat):
        return torch.distributions.Normal(self.mu(feat), self.log_std(feat).exp())   # net.py:530-533


class CDT(EngineModel):
    algo = "cdt"

    def __init__(self, state_dim: int, action_dim: int, max_action: float, seq_len: int = 10, episode_len: int = 1000,
                 embedding_dim: int = 128, num_layers: int = 4, num_heads: int = 8, attention_dropout: float = 0.0,
                 residual_dropout: float = 0.0, embedding_dropout: float = 0.0, time_emb: bool = True,
                 use_rew: bool = False, use_cost: bool = False, cost_transform: bool = False,
                 add_cost_feat: bool = False, mul_cost_feat: bool = False, cat_cost_feat: bool = False,
                 action_head_layers: int = 1, cost_prefix: bool = False, stochastic: bool = False,
                 init_temperature=0.1, target_entropy=None, device: str = "cuda:0"):
        super().__init__()
        if not (time_emb and use_rew and use_cost and cost_transform and stochastic) or add_cost_feat or \
                mul_cost_feat or cat_cost_feat or cost_prefix or action_head_layers != 1:
            raise NotImplementedError("osrl_b200 CDT covers the configured reference mode (time_emb, use_rew, use_cost, "
                                      "cost_transform, stochastic, action_head_layers=1, no prefix / cost features)")
        self.attention_dropout, self.residual_dropout = float(attention_dropout), float(residual_dropout)
        self.embedding_dropout = float(embedding_dropout)
        self.seq_len, self.embedding_dim, self.state_dim, self.action_dim = seq_len, embedding_dim, state_dim, action_dim
        self.episode_len, self.max_action, self.stochastic, self.device = episode_len, max_action, stochastic, device
        self.num_layers, self.num_heads = num_layers, num_heads
        self.init_temperature = init_temperature
        self.target_entropy = -float(action_dim) if target_entropy is None else float(target_entropy)
        # registration / construction order of the reference (cdt.py:85-141), then _init_weights (:148-164)
        self.emb_norm = nn.LayerNorm(embedding_dim)
        self.out_norm = nn.LayerNorm(embedding_dim)
        self.timestep_emb = nn.Embedding(episode_len + seq_len, embedding_dim)
        self.state_emb = nn.Linear(state_dim, embedding_dim)
        self.action_emb = nn.Linear(action_dim, embedding_dim)
        self.cost_emb = nn.Linear(1, embedding_dim)
        self.return_emb = nn.Linear(1, embedding_dim)
        self.blocks = nn.ModuleList([_Block(4 * seq_len, embedding_dim, num_heads) for _ in range(num_layers)])
        self.action_head = _DiagGaussianActor(embedding_dim, action_dim)
        self.state_pred_head = nn.Linear(embedding_dim, state_dim)
        self.cost_pred_head = nn.Linear(embedding_dim, 2)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(module: nn.Module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            torch.nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                torch.nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            torch.nn.init.zeros_(module.bias)
            torch.nn.init.ones_(module.weight)

    def forward(self, states, actions, returns_to_go, costs_to_go, time_steps, padding_mask=None, episode_cost=None):
        """cdt.py:166-265 for the configured mode: tokens (rtg, ctg, state, action) per step, action distribution from
        the state token, cost class log-probabilities and next-state prediction from the action token."""
        B, T = states.shape[0], states.shape[1]
        te = self.timestep_emb(time_steps)
        tok = [self.return_emb(returns_to_go.unsqueeze(-1)) + te,
               self.cost_emb((50 - costs_to_go.detach()).unsqueeze(-1)) + te,      # cost_transform (cdt.py:79)
               self.state_emb(states) + te, self.action_emb(actions) + te]
        x = torch.stack(tok, dim=1).permute(0, 2, 1, 3).reshape(B, 4 * T, self.embedding_dim)
        if padding_mask is not None:
            padding_mask = torch.stack([padding_mask] * 4, dim=1).permute(0, 2, 1).reshape(B, -1)
        x = self.emb_norm(x)
        for blk in self.blocks:
            x = blk(x, padding_mask=padding_mask)
        x = self.out_norm(x).reshape(B, T, 4, self.embedding_dim).permute(0, 2, 1, 3)
        act_feat, state_feat = x[:, 3], x[:, 2]
        return (self.action_head(state_feat), torch.log_softmax(self.cost_pred_head(act_feat), dim=-1),
                self.state_pred_head(act_feat))

    def temperature(self):
        e = self.engine
        return torch.tensor(np.exp(e.scalars()["log_temperature"]) if e else self.init_temperature)

    def _hyper(self):
        return dict(state_dim=self.state_dim, action_dim=self.action_dim, max_action=self.max_action,
                    seq_len=self.seq_len, episode_len=self.episode_len, embedding_dim=self.embedding_dim,
                    num_layers=self.num_layers, num_heads=self.num_heads, use_rew=1, use_cost=1, cost_transform=1,
                    stochastic=1, init_temperature=self.init_temperature, target_entropy=self.target_entropy,
                    attention_dropout=self.attention_dropout, residual_dropout=self.residual_dropout,
                    embedding_dropout=self.embedding_dropout)


class CDTTrainer(EngineTrainer):
    def __init__(self, model: CDT, env=None, logger=None, learning_rate: float = 1e-4, weight_decay: float = 1e-4,
                 betas: Tuple[float, ...] = (0.9, 0.999), clip_grad: float = 0.25, lr_warmup_steps: int = 10000,
                 reward_scale: float = 1.0, cost_scale: float = 1.0, loss_cost_weight: float = 0.0,
                 loss_state_weight: float = 0.0, cost_reverse: bool = False, no_entropy: bool = False,
                 device="cuda:0", **kw):
        super().__init__(model, env, logger, reward_scale, cost_scale, device, **kw)
        if no_entropy:
            raise NotImplementedError("no_entropy=True is not built")
        self.clip_grad, self.cost_weight, self.state_weight = clip_grad, loss_cost_weight, loss_state_weight
        self._lrs = dict(learning_rate=learning_rate, weight_decay=weight_decay, betas=tuple(betas),
                         clip_grad=0.0 if clip_grad is None else clip_grad, lr_warmup_steps=lr_warmup_steps,
                         loss_cost_weight=loss_cost_weight, loss_state_weight=loss_state_weight)
        self.stochastic, self.max_action, self.cost_reverse = model.stochastic, model.max_action, cost_reverse

    def _store(self, stats):
        self.logger.store(tab="train", **stats)

    def set_dataset(self, seq_dataset):
        """A SequenceDataset -> resident in HBM (fast path)."""
        self._dataset = seq_dataset
        if self.model.engine is not None:
            seq_dataset.to_engine(self.model.engine)

    def _engine(self, batch_size: int):
        eng = self.model.engine
        if eng is None:
            eng = self.model._bind(batch_size, self._lrs, seed=self.seed)
            if self._dataset is not None:
                self._dataset.to_engine(eng)
        elif eng.batch_size != batch_size:
            raise RuntimeError(f"engine was built for batch_size={eng.batch_size}, got a batch of {batch_size}")
        return eng

    def evaluate(self, num_rollouts, target_return, target_cost):
        """cdt.py:420-435: mean return / cost / length over `num_rollouts` episodes conditioned on the targets."""
        self.model.eval()
        out = [self.rollout(self.model, self.env, target_return, target_cost) for _ in range(num_rollouts)]
        self.model.train()
        return (np.mean([o[0] for o in out]) / self.reward_scale, np.mean([o[2] for o in out]) / self.cost_scale,
                np.mean([o[1] for o in out]))

    @torch.no_grad()
    def rollout(self, model, env, target_return: float, target_cost: float):
        """cdt.py:437-518: autoregressive episode -- the last seq_len steps of (state, action, return-to-go,
        cost-to-go) condition the next action; the targets are decremented by what the environment returned."""
        dev = model.state_emb.weight.device
        L = model.episode_len
        states = torch.zeros(1, L + 1, model.state_dim, device=dev)
        actions = torch.zeros(1, L, model.action_dim, device=dev)
        rtg, ctg = torch.zeros(1, L + 1, device=dev), torch.zeros(1, L + 1, device=dev)
        ts = torch.arange(L, dtype=torch.long, device=dev).view(1, -1)
        obs, info = env.reset()
        states[:, 0] = torch.as_tensor(obs, device=dev)
        rtg[:, 0], ctg[:, 0] = float(target_return), float(target_cost)
        epi_cost = torch.tensor([float(target_cost)], device=dev)
        ret, cost, n = 0.0, 0.0, 0
        for t in range(L):
            w = slice(max(0, t + 1 - model.seq_len), t + 1)
            dist, _, _ = model(states[:, w], actions[:, w], rtg[:, w], ctg[:, w], ts[:, w], None, epi_cost)
            a = (dist.mean if self.stochastic else dist).clamp(-self.max_action, self.max_action)[0, -1].cpu().numpy()
            obs, reward, terminated, truncated, info = env.step(a)
            c = ((1.0 - info["cost"]) if self.cost_reverse else info["cost"]) * self.cost_scale
            actions[:, t] = torch.as_tensor(a, device=dev)
            states[:, t + 1] = torch.as_tensor(obs, device=dev)
            rtg[:, t + 1] = rtg[:, t] - reward
            ctg[:, t + 1] = ctg[:, t] - c
            ret += reward
            n += 1
            cost += info["cost"]
            if terminated or truncated:
                break
        return ret, n, cost

    def train_one_step(self, states, actions, returns, costs_return, time_steps, mask, episode_cost, costs):
        eng = self._engine(states.shape[0])
        eng.step_seq({"states": states, "actions": actions, "returns": returns, "costs_return": costs_return,
                      "time_steps": time_steps, "mask": mask, "costs": costs})
        self._n += 1
        if self.log_every and self._n % self.log_every == 0:
            self._store(eng.stats())
