"""Shared host-side glue of the model/trainer mirrors: engine binding, stepping, logging."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from ..engine import Engine


class DummyLogger:
    """Stand-in for fsrl.utils.DummyLogger when fsrl is not installed (bcql.py:8)."""

    def store(self, tab=None, **kw):
        pass

    def write(self, *a, **k):
        pass

    write_without_reset = save_config = setup_checkpoint_fn = save_checkpoint = write


def device_index(device) -> int:
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(
            f"osrl_b200 models train on a CUDA device only (got device={device!r}); there is no CPU fallback. "
            "Use the reference implementation for CPU runs.")
    return 0 if d.index is None else d.index


class EngineModel(nn.Module):
    """nn.Module whose parameters become views of an engine arena once bound."""

    algo: str = ""

    def _hyper(self) -> dict:  # constructor hyper-parameters forwarded to the engine
        raise NotImplementedError

    def _bind(self, batch_size: int, lrs: dict, seed: int = 0, world_size: int = 1, rank: int = 0) -> Engine:
        eng = Engine(self.algo, batch_size=batch_size, device=device_index(self.device), seed=seed,
                     world_size=world_size, rank=rank, **self._hyper(), **lrs)
        eng.load_params(self.state_dict())          # reference-order torch init -> arena
        views = eng.param_views()
        for name, p in self.named_parameters():
            p.data = views[name]                    # zero-copy: state_dict()/load_state_dict() hit the arena
            p.requires_grad_(False)
        self._engine = eng
        return eng

    @property
    def engine(self) -> Optional[Engine]:
        return getattr(self, "_engine", None)


class EngineTrainer:
    """Common train_one_step / train_steps plumbing.

    noise="device": Philox noise generated on the GPU (fast path).
    noise="torch":  the trainer draws the raw normals with torch's global generator in the order and
                    shapes the reference consumes them -- same seed, same stream.
    """

    lr_names = ()

    def __init__(self, model: EngineModel, env=None, logger=None, reward_scale: float = 1.0, cost_scale: float = 1.0,
                 device="cuda:0", noise: str = "device", seed: int = 0, log_every: int = 1, lag_stats: bool = False):
        self.model = model
        self.logger = logger if logger is not None else DummyLogger()
        self.env = env
        self.reward_scale = reward_scale
        self.cost_scale = cost_scale
        self.device = device
        self.noise_mode = noise
        self.seed = seed
        self.log_every = log_every
        # lag_stats: logger.store() receives the stats of the PREVIOUS step, read without a stream synchronisation
        # (osrl_stats_lagged): the host queues step s+1 while the GPU still runs step s
        self.lag_stats = lag_stats
        self._n = 0
        self._dataset = None

    # -- engine
    def _engine(self, batch_size: int) -> Engine:
        eng = self.model.engine
        if eng is None:
            eng = self.model._bind(batch_size, self._lrs, seed=self.seed)
            if self._dataset is not None:
                eng.upload_dataset(*self._dataset)
        elif eng.batch_size != batch_size:
            raise RuntimeError(f"engine was built for batch_size={eng.batch_size}, got a batch of {batch_size}")
        return eng

    def _torch_noise(self, eng: Engine) -> Optional[Dict[str, torch.Tensor]]:
        return None

    def _store(self, stats: Dict[str, float]) -> None:
        self.logger.store(**stats)

    def _step(self, batch: dict):
        B = batch["observations"].shape[0]
        eng = self._engine(B)
        noise = self._torch_noise(eng) if self.noise_mode == "torch" else None
        eng.step(batch, noise)
        self._n += 1
        if self.log_every and self._n % self.log_every == 0:
            if self.lag_stats:
                st = eng.stats_lagged()
                if st is not None:
                    self._store(st)
            else:
                self._store(eng.stats())

    # -- resumable checkpoint: {"model_state": ...} stays what the reference writes (train_bcql.py:108-109); the extra
    # key carries what it loses (optimiser moments, targets, PID / dual variables, step counters)
    def checkpoint(self) -> dict:
        out = {"model_state": self.model.state_dict()}
        if self.model.engine is not None:
            out["engine_state"] = self.model.engine.state_blob()
            out["n_steps"] = self._n
        return out

    def load_checkpoint(self, ckpt: dict, batch_size: Optional[int] = None) -> None:
        self.model.load_state_dict(ckpt["model_state"])
        if "engine_state" in ckpt:
            eng = self._engine(batch_size if batch_size is not None else (self.model.engine.batch_size
                                                                          if self.model.engine else 256))
            eng.load_state_blob(ckpt["engine_state"])
            self._n = int(ckpt.get("n_steps", 0))

    # -- fast path: dataset resident in HBM, sampling on the device
    def set_dataset(self, dataset: dict, reward_scale: Optional[float] = None, cost_scale: Optional[float] = None):
        rs = self.reward_scale if reward_scale is None else reward_scale
        cs = self.cost_scale if cost_scale is None else cost_scale
        self._dataset = (dataset, rs, cs)
        if self.model.engine is not None:
            self.model.engine.upload_dataset(dataset, rs, cs)

    # -- evaluation (bcql.py:308-340, cpq.py:315-347, bearl.py:414-446): batch-1 policy calls against a CPU environment,
    # outside the hot path -- plain torch on the arena views, so it always sees the weights the engine just wrote
    def evaluate(self, eval_episodes):
        self.model.eval()
        rets, costs, lens = [], [], []
        for _ in range(eval_episodes):
            r, n, c = self.rollout()
            rets.append(r); lens.append(n); costs.append(c)
        self.model.train()
        return np.mean(rets) / self.reward_scale, np.mean(costs) / self.cost_scale, np.mean(lens)

    # how rollout() calls model.act: BCQ-Lag `act(obs)` (bcql.py:331); CPQ / BEAR-Lag / COptiDICE roll out the
    # DETERMINISTIC policy, `act(obs, True, True)` (cpq.py:338, bearl.py:437, coptidice.py:312)
    act_args = ()

    @torch.no_grad()
    def rollout(self):
        """One episode with the current policy: (return, length, cost) in the trainer's scaled units."""
        obs, info = self.env.reset()
        ret, cost, n = 0.0, 0.0, 0
        for _ in range(self.model.episode_len):
            act, _ = self.model.act(obs, *self.act_args)
            obs, reward, terminated, truncated, info = self.env.step(act)
            ret += reward
            cost += info["cost"] * self.cost_scale
            n += 1
            if terminated or truncated:
                break
        return ret, n, cost

    # -- batched host path: k iterations of the loop body train_bcql.py:142-148 per call, the caller keeps its DataLoader
    batch_keys = ("observations", "next_observations", "actions", "rewards", "costs", "done")

    def train_batches(self, batches, store: bool = True):
        """`train_one_step` on each of k host minibatches, queued in one call (osrl_steps_host): the transfer of batch
        j+1 overlaps the compute of batch j and the host synchronises once.  `batches`: a list of what the reference's
        loader yields (tuples in train_one_step's argument order, or dicts), or one dict of stacked [k, B, ...] tensors.
        Noise is Philox on the device.  Returns the per-step stats; every step's stats go to the logger (store=True)."""
        if isinstance(batches, (list, tuple)) and len(batches) and not isinstance(batches[0], dict):
            batches = [dict(zip(self.batch_keys, b)) for b in batches]
        first = batches[0] if isinstance(batches, (list, tuple)) else batches
        B = first["observations"].shape[-2]
        eng = self._engine(B)
        if self.noise_mode == "torch":
            raise RuntimeError('train_batches draws its noise on the device; noise="torch" needs train_one_step')
        out = eng.steps_host(batches)
        for st in out:
            self._n += 1
            if store and self.log_every and self._n % self.log_every == 0:
                self._store(st)
        return out

    def train_steps(self, n: int, batch_size: Optional[int] = None) -> Dict[str, float]:
        """n gradient steps without touching the host (replaces n iterations of train_bcql.py:142-148)."""
        if self._dataset is None:
            raise RuntimeError("call set_dataset(...) first")
        eng = self._engine(batch_size if batch_size is not None else (self.model.engine.batch_size
                                                                      if self.model.engine else 256))
        eng.steps(n)
        self._n += n
        stats = eng.stats()
        self._store(stats)
        return stats
