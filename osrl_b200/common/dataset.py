"""Minibatch sources and one-time dataset preprocessing with the reference's names (osrl/common/dataset.py).

`TransitionDataset` / `SequenceDataset` keep the reference's constructor and iterator contract (so they can be handed
to ``torch.utils.data.DataLoader`` by the unchanged example scripts) and additionally know how to make themselves
resident in HBM (``to_engine`` / ``trainer.set_dataset``): the data is packed once and all later draws happen on the
device (``osrl_steps``).  The preprocessing helpers (`process_bc_dataset`, Pareto-frontier augmentation, sampling
probabilities) run once before training -- off the hot path -- and are restated here on flat arrays; the 2-D Pareto
front is a sort + scan, so the reference's `oapackage` dependency (dataset.py:9-12, 82-87, 358-365) is not needed.
"""
from __future__ import annotations

import heapq
import random
from collections import Counter

import numpy as np
from torch.utils.data import IterableDataset


def episode_returns(x: np.ndarray, offsets: np.ndarray, gamma: float) -> np.ndarray:
    """First element of discounted_cumsum (dataset.py:19-27) of every episode [offsets[i], offsets[i+1]), evaluated for
    all episodes in lock-step from their last transition backwards: per episode the same float operations in the same
    order as the reference's loop (acc = x[t] + gamma * acc, in x's dtype), without the per-transition Python loop."""
    lens = np.diff(offsets)
    ends = offsets[1:] - 1
    acc = x[ends].copy()
    for back in range(1, int(lens.max()) if lens.size else 0):
        live = lens > back
        idx = ends[live] - back
        acc[live] = x[idx] + gamma * acc[live]
    return acc


def pareto_front_2d(cost: np.ndarray, rew: np.ndarray) -> list:
    """Indices of the points not dominated under (minimise cost, maximise reward) -- what the reference obtains from
    oapackage.ParetoDoubleLong on (-cost, reward) (dataset.py:82-87): sort by cost, scan with the best reward so far.
    Points with identical (cost, reward) on the front are all kept."""
    order = np.lexsort((-rew, cost))          # cost ascending, reward descending inside equal costs
    keep, best = [], -np.inf
    i, n = 0, len(order)
    while i < n:
        j = i
        c = cost[order[i]]
        while j < n and cost[order[j]] == c:
            j += 1
        top = rew[order[i]]
        if top > best:                        # strictly better reward than anything cheaper
            keep.extend(int(order[k]) for k in range(i, j) if rew[order[k]] == top)
            best = top
        i = j
    return sorted(keep)


def process_bc_dataset(dataset: dict, cost_limit: float, gamma: float, bc_mode: str):
    """Filter a transition dataset for the BC variants, in place (dataset.py:30-134; train_bc.py:77).

    Adds "cost_returns" / "rew_returns" (the episode's discounted return, repeated on each of its transitions; a
    trailing unfinished episode keeps 0), then keeps the transitions of: every episode ("all", "multi-task" -- which
    also appends the cost return as an observation feature), episodes within the limit ("safe"), beyond twice the limit
    ("risky"), within (0.5, 1.5] x limit ("boundary"), or within a fifth of the reward range of the fitted Pareto
    frontier ("frontier")."""
    done = np.logical_or(dataset["terminals"] == 1, dataset["timeouts"] == 1)
    ends = np.flatnonzero(done)
    n = dataset["observations"].shape[0]
    offsets = np.concatenate([[0], ends + 1]).astype(np.int64)
    dataset["cost_returns"] = np.zeros_like(dataset["costs"])
    dataset["rew_returns"] = np.zeros_like(dataset["rewards"])
    cost_ret = episode_returns(dataset["costs"], offsets, gamma)
    rew_ret = episode_returns(dataset["rewards"], offsets, gamma)
    lens = np.diff(offsets)
    covered = int(offsets[-1])
    dataset["cost_returns"][:covered] = np.repeat(cost_ret, lens)
    dataset["rew_returns"][:covered] = np.repeat(rew_ret, lens)

    cr = dataset["cost_returns"]
    if bc_mode in ("all", "multi-task"):
        keep = np.ones(n, dtype=bool)
    elif bc_mode == "safe":
        keep = cr <= cost_limit
    elif bc_mode == "risky":
        keep = cr >= 2 * cost_limit
    elif bc_mode == "boundary":
        keep = np.logical_and(0.5 * cost_limit < cr, cr <= 1.5 * cost_limit)
    elif bc_mode == "frontier":
        c64, r64 = cost_ret.astype(np.float64), rew_ret.astype(np.float64)
        span = (np.max(r64) - np.min(r64)) / 5
        pf = pareto_front_2d(c64, r64)
        frontier = None
        for deg in (0, 1, 2):                 # lowest degree that explains 90 % of the front's variance
            frontier = np.poly1d(np.polyfit(c64[pf], r64[pf], deg=deg))
            resid = np.sum((r64[pf] - frontier(c64[pf])) ** 2)
            if 1 - resid / np.sum((r64[pf] - np.mean(r64[pf])) ** 2) >= 0.9:
                break
        curve = frontier(cr)
        keep = np.logical_and(curve - span <= dataset["rew_returns"], dataset["rew_returns"] <= curve + span)
    else:
        raise NotImplementedError
    for k, v in dataset.items():
        dataset[k] = v[keep]
    if bc_mode == "multi-task":
        dataset["observations"] = np.hstack((dataset["observations"], dataset["cost_returns"].reshape(-1, 1)))
    print(f"original size = {n}, cost limit = {cost_limit}, filtered size = {int(np.sum(keep))}")


class TransitionDataset(IterableDataset):
    def __init__(self, dataset: dict, reward_scale: float = 1.0, cost_scale: float = 1.0, state_init: bool = False):
        self.dataset = dataset
        self.reward_scale = reward_scale
        self.cost_scale = cost_scale
        self.sample_prob = None
        self.state_init = state_init
        self.dataset_size = self.dataset["observations"].shape[0]
        self.dataset["done"] = np.logical_or(self.dataset["terminals"], self.dataset["timeouts"]).astype(np.float32)
        if self.state_init:
            init = self.dataset["done"].copy()
            init[1:] = init[:-1]
            init[0] = 1.0
            self.dataset["is_init"] = init

    def get_dataset_states(self):
        return (self.dataset["is_init"].mean(), self.dataset["observations"].std(0, keepdims=True),
                self.dataset["actions"].std(0, keepdims=True))

    def to_engine(self, engine) -> None:
        """Pack into HBM once (osrl_buffer_upload); sampling then happens on the device."""
        engine.upload_dataset(self.dataset, self.reward_scale, self.cost_scale)

    def sample(self, idx):
        d = self.dataset
        out = (d["observations"][idx, :], d["next_observations"][idx, :], d["actions"][idx, :],
               d["rewards"][idx] * self.reward_scale, d["costs"][idx] * self.cost_scale, d["done"][idx])
        return out + ((d["is_init"][idx],) if self.state_init else ())

    def __iter__(self):
        while True:
            yield self.sample(np.random.choice(self.dataset_size, p=self.sample_prob))


def _suffix_sums(x: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """Per-trajectory reward-to-go (discounted_cumsum with gamma=1, dataset.py:19-27), vectorised: a reversed
    float32 running sum restarted at every trajectory end (same left-to-right order of additions per step)."""
    out = np.empty_like(x)
    for s, e in zip(offsets[:-1], offsets[1:]):
        out[s:e] = np.cumsum(x[s:e][::-1], dtype=np.float32)[::-1]
    return out


class SequenceDataset(IterableDataset):
    """Trajectory windows for CDT.  Covers the sampling modes that need no Pareto-frontier augmentation
    (augment_percent=0, random_aug=0, pf_only/pf_sample off -- those require the un-vendored `oapackage`):
    uniform or cost-based trajectory sampling (cost_sample, dataset.py:439-459) and uniform start index."""

    def __init__(self, dataset: dict, seq_len: int = 10, reward_scale: float = 1.0, cost_scale: float = 1.0,
                 deg: int = 3, pf_sample: bool = False, max_rew_decrease: float = 1.0, beta: float = 1.0,
                 augment_percent: float = 0, max_reward: float = 1000.0, min_reward: float = 5,
                 cost_reverse: bool = False, pf_only: bool = False, rmin: float = 0, cost_bins: int = 60, npb: int = 5,
                 cost_sample: bool = False, cost_transform=lambda x: 50 - x, prob: float = 0.4,
                 start_sampling: bool = False, random_aug: float = 0, **_unused):
        if pf_only or pf_sample or augment_percent > 0 or random_aug > 0 or start_sampling:
            raise NotImplementedError("Pareto-frontier / random augmentation and start-index sampling are one-time "
                                      "CPU preprocessing outside the hot path and are not rebuilt here")
        self.seq_len, self.reward_scale, self.cost_scale = seq_len, reward_scale, cost_scale
        ends = np.flatnonzero(np.logical_or(dataset["terminals"], dataset["timeouts"]))
        n = int(ends[-1]) + 1 if ends.size else 0          # a trailing unfinished episode is dropped (:160-168)
        self.offsets = np.concatenate([[0], ends + 1]).astype(np.int64)
        costs = np.asarray(dataset["costs"][:n], dtype=np.float32)
        if cost_reverse:
            costs = 1.0 - costs
        self.flat = {
            "observations": np.asarray(dataset["observations"][:n], dtype=np.float32),
            "actions": np.asarray(dataset["actions"][:n], dtype=np.float32),
            "costs": costs,
        }
        self.flat["returns"] = _suffix_sums(np.asarray(dataset["rewards"][:n], dtype=np.float32), self.offsets)
        self.flat["cost_returns"] = _suffix_sums(costs, self.offsets)
        self.sample_prob = None
        if cost_sample:
            p = np.array([cost_transform(self.flat["cost_returns"][s]) for s in self.offsets[:-1]])  # float32 (:452-458)
            p[p < 0] = 0
            p /= np.sum(p)
            self.sample_prob = p
        print(f"original data: {len(self.offsets) - 1}, augment data: 0, total: {len(self.offsets) - 1}")

    def __len__(self):
        return len(self.offsets) - 1

    def to_engine(self, engine) -> None:
        """Pack into HBM once (osrl_seq_buffer_upload); windows are then drawn on the device."""
        engine.upload_seq_dataset(dict(self.flat, traj_offsets=self.offsets, sample_prob=self.sample_prob),
                                  self.reward_scale, self.cost_scale)

    def sample(self, traj_idx: int, start_idx: int):
        s, e = self.offsets[traj_idx], self.offsets[traj_idx + 1]
        lo, hi = s + start_idx, min(s + start_idx + self.seq_len, e)
        n, T = hi - lo, self.seq_len
        pad = lambda x: np.concatenate([x, np.zeros((T - n,) + x.shape[1:], dtype=x.dtype)], 0) if n < T else x
        f = self.flat
        mask = np.hstack([np.ones(n), np.zeros(T - n)])
        return (pad(f["observations"][lo:hi]), pad(f["actions"][lo:hi]), pad(f["returns"][lo:hi] * self.reward_scale),
                pad(f["cost_returns"][lo:hi] * self.cost_scale), np.arange(start_idx, start_idx + T), mask,
                f["cost_returns"][s] * self.cost_scale, pad(f["costs"][lo:hi]))

    def __iter__(self):
        while True:
            t = np.random.choice(len(self), p=self.sample_prob)
            yield self.sample(t, random.randint(0, int(self.offsets[t + 1] - self.offsets[t]) - 1))
