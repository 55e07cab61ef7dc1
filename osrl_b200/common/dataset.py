"""Minibatch sources and one-time dataset preprocessing with the reference's names (osrl/common/dataset.py).

`TransitionDataset` / `SequenceDataset` keep the reference's constructor and iterator contract (so they can be handed
to ``torch.utils.data.DataLoader`` by the unchanged example scripts) and additionally know how to make themselves
resident in HBM (``to_engine`` / ``trainer.set_dataset``): the data is packed once and all later draws happen on the
device (``osrl_steps``).  The preprocessing helpers (`process_bc_dataset`, Pareto-frontier augmentation, sampling
probabilities) run once before training -- off the hot path -- and are restated here on flat arrays; the 2-D Pareto
front is a sort + scan, so the reference's `oapackage` dependency (dataset.py:9-12, 82-87, 358-365) is not needed.
"""
from __future__ import annotations

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


def _grid_filter(x, y, xbins, ybins, max_per_bin, min_per_bin):
    """dataset.py:228-261: bin the points on an xbins x ybins grid over their bounding box; bins with more than
    max_per_bin points are thinned with random.sample, bins with <= min_per_bin points are dropped (outliers).
    Bins are visited in order of first appearance, like the reference's dict, so `random` is consumed identically."""
    xmin, xmax, ymin, ymax = min(x), max(x), min(y), max(y)
    xs, ys = (xmax - xmin) / xbins, (ymax - ymin) / ybins
    bins = {}
    for i in range(len(x)):
        bins.setdefault(((x[i] - xmin) // xs, (y[i] - ymin) // ys), []).append(i)
    keep = []
    for members in bins.values():
        if len(members) > max_per_bin:
            keep += random.sample(members, max_per_bin)
        elif len(members) > min_per_bin:
            keep += members
    return keep


def _nearest_sources(points: np.ndarray, targets: np.ndarray, max_rew_decrease: float, beta: float) -> list:
    """dataset.py:186-225: for every sampled (cost, reward) target the closest trajectory that is not costlier; when a
    trajectory is chosen k > 1 times the k-1 extra copies are redistributed over its cheaper, not-much-worse
    neighbours with probability ~ 1 / (distance + beta)."""
    ids = np.arange(points.shape[0])
    first = []
    for p in targets:
        ok = points[:, 0] <= p[0]
        d = points[ok] - p
        first.append(ids[ok][np.argmin(np.hypot(d[:, 0], d[:, 1]))])
    out = []
    for idx, k in Counter(first).items():
        out.append(idx)
        if k > 1:
            p = points[idx]
            ok = np.logical_and(points[:, 0] <= p[0], points[:, 1] >= p[1] - max_rew_decrease)
            d = points[ok] - p
            w = 1 / (np.hypot(d[:, 0], d[:, 1]) + beta)
            pick = np.random.choice(w.shape[0], size=k - 1, p=w / np.sum(w))
            out.extend(ids[ok][pick.tolist()])
    return out


def _gauss_kernel(size, std=1.0):
    x = np.linspace(-int(size), int(size), 2 * int(size) + 1)
    return np.exp(-(x ** 2 / std))


class SequenceDataset(IterableDataset):
    """Trajectory windows for CDT (dataset.py:633-787), every sampling / augmentation mode of the reference:
    Pareto-frontier augmentation (`augment_percent`, the default of examples/configs/cdt_configs.py), random
    augmentation (`random_aug`), frontier-only data (`pf_only`), cost-based or frontier-distance trajectory
    sampling (`cost_sample`, `pf_sample`), start-index sampling (`start_sampling`).

    Trajectories are kept as slices of flat arrays (no per-transition Python loop as in process_sequence_dataset,
    dataset.py:137-183); an augmented trajectory shares the observations / actions / costs of its source and owns
    only its relabelled return-to-go and cost-to-go (dataset.py:378-386).  `to_engine` materialises everything into
    one packed buffer in HBM; `start_sampling` is a host-iterator feature only (the device sampler draws uniform
    start indices)."""

    def __init__(self, dataset: dict, seq_len: int = 10, reward_scale: float = 1.0, cost_scale: float = 1.0,
                 deg: int = 3, pf_sample: bool = False, max_rew_decrease: float = 1.0, beta: float = 1.0,
                 augment_percent: float = 0, max_reward: float = 1000.0, min_reward: float = 5,
                 cost_reverse: bool = False, pf_only: bool = False, rmin: float = 0, cost_bins: int = 60, npb: int = 5,
                 cost_sample: bool = False, cost_transform=lambda x: 50 - x, prob: float = 0.4,
                 start_sampling: bool = False, random_aug: float = 0, aug_rmin: float = 0, aug_rmax: float = 600,
                 aug_cmin: float = 5, aug_cmax: float = 50, cgap: float = 5, rstd: float = 1, cstd: float = 0.2):
        self.seq_len, self.reward_scale, self.cost_scale = seq_len, reward_scale, cost_scale
        self.start_sampling = start_sampling
        ends = np.flatnonzero(np.logical_or(dataset["terminals"], dataset["timeouts"]))
        n = int(ends[-1]) + 1 if ends.size else 0          # a trailing unfinished episode is dropped (:160-168)
        offsets = np.concatenate([[0], ends + 1]).astype(np.int64)
        costs = np.asarray(dataset["costs"][:n], dtype=np.float32)
        if cost_reverse:
            costs = 1.0 - costs
        obs = np.asarray(dataset["observations"][:n], dtype=np.float32)
        act = np.asarray(dataset["actions"][:n], dtype=np.float32)
        rew = np.asarray(dataset["rewards"][:n], dtype=np.float32)
        rtg, ctg = _suffix_sums(rew, offsets), _suffix_sums(costs, offsets)
        n_orig = len(offsets) - 1
        # trajectory table: source slice [lo, hi) of the flat arrays + this trajectory's own return / cost to go
        self._obs, self._act, self._cost, self._rew = obs, act, costs, rew
        traj = [(int(offsets[i]), int(offsets[i + 1]), rtg[offsets[i]:offsets[i + 1]], ctg[offsets[i]:offsets[i + 1]])
                for i in range(n_orig)]
        r0 = np.array([t[2][0] for t in traj], dtype=np.float64)
        c0 = np.array([t[3][0] for t in traj], dtype=np.float64)
        aug = []
        self.pareto_frontier = None
        if pf_only:
            # dataset.py:702-707 selects the best `npb` returns per cost bin (select_optimal_trajectory, :509-547) into
            # self.dataset -- and :725 then overwrites self.dataset with original + augmented unconditionally, so in the
            # reference pf_only only prints its banner and suppresses augmentation.  Reproduced as is.
            print("*" * 100 + "\nUsing pareto frontier data points only!!!!!\n" + "*" * 100)
        elif random_aug > 0:              # dataset.py:550-630
            num = int(random_aug * n_orig)
            tgt = np.random.uniform(low=(aug_cmin, aug_rmin), high=(aug_cmax, aug_rmax), size=(num, 2))
            pts, ids, cmin = np.stack([c0, r0], 1), np.arange(n_orig), np.min(c0)
            srcs = []
            for p in tgt:
                ok = pts[:, 0] <= max(p[0] - cgap, cmin + 1)
                d = pts[ok] - p
                srcs.append(ids[ok][np.argmin(np.hypot(d[:, 0], d[:, 1]))])
            for i, p in zip(srcs, tgt):
                lo, hi, r, c = traj[i]
                c2 = c + (p[0] - c[0] + np.random.normal(loc=0, scale=cstd, size=c.shape)).astype(c.dtype)
                r2 = r + (p[1] - r[0] + np.random.normal(loc=0, scale=rstd, size=r.shape)).astype(r.dtype)
                aug.append((lo, hi, r2, c2))
            self.idx = srcs
        elif augment_percent > 0:         # dataset.py:290-387
            keep = _grid_filter(c0, r0, 10, 50, 10, 2)
            print(f"after filter {len(keep)}")
            fc, fr = c0[keep], r0[keep]
            pf = pareto_front_2d(fc, fr)
            self.pareto_frontier = np.poly1d(np.polyfit(fc[pf], fr[pf], deg=deg))
            num = int(augment_percent * fc.shape[0])
            grid = np.linspace(np.min(fc), np.max(fc), num)
            curve = self.pareto_frontier(grid)
            sampled = np.random.uniform(low=curve + min_reward, high=max_reward * np.ones(curve.shape), size=num)
            tgt = np.stack([grid, sampled], 1)
            srcs = _nearest_sources(np.stack([fc, fr], 1), tgt, max_rew_decrease, beta)
            for i, p in zip(srcs, tgt):
                lo, hi, r, c = traj[keep[i]]
                aug.append((lo, hi, (r + (p[1] - r[0])).astype(r.dtype), (c + (p[0] - c[0])).astype(c.dtype)))
            self.idx, self.indices = srcs, keep
        self._traj = traj + aug
        self.n_original, self.n_augmented = len(traj), len(aug)
        print(f"original data: {len(traj)}, augment data: {len(aug)}, total: {len(self._traj)}")

        first_c = np.array([t[3][0] for t in self._traj])          # float32, like dataset.py:452-458
        first_r = np.array([t[2][0] for t in self._traj])
        self.sample_prob = None
        if cost_sample:
            p = np.array([cost_transform(c) for c in first_c])
            p[p < 0] = 0
            p /= np.sum(p)
            self.sample_prob = p
        elif pf_sample:                   # dataset.py:390-430: ~ 1 / (distance to the frontier curve + 1)
            from scipy.optimize import minimize
            pr = []
            for r, c in zip(first_r.astype(np.float64), first_c.astype(np.float64)):
                f = lambda x: (x - c) ** 2 + (self.pareto_frontier(x) - r) ** 2
                x = np.max([0, minimize(f, x0=c, method="bfgs", tol=1e-4).x[0]])
                pr.append(1 / (np.sqrt(f(x)) + 1))
            self.sample_prob = np.array(pr) / np.sum(pr)
        if start_sampling:                # dataset.py:470-492
            self.start_idx_sample_prob = []
            kern = _gauss_kernel(10, 10)
            for lo, hi, _, _ in self._traj:
                cs = self._cost[lo:hi]
                k, length = np.sum(cs), len(cs)
                x = 100 if prob * length - k <= 0 else k * (1 - prob) / (prob * length - k)
                x = 1 if x <= 0 else x
                w = np.convolve(np.array(cs), kern)[10:-10] + x
                self.start_idx_sample_prob.append(w / w.sum())

    def __len__(self):
        return len(self._traj)

    def compute_pareto_return(self, cost):
        return self.pareto_frontier(cost)

    @property
    def offsets(self):
        """Trajectory boundaries of the materialised (original + augmented) buffer."""
        return np.concatenate([[0], np.cumsum([hi - lo for lo, hi, _, _ in self._traj])]).astype(np.int64)

    @property
    def flat(self):
        cat = lambda parts: np.concatenate(parts, 0)
        return {"observations": cat([self._obs[lo:hi] for lo, hi, _, _ in self._traj]),
                "actions": cat([self._act[lo:hi] for lo, hi, _, _ in self._traj]),
                "costs": cat([self._cost[lo:hi] for lo, hi, _, _ in self._traj]),
                "returns": cat([r for _, _, r, _ in self._traj]), "cost_returns": cat([c for _, _, _, c in self._traj])}

    @staticmethod
    def device_resident(engine, dataset: dict, reward_scale: float = 1.0, cost_scale: float = 1.0,
                        cost_reverse: bool = False, cost_sample: bool = False, cost_transform=lambda x: 50 - x) -> dict:
        """The un-augmented SequenceDataset built entirely on the GPU (osrl_seq_preprocess): the episode split and the
        two discounted_cumsum passes of process_sequence_dataset (dataset.py:137-183) run as kernels on the raw DSRL
        arrays and leave the packed trajectory buffer resident; only the per-episode first returns come back, for the
        cost-based episode distribution (dataset.py:452-459).  Augmentation / Pareto sampling need the host class."""
        info = engine.preprocess_seq_dataset(dataset, reward_scale, cost_scale, cost_reverse)
        if cost_sample:
            p = np.array([cost_transform(c) for c in info["cost_returns"]])
            p[p < 0] = 0
            engine.set_seq_sample_prob(p / np.sum(p))
            info["sample_prob"] = p / np.sum(p)
        return info

    def to_engine(self, engine) -> None:
        """Pack into HBM once (osrl_seq_buffer_upload); windows are then drawn on the device."""
        if self.start_sampling:
            raise NotImplementedError("start_sampling draws non-uniform start indices: use the DataLoader path")
        engine.upload_seq_dataset(dict(self.flat, traj_offsets=self.offsets, sample_prob=self.sample_prob),
                                  self.reward_scale, self.cost_scale)

    def sample(self, traj_idx: int, start_idx: int):
        """dataset.py:749-775 (__prepare_sample): window, float32 scaling, zero padding at the end, unclipped time steps."""
        lo0, hi0, rtg, ctg = self._traj[traj_idx]
        lo, hi = lo0 + start_idx, min(lo0 + start_idx + self.seq_len, hi0)
        n, T = hi - lo, self.seq_len
        pad = lambda x: np.concatenate([x, np.zeros((T - n,) + x.shape[1:], dtype=x.dtype)], 0) if n < T else x
        a, b = start_idx, start_idx + n
        mask = np.hstack([np.ones(n), np.zeros(T - n)])
        return (pad(self._obs[lo:hi]), pad(self._act[lo:hi]), pad(rtg[a:b] * self.reward_scale),
                pad(ctg[a:b] * self.cost_scale), np.arange(start_idx, start_idx + T), mask,
                ctg[0] * self.cost_scale, pad(self._cost[lo:hi]))

    def __iter__(self):
        while True:
            t = np.random.choice(len(self), p=self.sample_prob)
            length = self._traj[t][1] - self._traj[t][0]
            if self.start_sampling:
                s0 = np.random.choice(length, p=self.start_idx_sample_prob[t])
            else:
                s0 = random.randint(0, length - 1)
            yield self.sample(t, s0)
