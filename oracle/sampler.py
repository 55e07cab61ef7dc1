"""Oracle restatement of the minibatch samplers (test infrastructure only).

* ``transition_sample`` -- TransitionDataset.__prepare_sample (dataset.py:832-842): numpy fancy
  indexing of six fields, float32 reward/cost scaling.
* ``philox_indices`` -- the engine's on-device index draw (Philox4x32-10, kernels.cuh draw_index)
  restated in numpy so the sampled path is bit-checkable.  The reference itself draws with
  numpy's MT19937 per DataLoader worker (dataset.py:844-847), which no other generator can
  reproduce stream-for-stream; uniform-with-replacement is the property that is kept.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & _MASK).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_indices(seed: int, step: int, rank: int, rows: int, n: int) -> np.ndarray:
    i = np.arange(rows, dtype=np.uint32)
    c1 = np.full(rows, step & 0xFFFFFFFF, dtype=np.uint32)
    c2 = np.full(rows, 1 | (((step >> 32) & 0xFFFFFF) << 8), dtype=np.uint32)
    c3 = np.full(rows, rank, dtype=np.uint32)
    x0, _, _, _ = philox4x32_10(i, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return ((x0.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def transition_sample(dataset: dict, idx, reward_scale: float = 1.0, cost_scale: float = 1.0):
    done = np.logical_or(dataset["terminals"], dataset["timeouts"]).astype(np.float32)  # dataset.py:815-816
    return (dataset["observations"][idx, :], dataset["next_observations"][idx, :], dataset["actions"][idx, :],
            dataset["rewards"][idx] * reward_scale, dataset["costs"][idx] * cost_scale, done[idx])


def philox_sequences(seed: int, step: int, rank: int, rows: int, prob: np.ndarray, alias: np.ndarray,
                     traj_offsets: np.ndarray):
    """The engine's on-device trajectory draw (cdt_kernels.cuh draw_sequence) restated: alias-table categorical
    over trajectories + uniform start index, both from one Philox4x32-10 block per batch row."""
    n_traj = prob.shape[0]
    i = np.arange(rows, dtype=np.uint32)
    c1 = np.full(rows, step & 0xFFFFFFFF, dtype=np.uint32)
    c2 = np.full(rows, 2 | (((step >> 32) & 0xFFFFFF) << 8), dtype=np.uint32)
    c3 = np.full(rows, rank, dtype=np.uint32)
    x0, x1, x2, _ = philox4x32_10(i, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    slot = ((x0.astype(np.uint64) * np.uint64(n_traj)) >> np.uint64(32)).astype(np.int64)
    u = (x1 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    traj = np.where(u < prob[slot].astype(np.float32), slot, alias[slot].astype(np.int64))
    lens = (traj_offsets[traj + 1] - traj_offsets[traj]).astype(np.uint64)
    start = ((x2.astype(np.uint64) * lens) >> np.uint64(32)).astype(np.int64)
    return traj.astype(np.int32), start.astype(np.int32)
