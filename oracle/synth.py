"""Synthetic DSRL-shaped data.  Test/bench infrastructure only.

The DSRL package (datasets + envs) is not vendored and not installable offline, so
every measurement and parity test uses seeded synthetic arrays with DSRL's shapes and
dtypes: float32 observations / next_observations / actions / rewards / costs and
bool-like terminals / timeouts.
"""
from __future__ import annotations

import numpy as np

TASK_DIMS = {  # obs, act, episode_len, episodes
    "OfflineCarCircle-v0": (8, 2, 300, 1450),
    "OfflineAntRun-v0": (33, 8, 200, 1816),
    "OfflinePointCircle1Gymnasium-v0": (28, 2, 500, 1000),
    "OfflineHalfCheetahVelocityGymnasium-v1": (17, 6, 1000, 1000),
}


def make_dataset(obs_dim: int, act_dim: int, episode_len: int, episodes: int, seed: int = 0) -> dict:
    rng = np.random.default_rng(seed)
    n = episode_len * episodes
    obs_all = rng.standard_normal((n + 1, obs_dim), dtype=np.float32)
    timeouts = np.zeros(n, dtype=bool)
    timeouts[episode_len - 1::episode_len] = True
    nxt = obs_all[1:].copy()
    return {
        "observations": obs_all[:-1].copy(),
        "next_observations": nxt,
        "actions": rng.uniform(-1.0, 1.0, (n, act_dim)).astype(np.float32),
        "rewards": (0.5 + 0.5 * rng.standard_normal(n)).astype(np.float32),
        "costs": (rng.random(n) < 0.1).astype(np.float32),
        "terminals": np.zeros(n, dtype=bool),
        "timeouts": timeouts,
    }


def make_task_dataset(task: str, seed: int = 0, episodes: int | None = None) -> dict:
    o, a, T, E = TASK_DIMS[task]
    return make_dataset(o, a, T, E if episodes is None else episodes, seed)


def make_batch(rng: np.random.Generator, B: int, obs_dim: int, act_dim: int) -> dict:
    """One already-scaled transition minibatch (what TransitionDataset yields, collated)."""
    return {
        "observations": rng.standard_normal((B, obs_dim), dtype=np.float32),
        "next_observations": rng.standard_normal((B, obs_dim), dtype=np.float32),
        "actions": rng.uniform(-1.0, 1.0, (B, act_dim)).astype(np.float32),
        "rewards": (0.1 * (0.5 + 0.5 * rng.standard_normal(B))).astype(np.float32),
        "costs": (rng.random(B) < 0.1).astype(np.float32),
        "done": (rng.random(B) < 0.02).astype(np.float32),
    }
