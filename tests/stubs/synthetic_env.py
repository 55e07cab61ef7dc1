"""Seeded synthetic stand-in for a DSRL offline environment (the API of SURVEY.md Appendix D)."""
import numpy as np


class _Box:
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = np.full(shape, low, np.float32), np.full(shape, high, np.float32), shape

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)


class SyntheticOfflineEnv:
    """obs ~ N(0,1), reward = -|a|^2 + noise, cost ~ Bernoulli(0.1); episodes of `episode_len` steps."""

    def __init__(self, obs_dim=8, act_dim=2, episode_len=30, episodes=40, seed=0):
        self.observation_space = _Box(-np.inf, np.inf, (obs_dim,))
        self.action_space = _Box(-1.0, 1.0, (act_dim,))
        self.obs_dim, self.act_dim, self.episode_len, self.episodes = obs_dim, act_dim, episode_len, episodes
        self.rng = np.random.default_rng(seed)
        self.target_cost, self._t = None, 0

    def get_dataset(self):
        import os
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        from oracle import synth
        return synth.make_dataset(self.obs_dim, self.act_dim, self.episode_len, self.episodes, seed=1)

    def set_target_cost(self, c):
        self.target_cost = c

    def pre_process_data(self, data, *a, **k):
        return data

    def get_normalized_score(self, ret, cost):
        return ret, cost

    def reset(self, **k):
        self._t = 0
        return self.rng.standard_normal(self.obs_dim).astype(np.float32), {}

    def step(self, a):
        self._t += 1
        a = np.asarray(a, dtype=np.float32)
        assert a.shape == (self.act_dim,) and np.all(np.isfinite(a))
        obs = self.rng.standard_normal(self.obs_dim).astype(np.float32)
        return obs, float(1.0 - np.sum(a * a)), False, self._t >= self.episode_len, {"cost": float(self.rng.random() < 0.1)}


def make(task, **k):
    dims = {"Ant": (33, 8), "Cheetah": (17, 6)}
    o, a = next((v for key, v in dims.items() if key in task), (8, 2))
    return SyntheticOfflineEnv(o, a)
