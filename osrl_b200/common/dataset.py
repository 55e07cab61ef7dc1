"""Minibatch sources with the reference's class names (osrl/common/dataset.py:633-847).

``TransitionDataset`` keeps the reference's constructor and iterator contract (so it can be
handed to ``torch.utils.data.DataLoader`` by the unchanged example scripts), and additionally
knows how to make itself resident in HBM: ``trainer.set_dataset(ds.dataset, ...)`` /
``ds.to_engine(engine)`` packs it once and all later draws happen on the device
(``osrl_steps``).
"""
from __future__ import annotations

import numpy as np
from torch.utils.data import IterableDataset


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


class SequenceDataset(IterableDataset):
    """Placeholder for the CDT trajectory sampler (dataset.py:633-787); built with the CDT path."""

    def __init__(self, *a, **k):
        raise NotImplementedError("SequenceDataset is part of the CDT path, which this build does not include yet")
