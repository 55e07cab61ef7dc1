"""Shared helpers of the parity tests (the oracle is the checker, never the product)."""
import dataclasses
import json
import os

import numpy as np
import torch

from oracle import algos

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORACLES = {"bc": (algos.BCOracle, algos.BCConfig), "bcql": (algos.BCQLOracle, algos.BCQLConfig),
           "cpq": (algos.CPQOracle, algos.CPQConfig), "bearl": (algos.BEARLOracle, algos.BEARLConfig)}
# tolerance named by BASELINE.json north_star: losses and parameter deltas within 1e-5 relative
RTOL = 1e-5


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def make_oracle(algo, cfg_dict, init_seed=0):
    cls, cfgcls = ORACLES[algo]
    torch.manual_seed(init_seed)
    return cls(cfgcls(**cfg_dict))


def engine_kwargs(algo, cfg_dict):
    return dict(cfg_dict)


def batch_tuple(algo, b):
    t = {k: torch.as_tensor(np.asarray(v)) for k, v in b.items()}
    if algo == "bc":
        return (t["observations"], t["actions"])
    return (t["observations"], t["next_observations"], t["actions"], t["rewards"], t["costs"], t["done"])


def rel_delta_err(p_eng, p_ref, p_init):
    """max |p_eng - p_ref| relative to the largest parameter delta of that tensor."""
    d = (p_ref - p_init).abs().max().item()
    return (p_eng - p_ref).abs().max().item() / (d + 1e-12)
