"""Experiment glue with the reference's names (osrl/common/exp_util.py): seeding, run naming, checkpoint loading.

None of this is on the hot path; it exists so that the unchanged example scripts (examples/train/*.py:22,37;
examples/eval/*.py:12,28) find the same functions with the same behaviour when `osrl` resolves to this package.
"""
from __future__ import annotations

import os
import os.path as osp
import random
import uuid
from collections.abc import Mapping, Sequence

import numpy as np
import torch
import yaml


def seed_all(seed=1029, others=None):
    """exp_util.py:12-31: the global generators the reference seeds, in the same order."""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    if others is not None:
        if hasattr(others, "seed"):
            others.seed(seed)
            return True
        try:
            for item in others:
                if hasattr(item, "seed"):
                    item.seed(seed)
        except TypeError:
            pass


def get_cfg_value(config, key):
    """exp_util.py:34-48: depth-first lookup of `key`; lists are rendered as their concatenated items."""
    if key in config:
        v = config[key]
        return "".join(str(i) for i in v) if isinstance(v, list) else str(v)
    for v in config.values():
        if isinstance(v, dict):
            found = get_cfg_value(v, key)
            if found is not None:
                return found
    return "None"


def load_config_and_model(path: str, best: bool = False):
    """exp_util.py:51-74: (config.yaml, checkpoint/model[_best].pt) of a finished run directory.  The checkpoint is
    the reference's `{"model_state": state_dict}`; loading it into one of this package's models writes straight into
    the engine arena (the parameters are views)."""
    if not osp.exists(path):
        raise ValueError(f"{path} doesn't exist!")
    config_file = osp.join(path, "config.yaml")
    print(f"load config from {config_file}")
    with open(config_file) as f:
        config = yaml.load(f.read(), Loader=yaml.FullLoader)
    model_path = osp.join(path, "checkpoint", "model_best.pt" if best else "model.pt")
    print(f"load model from {model_path}")
    return config, torch.load(model_path)


def to_string(values) -> str:
    """exp_util.py:77-96: flatten nested sequences / dicts (keys sorted) into an underscore-joined string."""
    if isinstance(values, Sequence) and not isinstance(values, str):
        return "_".join(to_string(v) for v in values)
    if isinstance(values, Mapping):
        return "_".join(to_string(values[k]) for k in sorted(values.keys()))
    return str(values)


DEFAULT_SKIP_KEY = [
    "task", "reward_threshold", "logdir", "worker", "project", "group", "name", "prefix", "suffix", "save_interval",
    "render", "verbose", "save_ckpt", "training_num", "testing_num", "epoch", "device", "thread"
]

DEFAULT_KEY_ABBRE = {
    "cost_limit": "cost", "mstep_iter_num": "mnum", "estep_iter_num": "enum", "estep_kl": "ekl",
    "mstep_kl_mu": "kl_mu", "mstep_kl_std": "kl_std", "mstep_dual_lr": "mlr", "estep_dual_lr": "elr",
    "update_per_step": "update"
}


def auto_name(default_cfg: dict, current_cfg: dict, prefix: str = "", suffix: str = "",
              skip_keys: list = DEFAULT_SKIP_KEY, key_abbre: dict = DEFAULT_KEY_ABBRE) -> str:
    """exp_util.py:117-151: run name = every non-default, non-skipped setting as `<key><value>`, joined by '_',
    between `prefix` and `suffix`, plus a 4-hex-digit tag ("default-xxxx" when nothing differs)."""
    parts = [prefix] if prefix else []
    for k in sorted(default_cfg.keys()):
        if k in skip_keys or default_cfg[k] == current_cfg[k]:
            continue
        parts.append(key_abbre.get(k, k) + to_string(current_cfg[k]))
    name = "_".join(parts)
    if suffix:
        name = f"{name}_{suffix}" if name else suffix
    return f"{name or 'default'}-{str(uuid.uuid4())[:4]}"
