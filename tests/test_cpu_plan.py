"""CPU-only checks: the C-ABI library loads, exports every declared symbol, and its parameter
plan (names / shapes / order) equals the reference's state_dict as pinned in tests/golden."""
import ctypes
import re
import os

import numpy as np
import pytest

from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "osrl_b200.h")).read()
    declared = set(re.findall(r"\b(osrl_[a-z_]+)\s*\(", hdr))
    from osrl_b200 import _lib
    bound = {s[0] for s in _lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib_built, name)
    assert lib_built.osrl_abi_version() == int(re.search(r"#define\s+OSRL_ABI_VERSION\s+(\d+)", hdr).group(1)) == 3


@pytest.mark.parametrize("case", ["bc_small", "bcql_small", "cpq_small", "bearl_small", "bcql_full", "cdt_small"])
def test_plan_matches_reference_state_dict(lib_built, case):
    from osrl_b200 import plan
    z, meta = load_golden(case)
    cfg = dict(meta["cfg"])
    if meta["algo"] == "cdt":
        cfg["target_entropy"] = -float(cfg["action_dim"])
    table = plan(meta["algo"], **cfg)
    names = [t[0] for t in table]
    assert names == meta["keys"]
    if meta["full"]:
        for name, shape, _, _ in table:
            assert tuple(z["init/" + name].shape) == tuple(shape), name


def test_plan_rejects_bad_config(lib_built):
    from osrl_b200 import plan
    with pytest.raises(RuntimeError):
        plan("bcql", state_dim=0, action_dim=2, a_hidden_sizes=[8], c_hidden_sizes=[8], vae_hidden_sizes=8)


def test_engine_create_fails_loudly_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from osrl_b200 import Engine
    with pytest.raises(RuntimeError, match="CUDA"):
        Engine("bc", batch_size=4, state_dim=3, action_dim=2, a_hidden_sizes=[8, 8])
