"""osrl_b200.common.exp_util against the UNMODIFIED reference module (osrl/common/exp_util.py) on the same inputs: run
naming (auto_name / to_string incl. abbreviations, skip keys, prefix / suffix, the "default" fallback), the depth-first
config lookup with its quirk (a nested dict that lacks the key answers the string "None" and ends the search), and the
generator streams seed_all leaves behind.  Skipped where /root/reference is absent."""
import random
import sys
import uuid

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not available")


@pytest.fixture(scope="module")
def ref():
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]
    ref_shim.import_reference()
    import osrl.common.exp_util as r
    yield r
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]


CASES = [
    (dict(a=1, b=[1, 2], task="x", cost_limit=10), dict(a=2, b=[1, 2], task="y", cost_limit=20), "BCQL", ""),
    (dict(a=1, b=[1, 2]), dict(a=1, b=[1, 2]), "", ""),
    (dict(a=1, b=[1, 2]), dict(a=1, b=[1, 2]), "", "tail"),
    (dict(a=1, b=[1, [2, 3]], c={"z": 1, "y": (2, 3)}), dict(a=1, b=[4, [5, 6]], c={"z": 7, "y": (8, 9)}), "P", "S"),
    (dict(update_per_step=1, seed=0, device="cpu"), dict(update_per_step=4, seed=3, device="cuda"), "", ""),
]


@pytest.mark.parametrize("default,current,prefix,suffix", CASES)
def test_auto_name(ref, monkeypatch, default, current, prefix, suffix):
    from osrl_b200.common import exp_util as mine
    monkeypatch.setattr(uuid, "uuid4", lambda: "abcd-ef")
    assert mine.auto_name(default, current, prefix, suffix) == ref.auto_name(default, current, prefix, suffix)
    assert mine.DEFAULT_SKIP_KEY == ref.DEFAULT_SKIP_KEY and mine.DEFAULT_KEY_ABBRE == ref.DEFAULT_KEY_ABBRE


def test_to_string_and_get_cfg_value(ref):
    from osrl_b200.common import exp_util as mine
    for v in (3, "s", [1, [2, "x"], (3, 4)], {"b": 1, "a": [2, 3]}, [], {}):
        assert mine.to_string(v) == ref.to_string(v)
    cfg = {"a": 1, "l": [1, 2, 3], "n1": {"x": 5}, "n2": {"y": {"deep": [7, 8]}, "x": 6}}
    for key in ("a", "l", "x", "deep", "y", "missing"):
        assert mine.get_cfg_value(cfg, key) == ref.get_cfg_value(cfg, key), key


def test_seed_all_streams(ref):
    from osrl_b200.common import exp_util as mine

    class Env:
        def seed(self, s):
            self.s = s

    def draw(fn):
        e = Env()
        fn(123, others=[e])
        return random.random(), np.random.rand(3).tolist(), torch.rand(3).tolist(), e.s

    assert draw(mine.seed_all) == draw(ref.seed_all)
    e1, e2 = Env(), Env()
    assert mine.seed_all(5, others=e1) == ref.seed_all(5, others=e2) and e1.s == e2.s == 5


def test_load_config_and_model_roundtrip(ref, tmp_path):
    from osrl_b200.common import exp_util as mine
    import yaml
    (tmp_path / "checkpoint").mkdir()
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.dump({"task": "t", "seed": 3}, f)
    torch.save({"model_state": {"w": torch.arange(3.0)}}, tmp_path / "checkpoint" / "model.pt")
    torch.save({"model_state": {"w": torch.arange(4.0)}}, tmp_path / "checkpoint" / "model_best.pt")
    for best in (False, True):
        c1, m1 = mine.load_config_and_model(str(tmp_path), best)
        c2, m2 = ref.load_config_and_model(str(tmp_path), best)
        assert c1 == c2 and torch.equal(m1["model_state"]["w"], m2["model_state"]["w"])
    with pytest.raises(ValueError):
        mine.load_config_and_model(str(tmp_path / "nope"))
