"""Constructor / function signatures of the mirrors against the reference's: same parameter names in the same order
with the same default values (so that code relying on positional arguments or on defaults behaves identically); the
mirrors may only append keyword parameters of their own.  `device` defaults differ on purpose (the reference defaults
to "cpu", this package has no CPU path) and logger defaults are instances.  Skipped where /root/reference is absent."""
import inspect
import sys

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not available")


@pytest.fixture(scope="module")
def ref():
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs"))
    import oapackage  # noqa: F401
    ref_shim.import_reference()
    import osrl.algorithms as ra
    import osrl.common.dataset as rd
    import osrl.common.exp_util as re_
    yield {"algorithms": ra, "dataset": rd, "exp_util": re_}
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(a, b)
    if callable(a) and callable(b) and getattr(a, "__name__", "") == "<lambda>":
        return all(a(x) == b(x) for x in (0.0, 1.0, 37.5))     # SequenceDataset's cost_transform default
    return a == b and type(a) is type(b)


def _check(name, theirs, ours, skip=("device", "logger")):
    pt = list(inspect.signature(theirs).parameters.values())
    po = list(inspect.signature(ours).parameters.values())
    pt = [p for p in pt if p.name != "self"]
    po = [p for p in po if p.name != "self"]
    assert [p.name for p in po[:len(pt)]] == [p.name for p in pt], f"{name}: parameter names / order"
    for t, o in zip(pt, po):
        if t.name in skip:
            continue
        if t.default is inspect.Parameter.empty:
            continue    # required in the reference: the mirror may relax it (env=None for the device-resident path)
        assert o.default is not inspect.Parameter.empty, f"{name}.{t.name}: default dropped"
        if True:
            assert _same(t.default, o.default), f"{name}.{t.name}: default {o.default!r} vs reference {t.default!r}"
    for extra in po[len(pt):]:
        assert extra.kind in (extra.VAR_KEYWORD, extra.KEYWORD_ONLY) or extra.default is not inspect.Parameter.empty, \
            f"{name}: extra parameter {extra.name} needs a default"


ALGOS = ["BC", "BCQL", "CPQ", "BEARL", "CDT", "COptiDICE"]


@pytest.mark.parametrize("name", ALGOS + [a + "Trainer" for a in ALGOS])
def test_algorithm_signatures(ref, name):
    import osrl_b200.algorithms as mine
    _check(name, getattr(ref["algorithms"], name).__init__, getattr(mine, name).__init__)


@pytest.mark.parametrize("name", ["TransitionDataset", "SequenceDataset"])
def test_dataset_signatures(ref, name):
    import osrl_b200.common.dataset as mine
    _check(name, getattr(ref["dataset"], name).__init__, getattr(mine, name).__init__)


@pytest.mark.parametrize("name", ["process_bc_dataset", "seed_all", "get_cfg_value", "load_config_and_model", "to_string",
                                  "auto_name"])
def test_function_signatures(ref, name):
    import osrl_b200.common.dataset as md
    import osrl_b200.common.exp_util as me
    theirs = getattr(ref["dataset"], name, None) or getattr(ref["exp_util"], name)
    ours = getattr(md, name, None) or getattr(me, name)
    _check(name, theirs, ours)


@pytest.mark.parametrize("name", ALGOS)
def test_trainer_methods_exist_with_reference_signatures(ref, name):
    import osrl_b200.algorithms as mine
    t, o = getattr(ref["algorithms"], name + "Trainer"), getattr(mine, name + "Trainer")
    for meth in ("train_one_step", "evaluate", "rollout"):
        _check(f"{name}Trainer.{meth}", getattr(t, meth), getattr(o, meth))
