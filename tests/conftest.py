import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    """The in-tree shared library; built on demand when nvcc is present."""
    from osrl_b200 import _lib, build
    try:
        build.build()          # no-op when the .so is newer than every source / header (build._stale)
    except RuntimeError:
        if not os.path.exists(_lib.LIB_PATH):   # no nvcc here: a prebuilt library is fine, none is not
            raise
    return _lib.load()


def pytest_sessionfinish(session, exitstatus):
    """Dump the measured parity margins of a GPU session (tests/helpers.py: record_margin)."""
    import json
    from tests import helpers
    if not helpers.MARGINS:
        return
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    rows = dict(sorted(helpers.MARGINS.items(), key=lambda kv: -kv[1]["worst_ratio"]))
    with open(os.path.join(out, "parity_margins.json"), "w") as f:
        json.dump({"exitstatus": int(exitstatus), "margins": rows}, f, indent=1)
