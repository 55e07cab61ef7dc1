"""Import the UNMODIFIED reference from /root/reference inside the build container.

Test infrastructure only, and only usable where /root/reference exists (it does not
on the GPU box).  The reference's algorithm modules import ``gymnasium`` (for a type
annotation) and ``fsrl.utils`` (logger classes) which are not installed here; two
stub modules are injected so ``import osrl.algorithms`` succeeds.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OSRL_REFERENCE_ROOT", "/root/reference")


class _NullLogger:
    def __init__(self, *a, **k):
        self.rows = []

    def store(self, tab=None, **kw):
        self.rows.append(dict(kw))

    def write(self, *a, **k):
        pass

    write_without_reset = save_config = setup_checkpoint_fn = save_checkpoint = write


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "osrl"))


def import_reference():
    """Returns the reference's ``osrl`` package (algorithms + common)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "gymnasium" not in sys.modules:
        try:
            import gymnasium  # noqa: F401
        except Exception:
            g = types.ModuleType("gymnasium")
            g.Env = object
            sys.modules["gymnasium"] = g
    if "fsrl" not in sys.modules:
        f = types.ModuleType("fsrl")
        fu = types.ModuleType("fsrl.utils")
        fu.DummyLogger = _NullLogger
        fu.WandbLogger = _NullLogger
        f.utils = fu
        sys.modules["fsrl"] = f
        sys.modules["fsrl.utils"] = fu
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own repo must not shadow the reference's top-level ``osrl`` package
    for name in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "") or REFERENCE_ROOT not in (mod.__file__ or ""):
            del sys.modules[name]
    import osrl.algorithms  # noqa: F401
    import osrl.common  # noqa: F401
    import osrl
    return osrl


NullLogger = _NullLogger
