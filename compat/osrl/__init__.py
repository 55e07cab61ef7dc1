"""Drop-in alias: ``import osrl`` resolves to the B200 engine's mirrors of the reference package.

Put ``<repo>/compat`` (and ``<repo>``) ahead of the reference checkout on ``PYTHONPATH`` and the unchanged
``examples/train/*.py`` / ``examples/eval/*.py`` import ``osrl.algorithms`` / ``osrl.common`` from here:
same class names, constructor signatures, ``train_one_step`` arguments, ``state_dict`` keys and logger keys
(osrl/__init__.py, osrl/algorithms/__init__.py, osrl/common/__init__.py of the reference); the per-step
arithmetic runs in libosrl_b200.so.  See INTEGRATION.md.
"""
__version__ = "0.1.0"

__all__ = ["algorithms", "common"]
