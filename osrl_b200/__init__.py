"""osrl_b200 -- B200-native engine for the OSRL per-step training hot path.

The arithmetic lives in ``libosrl_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/osrl_b200.h``); this package is the host-side mirror of the reference's Python
surface (``osrl.algorithms.*``, ``osrl.common.*``).  There is no CPU / PyTorch fallback.
"""
from .engine import Engine, comm_unique_id, make_config, plan  # noqa: F401

__version__ = "0.1.0"
