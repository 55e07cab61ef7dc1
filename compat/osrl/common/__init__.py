"""osrl.common -> osrl_b200.common (reference: osrl/common/__init__.py)."""
from osrl.common.dataset import SequenceDataset, TransitionDataset  # noqa: F401
from osrl.common.exp_util import *  # noqa: F401,F403
from osrl.common.net import *  # noqa: F401,F403
