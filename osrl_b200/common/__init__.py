from .dataset import SequenceDataset, TransitionDataset  # noqa: F401
from .exp_util import seed_all  # noqa: F401
from .net import *  # noqa: F401,F403
