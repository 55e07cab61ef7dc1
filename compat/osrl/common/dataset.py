"""osrl.common.dataset -> osrl_b200.common.dataset"""
from osrl_b200.common.dataset import *  # noqa: F401,F403
from osrl_b200.common import dataset as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
