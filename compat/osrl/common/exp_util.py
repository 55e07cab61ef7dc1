"""osrl.common.exp_util -> osrl_b200.common.exp_util"""
from osrl_b200.common.exp_util import *  # noqa: F401,F403
from osrl_b200.common import exp_util as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
