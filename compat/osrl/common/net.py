"""osrl.common.net -> osrl_b200.common.net"""
from osrl_b200.common.net import *  # noqa: F401,F403
from osrl_b200.common import net as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
