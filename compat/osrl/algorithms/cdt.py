"""osrl.algorithms.cdt -> osrl_b200.algorithms.cdt"""
from osrl_b200.algorithms import cdt as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
