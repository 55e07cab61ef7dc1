"""osrl.algorithms.bc -> osrl_b200.algorithms.bc"""
from osrl_b200.algorithms import bc as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
