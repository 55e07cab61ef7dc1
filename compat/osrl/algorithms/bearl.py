"""osrl.algorithms.bearl -> osrl_b200.algorithms.bearl"""
from osrl_b200.algorithms import bearl as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
