"""osrl.algorithms.bcql -> osrl_b200.algorithms.bcql"""
from osrl_b200.algorithms import bcql as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
