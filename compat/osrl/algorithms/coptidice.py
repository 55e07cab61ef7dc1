"""osrl.algorithms.coptidice -> osrl_b200.algorithms.coptidice"""
from osrl_b200.algorithms import coptidice as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
