"""osrl.algorithms.cpq -> osrl_b200.algorithms.cpq"""
from osrl_b200.algorithms import cpq as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
