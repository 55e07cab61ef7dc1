"""Shared helpers of the parity tests (the oracle is the checker, never the product)."""
import copy
import json
import os

import numpy as np
import torch

from oracle import algos

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ORACLES = {"bc": (algos.BCOracle, algos.BCConfig), "bcql": (algos.BCQLOracle, algos.BCQLConfig),
           "cpq": (algos.CPQOracle, algos.CPQConfig), "bearl": (algos.BEARLOracle, algos.BEARLConfig)}
# tolerance named by BASELINE.json north_star: losses and parameter deltas within 1e-5 relative
RTOL = 1e-5


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def make_oracle(algo, cfg_dict, init_seed=0):
    cls, cfgcls = ORACLES[algo]
    torch.manual_seed(init_seed)
    return cls(cfgcls(**cfg_dict))


def batch_tuple(algo, b, dtype=torch.float32):
    t = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in b.items()}
    if algo == "bc":
        return (t["observations"], t["actions"])
    return (t["observations"], t["next_observations"], t["actions"], t["rewards"], t["costs"], t["done"])


def to_double(orc):
    """float64 twin of an oracle at its current state (conditioning probe)."""
    o = copy.deepcopy(orc)
    for k in list(o.params):
        o.params[k] = o.params[k].detach().double()
    for st in (o.opt.values() if isinstance(o.opt, dict) else [o.opt]):
        st.m = {k: v.double() for k, v in st.m.items()}
        st.v = {k: v.double() for k, v in st.v.items()}
    if hasattr(o, "pid"):
        o.pid.e_old, o.pid.e_int = o.pid.e_old.double(), o.pid.e_int.double()
    if hasattr(o, "log_alpha"):
        o.log_alpha = o.log_alpha.double()
    return o


def probe_step(orc, algo, batch, noise=None):
    """One reference step in fp32 (the reference path) and one in fp64 from the same state with the same
    noise.  Returns (stats32, stats64, grads32, grads64, params_before, params64_after).  The fp32/fp64 gap
    measures how well-conditioned each quantity is: a ReLU unit whose pre-activation rounds to +-0, or an
    Adam update with |g| ~ eps, makes the reference itself irreproducible at 1e-5."""
    o64 = to_double(orc)
    before = {k: v.detach().clone() for k, v in orc.params.items()}
    s32 = orc.step(*batch_tuple(algo, batch), noise=noise)
    nz = {k: v.double() for k, v in orc.last_noise.items()}
    with algos.precision(torch.float64):
        s64 = o64.step(*batch_tuple(algo, batch, torch.float64), noise=nz)
    g32 = {k: v.detach() for k, v in orc.last_grads.items()}
    g64 = {k: v.detach() for k, v in o64.last_grads.items()}
    return s32, s64, g32, g64, before, {k: v.detach() for k, v in o64.params.items()}


def maxrel(a, b):
    """max |a-b| / max |b|"""
    return float((a.double() - b.double()).abs().max()) / (float(b.double().abs().max()) + 1e-300)


def l2rel(a, b):
    return float((a.double() - b.double()).norm()) / (float(b.double().norm()) + 1e-300)


# ---- measured parity margins ("print and commit the measured worst error per quantity").  Every GPU
# comparison reports (test, quantity, err, tol); conftest writes the per-quantity worst err / tol of the session to
# gpurun_out/parity_margins.json, which is committed as profiles/r02_parity_margins.json.
MARGINS = {}


def record_margin(test: str, quantity: str, err: float, tol: float):
    key = f"{test}::{quantity}"
    cur = MARGINS.get(key)
    ratio = float(err) / float(tol) if tol > 0 else float("inf")
    if cur is None or ratio > cur["worst_ratio"]:
        MARGINS[key] = {"worst_ratio": ratio, "err": float(err), "tol": float(tol), "n": (cur["n"] if cur else 0) + 1}
    else:
        cur["n"] += 1
