"""COptiDICE (the next algorithm for the CUDA engine): the CPU oracle restatement
reproduces the fixture the UNMODIFIED reference produced (oracle/make_golden.py coptidice) -- stats per step, final
networks and the softplus scalars -- on the stored batches and noise.  No GPU involved."""
import json
import os

import numpy as np
import torch

from oracle import coptidice as oc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coptidice_small.npz")


def test_coptidice_oracle_reproduces_reference_fixture():
    z = np.load(GOLD, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    cfg = oc.COptiDICEConfig(**meta["cfg"])
    torch.manual_seed(123)            # init is overwritten below; noise comes from the fixture
    orc = oc.COptiDICEOracle(cfg, z["observations_std"], z["actions_std"])
    for k in meta["keys"]:
        orc.params[k] = torch.from_numpy(z["init/" + k]).clone()
    orc.opt = {n: type(o)(o.names, orc.params, o.lr) for n, o in orc.opt.items()}
    keys = ("observations", "next_observations", "actions", "rewards", "costs", "done", "is_init")
    for s in range(meta["steps"]):
        args = [torch.from_numpy(z[f"batch{s}/{k}"]) for k in keys]
        noise = {k: torch.from_numpy(z[f"noise{s}/{k}"]) for k in ("obs_eps", "act_eps", "pi")}
        got = orc.step(*args, noise=noise)
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            assert abs(got[k] - w) <= 2e-5 * max(abs(w), 1e-3) + 1e-7, (s, k, got[k], w)
    for k in meta["keys"]:
        ref = torch.from_numpy(z["final/" + k])
        err = float((orc.params[k] - ref).abs().max())
        assert err <= 2e-5 * float(ref.abs().max()) + 1e-7, (k, err)


def test_f_divergences_match_their_definitions():
    x = torch.linspace(0.05, 3.0, 50)
    for name in ("chi2", "softchi", "kl"):
        f, finv = oc.f_div(name)
        xs = x.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(f(xs).sum(), xs)
        # f'^{-1}(f'(x)) = x wherever f' is invertible (softchi: everywhere on x > 0)
        assert torch.allclose(finv(g), x, atol=2e-4), name
