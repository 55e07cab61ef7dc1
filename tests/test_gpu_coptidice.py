"""GPU parity of the COptiDICE step (csrc/algo_coptidice.cu) against the fixture written by the UNMODIFIED reference
(tests/golden/coptidice_small.npz, oracle/make_golden.py) and against the live oracle at BASELINE layer sizes.
Tolerance policy: tests/test_gpu_parity.py module docstring."""
import numpy as np
import pytest
import torch

from oracle import coptidice as oc
from oracle import synth
from oracle import algos
from tests.helpers import RTOL, load_golden, maxrel, record_margin, to_double

pytestmark = pytest.mark.gpu
KEYS = ("observations", "next_observations", "actions", "rewards", "costs", "done", "is_init")


def _engine(cfg, B, obs_std, act_std, gemm="fz"):
    import os
    from osrl_b200 import Engine
    c = dict(cfg)
    os.environ["OSRL_GEMM"] = gemm
    try:
        return Engine("coptidice", batch_size=B, device=0, seed=5, observations_std=obs_std, actions_std=act_std, **c)
    finally:
        os.environ.pop("OSRL_GEMM", None)


def test_coptidice_small_golden(lib_built):
    z, meta = load_golden("coptidice_small")
    B, steps, cfg = meta["B"], meta["steps"], meta["cfg"]
    eng = _engine(cfg, B, z["observations_std"], z["actions_std"])
    keys = [k for k in meta["keys"] if k not in ("tau", "lmbda")]
    init = {k: torch.from_numpy(z["init/" + k]) for k in keys}
    eng.load_params(init)
    for s in range(steps):
        eng.step({k: z[f"batch{s}/{k}"] for k in KEYS}, {k: z[f"noise{s}/{k}"] for k in eng.noise_layout})
        got = eng.stats()
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            tol = 2e-5 * max(abs(w), 1e-3) + 1e-7
            record_margin("coptidice_golden", "stat " + k, abs(got[k] - w), tol)
            assert abs(got[k] - w) <= tol, f"step {s} {k}: {got[k]} vs reference {w}"
    P = eng.read_params()
    for k in keys:
        ref = torch.from_numpy(z["final/" + k])
        if float((ref - init[k]).abs().max()) > 0:
            err = float((P[k] - ref).norm())
            bound = 5e-4 * float((ref - init[k]).norm()) + 4e-7 * float(ref.norm())
            record_margin("coptidice_golden", "final params after k steps (l2 of delta)", err, bound)
            assert err <= bound, f"{k}: err {err:.3e} > {bound:.3e}"
    sc = eng.scalars()
    assert abs(sc["tau"] - float(z["final/tau"][0])) <= 1e-6 and abs(sc["lmbda"] - float(z["final/lmbda"][0])) <= 1e-6, sc
    eng.close()


@pytest.mark.parametrize("gemm", ["ffma", "fz"])
@pytest.mark.parametrize("f_type", ["softchi", "chi2", "kl"])
def test_coptidice_full_vs_live_oracle(lib_built, f_type, gemm):
    """BASELINE layer sizes (256 x 256, CarCircle dims, B=256): stats, gradients and the dual variables per step."""
    cfg = oc.COptiDICEConfig(8, 2, 1.0, f_type=f_type, init_state_propotion=0.2, a_hidden_sizes=[256, 256],
                             c_hidden_sizes=[256, 256], num_nu=2, num_chi=2, actor_lr=1e-3, critic_lr=1e-3, scalar_lr=1e-3)
    B = 256
    rng = np.random.default_rng(3)
    obs_std = rng.uniform(0.5, 1.5, 8).astype(np.float32)
    act_std = rng.uniform(0.3, 0.8, 2).astype(np.float32)
    torch.manual_seed(0)
    orc = oc.COptiDICEOracle(cfg, obs_std[None], act_std[None])
    import dataclasses
    eng = _engine(dataclasses.asdict(cfg), B, obs_std, act_std, gemm)
    eng.load_params({k: v for k, v in orc.params.items() if k not in ("tau", "lmbda")})
    # exact-fp32 GEMMs reproduce the reference's pre-activations to ~1e-7: kink flips are rare; the tensor-core path
    # is ~1e-6 off
    max_frac, cap = (0.06, 1e-2) if gemm == "ffma" else (0.4, 2e-1)   # one flipped unit of a nu member moves its 4 lower tensors and,
    # through w of the policy phase, the actor's 8: 12 of 34 tensors from a single ReLU flip
    torch.manual_seed(11)
    for s in range(3):
        b = synth.make_batch(rng, B, 8, 2)
        b["is_init"] = (rng.random(B) < 0.2).astype(np.float32)
        if s > 0:   # every step starts from the oracle's exact state: Adam's sign-like first updates amplify 1e-13
            # gradient differences into 1e-5 weight differences, which then flip ReLU units of the NEXT step; the
            # multi-step trajectory is covered by test_coptidice_small_golden
            eng.load_params({k: v for k, v in orc.params.items() if k not in ("tau", "lmbda")})
            eng.set_scalars(tau=float(orc.params["tau"]), lmbda=float(orc.params["lmbda"]))
        o64 = to_double(orc)                      # conditioning probe: the same step from the same state in float64
        o64.obs_std, o64.act_std = o64.obs_std.double(), o64.act_std.double()
        want = orc.step(*[torch.from_numpy(b[k]) for k in KEYS])
        with algos.precision(torch.float64):
            o64.step(*[torch.from_numpy(b[k]).double() for k in KEYS], noise={k: v.double() for k, v in orc.last_noise.items()})
        eng.step(b, {k: v for k, v in orc.last_noise.items() if k in eng.noise_layout})
        got = eng.stats()
        for k, w in want.items():
            tol = 2e-5 * max(abs(w), 1e-3) + 1e-7
            if k in ("loss/D_kl", "loss/tau_loss"):
                # D_kl = mean(w log w - w + 1) over weights of O(1): a ~2e-3 remainder of O(1) terms, so fp32 rounding
                # (1e-7 per term) is ~1e-4 of the value; held to 1e-6 absolute, i.e. 1e-6 of the terms it is made of
                tol = max(tol, 1e-6)
            record_margin("coptidice_vs_oracle/" + f_type, "stat " + k, abs(got[k] - w), tol)
            assert abs(got[k] - w) <= tol, f"{f_type} step {s} {k}: {got[k]} vs {w}"
        G = eng.read_section("grad")
        bad = []
        for k, g in orc.last_grads.items():
            if k in ("tau", "lmbda") or float(g.abs().max()) == 0.0:
                continue
            if g.numel() == 1:
                # head biases: one scalar = sum over the 2B rows of signed output gradients of O(1/B) that largely
                # cancel ((1-gamma) init/p0 - w against gamma (1-done) w): fp32 rounding of the summands, 1e-7 absolute
                aerr = float((G[k] - g).abs().max())
                record_margin("coptidice_vs_oracle/" + f_type, "head-bias grad (abs)", aerr, 1e-6)
                assert aerr <= 1e-6, f"{f_type} step {s} grad {k}: abs err {aerr:.2e}"
                continue
            # w = relu(f'^-1(e / alpha)) and the min over the ensemble are kinks: a row within rounding of one moves a
            # whole back-propagation path by O(1/B); the reference's own fp32-vs-fp64 gap measures that
            cond = maxrel(g, o64.last_grads[k])
            tol = max(2 * RTOL, 10 * cond)
            err = maxrel(G[k], g)
            record_margin("coptidice_vs_oracle/" + f_type, "grad (max-rel)", err, tol)
            if err > tol:
                bad.append((k, err, cond))
                assert err <= cap, f"{f_type}/{gemm} step {s} grad {k}: {err:.2e} (cond {cond:.1e})"
        # (a kink can also separate the engine from BOTH reference evaluations -- its pre-activations carry the
        # 3xTF32 error of ~1e-6: same budget as tests/test_gpu_parity.py for the tensor-core paths)
        record_margin(f"coptidice_vs_oracle/{f_type}/{gemm}", "grad outlier tensors (fraction of budget)", len(bad),
                      max(1, int(max_frac * len(orc.last_grads))))
        assert len(bad) <= max(1, int(max_frac * len(orc.last_grads))), bad[:6]
        sc = eng.scalars()
        assert abs(sc["tau"] - float(orc.params["tau"])) <= 1e-6 and abs(sc["lmbda"] - float(orc.params["lmbda"])) <= 1e-6
    eng.close()


def test_coptidice_trainer_api(lib_built):
    """The public mirror (COptiDICE / COptiDICETrainer, train_coptidice.py:101-143): batch as a 7-list, noise drawn on
    the device, resident dataset with is_init; stats finite, dual variables move."""
    from osrl_b200.algorithms import COptiDICE, COptiDICETrainer
    from osrl_b200.common.dataset import TransitionDataset
    data = synth.make_dataset(8, 2, 30, 40, seed=2)
    ds = TransitionDataset(data, reward_scale=0.1, cost_scale=1.0, state_init=True)
    p0, osd, asd = ds.get_dataset_states()
    torch.manual_seed(0)
    model = COptiDICE(8, 2, 1.0, "softchi", p0, osd, asd, [32, 32], [32, 32], num_nu=2, num_chi=2, device="cuda:0")

    class Log:
        rows = []

        def store(self, tab=None, **kw):
            self.rows.append(kw)

    tr = COptiDICETrainer(model, None, Log(), actor_lr=1e-3, critic_lr=1e-3, scalar_lr=1e-3, reward_scale=0.1, device="cuda:0")
    it = iter(torch.utils.data.DataLoader(ds, batch_size=64))
    for _ in range(5):
        tr.train_one_step([t.to("cuda:0") for t in next(it)])
    assert all(np.isfinite(v) for v in Log.rows[-1].values()) and "loss/nu_loss" in Log.rows[-1]
    assert float(model.tau) != 1.0 and float(model.lmbda) != 1.0
    a, logp = model.act(np.zeros(8, dtype=np.float32), True, True)
    assert a.shape == (2,) and np.isfinite(logp)
    eng = model.engine
    eng.upload_dataset(ds.dataset, 0.1, 1.0)
    eng.steps(3)
    assert all(np.isfinite(v) for v in eng.stats().values())
