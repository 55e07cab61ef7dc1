"""GPU parity of the CDT step (transformer fwd/bwd, masked losses, grad-norm clip, AdamW warm-up,
temperature Adam) against the oracle and the reference-produced fixtures.  Dropout: the multipliers are replayed
(fixture -> engine, or engine -> oracle), SURVEY.md 7.5-1.
Tolerance policy: tests/test_gpu_parity.py module docstring."""
import copy
import json

import numpy as np
import pytest
import torch

from oracle import algos
from oracle import cdt as ocdt
from oracle.make_golden import CDT_KEYS, make_seq_batch
from tests.helpers import RTOL, l2rel, load_golden, maxrel, record_margin

pytestmark = pytest.mark.gpu


def _cfg(meta):
    c = dict(meta["cfg"])
    c["betas"] = tuple(c["betas"])
    return ocdt.CDTConfig(**c)


def _engine(meta, B, gemm="fz"):
    import os
    from osrl_b200 import Engine
    os.environ["OSRL_GEMM"] = gemm
    c = dict(meta["cfg"])
    c.pop("max_action")
    c["target_entropy"] = -float(c["action_dim"]) if c["target_entropy"] is None else c["target_entropy"]
    try:
        return Engine("cdt", batch_size=B, device=0, seed=3, max_action=1.0, use_rew=1, use_cost=1, cost_transform=1,
                      stochastic=1, **c)
    finally:
        os.environ.pop("OSRL_GEMM", None)


def _args(b, dtype=torch.float32):
    out = []
    for k in CDT_KEYS:
        t = torch.from_numpy(np.asarray(b[k]))
        out.append(t if k == "time_steps" else (t.to(dtype) if k != "mask" else t.to(torch.float64)))
    return out


def _to_double(orc):
    o = copy.deepcopy(orc)
    o.params = type(o.params)((k, v.double()) for k, v in o.params.items())
    o.opt.m = {k: v.double() for k, v in o.opt.m.items()}
    o.opt.v = {k: v.double() for k, v in o.opt.v.items()}
    return o


def test_cdt_small_golden(lib_built):
    z, meta = load_golden("cdt_small")
    B, steps = meta["B"], meta["steps"]
    eng = _engine(meta, B)
    init = {k: torch.from_numpy(z["init/" + k]) for k in meta["keys"]}
    eng.load_params(init)
    for s in range(steps):
        eng.step_seq({k: z[f"batch{s}/{k}"] for k in CDT_KEYS})
        got = eng.stats()
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            assert abs(got[k] - w) <= 2e-5 * max(abs(w), 1e-3) + 1e-7, f"step {s} {k}: {got[k]} vs reference {w}"
    got = eng.read_params()
    for k in meta["keys"]:
        if "in_proj_bias" in k:
            continue   # key-bias slice: exactly-zero true gradient, Adam amplifies rounding noise (make_golden.py)
        ref = torch.from_numpy(z["final/" + k])
        err = float((got[k] - ref).norm())
        bound = 1e-3 * float((ref - init[k]).norm()) + 4e-7 * float(ref.norm()) + 1e-9
        assert err <= bound, f"{k}: err {err:.3e} > {bound:.3e}"
    assert abs(eng.scalars()["log_temperature"] - float(z["log_temperature"])) < 1e-6
    eng.close()


@pytest.mark.parametrize("gemm", ["ffma", "mma", "tc5", "fz"])
@pytest.mark.parametrize("case", ["cdt_small", "cdt_full"])
def test_cdt_against_live_oracle(lib_built, case, gemm):
    z, meta = load_golden(case)
    B, steps = meta["B"], meta["steps"]
    cfg = _cfg(meta)
    torch.manual_seed(0)
    orc = ocdt.CDTOracle(cfg)
    eng = _engine(meta, B, gemm)
    eng.load_params(orc.params)
    rng = np.random.default_rng(77)
    strict = total = 0
    # measured on B200 (profiles/r02_parity_margins.json): 239 gradient tensors per GEMM mode, worst max-rel 4.8e-6
    # (fused tcgen05 path), no tensor above the 2e-5 bound -> at most one outlier per step, and then below 1e-4
    max_frac, cap = 0.02, 1e-4
    for s in range(steps):
        b = make_seq_batch(rng, B, cfg.seq_len, cfg.state_dim, cfg.action_dim)
        o64 = _to_double(orc)
        before = {k: v.clone() for k, v in orc.params.items()}
        s32 = orc.step(*_args(b))
        with algos.precision(torch.float64):
            s64 = o64.step(*_args(b, torch.float64))
        eng.step_seq(b)
        got = eng.stats()
        for k, w in s32.items():
            scale = max(abs(s64[k]), 1e-3)
            cond = abs(w - s64[k]) / scale
            tol = max(RTOL, 10 * cond)
            total += 1; strict += tol == RTOL
            assert abs(got[k] - w) <= tol * scale + 1e-7, f"{case} step {s} stat {k}: {got[k]} vs {w} (cond {cond:.1e})"
        # pinned reference stats
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            assert abs(got[k] - w) <= 2e-5 * max(abs(w), 1e-3) + 1e-7, f"{case} step {s} {k} vs golden"
        G = eng.read_section("grad")
        bad = []
        for k, g in orc.last_grads.items():
            g64 = o64.last_grads[k]
            if float(g64.abs().max()) == 0.0:
                assert float(G[k].abs().max()) <= 1e-12, k
                continue
            cond = maxrel(g, g64)
            tol = max(2 * RTOL, 10 * cond)
            total += 1; strict += tol == 2 * RTOL
            err = maxrel(G[k], g)
            record_margin(f"cdt_vs_oracle/{gemm}", "grad (max-rel, all tensors)", err, tol)
            if err > tol:
                bad.append((k, err, cond))
                assert err <= cap, f"{case}/{gemm} step {s} grad {k}: {err:.2e} (cond {cond:.1e})"
        record_margin(f"cdt_vs_oracle/{gemm}", "grad outlier tensors (count)", len(bad), max(1, int(max_frac * len(orc.last_grads))))
        assert len(bad) <= max(1, int(max_frac * len(orc.last_grads))), bad[:6]
        if s == 0:
            P = eng.read_params()
            for k, ref in orc.params.items():
                d_ref = ref - before[k]
                if float(d_ref.abs().max()) == 0 or "in_proj_bias" in k:
                    continue
                cond = l2rel(d_ref, o64.params[k] - before[k].double())
                err = l2rel(P[k] - before[k], d_ref)
                assert err <= max(2 * RTOL, 10 * cond), f"{case} param delta {k}: {err:.2e} (cond {cond:.1e})"
    assert strict >= 0.85 * total, (strict, total)
    assert abs(eng.scalars()["log_temperature"] - float(orc.log_temperature)) < 1e-6
    eng.close()


def _golden_masks(z, meta, s, orc):
    """dropout multipliers of step s from the fixture's keep bits (make_golden.py)"""
    out = {}
    for k, (shape, pr) in orc.mask_shapes(meta["B"]).items():
        bits = np.unpackbits(z[f"drop{s}/{k}"])[:int(np.prod(shape))].astype(np.float32)
        out[k] = torch.from_numpy(bits / np.float32(1.0 - pr)).reshape(shape)
    return out


def test_cdt_dropout_golden(lib_built):
    """cdt_drop_small.npz: the UNMODIFIED reference stepped on prescribed dropout multipliers (0.1 at the embedding,
    attention-weight and both residual sites, as in cdt_configs.py); the engine replays the same multipliers."""
    z, meta = load_golden("cdt_drop_small")
    B, steps = meta["B"], meta["steps"]
    orc = ocdt.CDTOracle(_cfg(meta))
    eng = _engine(meta, B)
    assert list(eng.noise_layout) == list(orc.mask_shapes(B)), (list(eng.noise_layout), list(orc.mask_shapes(B)))
    init = {k: torch.from_numpy(z["init/" + k]) for k in meta["keys"]}
    eng.load_params(init)
    for s in range(steps):
        eng.step_seq({k: z[f"batch{s}/{k}"] for k in CDT_KEYS}, _golden_masks(z, meta, s, orc))
        got = eng.stats()
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            assert abs(got[k] - w) <= 2e-5 * max(abs(w), 1e-3) + 1e-7, f"step {s} {k}: {got[k]} vs reference {w}"
    got = eng.read_params()
    for k in meta["keys"]:
        if "in_proj_bias" in k:
            continue
        ref = torch.from_numpy(z["final/" + k])
        err = float((got[k] - ref).norm())
        bound = 1e-3 * float((ref - init[k]).norm()) + 4e-7 * float(ref.norm()) + 1e-9
        assert err <= bound, f"{k}: err {err:.3e} > {bound:.3e}"
    eng.close()


def test_cdt_dropout_device_masks_vs_live_oracle(lib_built):
    """Multipliers drawn on the device (Philox): right values and rates, and the oracle stepped on the dumped
    multipliers agrees with the engine (stats and gradients)."""
    z, meta = load_golden("cdt_drop_small")
    B = meta["B"]
    cfg = _cfg(meta)
    torch.manual_seed(0)
    orc = ocdt.CDTOracle(cfg)
    eng = _engine(meta, B)
    eng.load_params(orc.params)
    rng = np.random.default_rng(5)
    for s in range(2):
        b = make_seq_batch(rng, B, cfg.seq_len, cfg.state_dim, cfg.action_dim)
        eng.step_seq(b)                      # no noise given: every slot drawn on the device
        nz = eng.last_noise()
        for k, (shape, pr) in orc.mask_shapes(B).items():
            v = nz[k]
            keep = np.float32(1.0 / (1.0 - pr))
            assert np.all((v == 0) | (np.abs(v - keep) < 1e-6)), k
            assert abs(float((v == 0).mean()) - pr) < 0.02, (k, float((v == 0).mean()))
        o64 = _to_double(orc)
        masks = {k: torch.from_numpy(v) for k, v in nz.items()}
        s32 = orc.step(*_args(b), noise=masks)
        with algos.precision(torch.float64):
            s64 = o64.step(*_args(b, torch.float64), noise={k: v.double() for k, v in masks.items()})
        got = eng.stats()
        for k, w in s32.items():
            scale = max(abs(s64[k]), 1e-3)
            tol = max(RTOL, 10 * abs(w - s64[k]) / scale)
            assert abs(got[k] - w) <= tol * scale + 1e-7, f"step {s} stat {k}: {got[k]} vs {w}"
        G = eng.read_section("grad")
        bad = []
        for k, g in orc.last_grads.items():
            g64 = o64.last_grads[k]
            if float(g64.abs().max()) == 0.0:
                continue
            tol = max(2 * RTOL, 10 * maxrel(g, g64))
            err = maxrel(G[k], g)
            if err > tol:
                bad.append((k, err))
                assert err <= 2e-1, f"step {s} grad {k}: {err:.2e}"
        assert len(bad) <= max(1, int(0.2 * len(orc.last_grads))), bad[:6]
    eng.close()


def _seq_setup(B=32):
    from oracle import synth
    from osrl_b200 import Engine
    from osrl_b200.common.dataset import SequenceDataset
    d = synth.make_dataset(6, 3, 41, 25, seed=9)
    d["timeouts"][-7:] = False
    ds = SequenceDataset(d, seq_len=10, reward_scale=0.1, cost_scale=1.0, cost_sample=True,
                         cost_transform=lambda x: 70 - x)
    eng = Engine("cdt", batch_size=B, device=0, seed=4321, state_dim=6, action_dim=3, max_action=1.0, seq_len=10,
                 episode_len=100, embedding_dim=32, num_layers=1, num_heads=4, use_rew=1, use_cost=1, cost_transform=1,
                 stochastic=1, target_entropy=-3.0)
    ds.to_engine(eng)
    return d, ds, eng


def test_sequence_gather_bit_exact_and_sampler(lib_built):
    """K0s: windows for explicit (trajectory, start) pairs == SequenceDataset.__prepare_sample bit for bit
    (slice, float32 scaling, end zero-pad, mask, unclipped time_steps; dataset.py:749-775); the on-device draw
    == the numpy restatement; the alias table reproduces sample_prob."""
    from oracle.sampler import philox_sequences
    d, ds, eng = _seq_setup()
    tr = ocdt.split_trajectories(d)
    rng = np.random.default_rng(1)
    tj = rng.integers(0, len(tr), 300)
    st = np.array([rng.integers(0, tr[t]["rewards"].shape[0]) for t in tj])
    tj[:3], st[:3] = [0, len(tr) - 1, 5], [0, 40, 39]
    out = eng.seq_gather(tj, st)
    for i, (t, s) in enumerate(zip(tj, st)):
        want = ocdt.sequence_sample(tr, int(t), int(s), 10, 0.1, 1.0)
        for k, w in zip(("states", "actions", "returns", "costs_return", "time_steps", "mask"), want[:6]):
            assert np.array_equal(out[k][i].cpu().numpy().astype(np.asarray(w).dtype), np.asarray(w)), (k, t, s)
        assert np.array_equal(out["costs"][i].cpu().numpy(), want[7])
    prob, alias = eng.alias_table()
    n = len(tr)
    p = prob.astype(np.float64) / n
    for s_ in range(n):
        p[alias[s_]] += (1.0 - prob[s_]) / n
    assert np.allclose(p, ds.sample_prob, atol=2e-7)
    for step in range(3):
        eng.steps(1)
        t_eng, s_eng = eng.last_sequences()
        t_ref, s_ref = philox_sequences(4321, step, 0, 32, prob, alias, ds.offsets)
        assert np.array_equal(t_eng, t_ref) and np.array_equal(s_eng, s_ref)
        assert all(np.isfinite(v) for v in eng.stats().values())
    eng.close()


def test_cdt_trainer_api_matches_reference_fixture(lib_built):
    """The public mirror (osrl_b200.algorithms.CDT / CDTTrainer) driven like examples/train/train_cdt.py:181-187
    reproduces the reference-produced stats of tests/golden/cdt_small.npz."""
    from osrl_b200.algorithms import CDT, CDTTrainer
    z, meta = load_golden("cdt_small")
    c = meta["cfg"]

    class Log:
        rows = []

        def store(self, tab=None, **kw):
            self.rows.append(kw)

    torch.manual_seed(0)
    model = CDT(c["state_dim"], c["action_dim"], 1.0, seq_len=c["seq_len"], episode_len=c["episode_len"],
                embedding_dim=c["embedding_dim"], num_layers=c["num_layers"], num_heads=c["num_heads"], time_emb=True,
                use_rew=True, use_cost=True, cost_transform=True, stochastic=True, init_temperature=0.1,
                target_entropy=-c["action_dim"], device="cuda:0")
    for k in meta["keys"]:
        assert torch.equal(model.state_dict()[k], torch.from_numpy(z["init/" + k])), k   # same seed -> same init
    log = Log()
    tr = CDTTrainer(model, None, log, learning_rate=c["learning_rate"], weight_decay=c["weight_decay"],
                    betas=tuple(c["betas"]), clip_grad=c["clip_grad"], lr_warmup_steps=c["lr_warmup_steps"],
                    loss_cost_weight=c["loss_cost_weight"], loss_state_weight=c["loss_state_weight"], device="cuda:0")
    for s in range(meta["steps"]):
        tr.train_one_step(*[torch.from_numpy(z[f"batch{s}/{k}"]) for k in CDT_KEYS])
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            assert abs(log.rows[-1][k] - w) <= 2e-5 * max(abs(w), 1e-3) + 1e-7, (s, k)
    sd = model.state_dict()     # views of the engine arena
    k = "blocks.0.mlp.0.weight"
    ref = torch.from_numpy(z["final/" + k])
    assert sd[k].is_cuda and float((sd[k].cpu() - ref).norm()) <= 1e-3 * float((ref - torch.from_numpy(z["init/" + k])).norm())


def test_cdt_b2048_fixture_and_split_k(lib_built):
    """BASELINE.json configs[3] at its own size (B=2048, T=10, 3 layers, E=128, dropout 0.1 at all three sites):
    81,920 tokens put the weight gradients on the split-K path.  Stats vs tests/golden/cdt_b2048.npz (written by the
    UNMODIFIED reference on multipliers from torch.Generator(4242)); gradients vs the live oracle on the same
    multipliers; and run-to-run reproducibility of the gradients (the split-K partials are combined in a fixed order)."""
    from tests.helpers import record_margin
    z, meta = load_golden("cdt_b2048")
    B, steps = meta["B"], meta["steps"]
    cfg = _cfg(meta)
    torch.manual_seed(0)
    orc = ocdt.CDTOracle(cfg)
    ck = np.array([[float(v.double().sum()), float(v.double().abs().sum()), float((v.double() ** 2).sum())]
                   for v in orc.params.values()])
    assert np.allclose(ck, z["init_checksum"], rtol=1e-12, atol=0), "oracle init differs from the fixture's reference init"
    eng = _engine(meta, B)
    eng2 = _engine(meta, B)
    eng.load_params(orc.params)
    eng2.load_params(orc.params)
    rng = np.random.default_rng(77)
    mgen = torch.Generator().manual_seed(4242)
    for s in range(steps):
        b = make_seq_batch(rng, B, cfg.seq_len, cfg.state_dim, cfg.action_dim)
        masks = orc.draw_masks(B, generator=mgen)
        eng.step_seq(b, masks)
        got = eng.stats()
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            tol = 2e-5 * max(abs(w), 1e-3) + 1e-7
            record_margin("cdt_b2048", "stat " + k, abs(got[k] - w), tol)
            assert abs(got[k] - w) <= tol, f"step {s} {k}: {got[k]} vs reference {w}"
        orc.step(*_args(b), noise=masks)
        G = eng.read_section("grad")
        bad = []
        for k, g in orc.last_grads.items():
            if float(g.abs().max()) == 0.0 or "in_proj_bias" in k:
                continue
            err = maxrel(G[k], g)
            record_margin("cdt_b2048", "grad (max-rel)", err, 2e-5)
            if err > 2e-5:      # 81,920-term fp32 reductions in a different order than torch's: measured worst 4.2e-6
                bad.append((k, err))
                assert err <= 1e-4, f"step {s} grad {k}: {err:.2e}"
        assert len(bad) <= 1, bad[:6]
        eng2.step_seq(b, masks)
        G2 = eng2.read_section("grad")
        worst = max(maxrel(G2[k], G[k]) for k in G if float(G[k].abs().max()) > 0)
        # split-K partials and the timestep-embedding scatter are accumulated with float atomics: the order, hence the
        # last bits, vary from run to run; the measured spread is recorded in profiles/r02_parity_margins.json
        record_margin("cdt_b2048", "run-to-run gradient difference (max-rel)", worst, 1e-5)
        assert worst <= 1e-5, f"step {s}: gradients differ run to run by {worst:.2e}"
    P = eng.read_params()
    ckf = z["final_checksum"]
    for i, k in enumerate(meta["keys"]):
        if "in_proj_bias" in k:
            continue
        v = P[k].double()
        got_ck = np.array([float(v.sum()), float(v.abs().sum()), float((v * v).sum())])
        # sum |p| and sum p^2 are well conditioned: 1e-5 relative of the reference's
        assert abs(got_ck[1] - ckf[i][1]) <= 1e-5 * ckf[i][1] + 1e-9, (k, got_ck, ckf[i])
        assert abs(got_ck[2] - ckf[i][2]) <= 2e-5 * ckf[i][2] + 1e-12, (k, got_ck, ckf[i])
    eng.close()
    eng2.close()
