"""GPU parity: the CUDA engine (through the C ABI) against the oracle and the golden fixtures.

Bar (BASELINE.json north_star): losses and parameter deltas within 1e-5 relative of the
reference path on identical seeds / batches / replayed noise; index gather bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import synth
from tests.helpers import RTOL, batch_tuple, load_golden, make_oracle, rel_delta_err

pytestmark = pytest.mark.gpu


def _engine(meta, B):
    from osrl_b200 import Engine
    return Engine(meta["algo"], batch_size=B, device=0, seed=7, **meta["cfg"])


def _check_stats(got, want_row, names, tag):
    for k, w in zip(names, want_row):
        g = got[k]
        assert abs(g - w) <= RTOL * max(abs(w), 1e-3) + 1e-7, f"{tag}: {k}: engine {g} vs reference {w}"


@pytest.mark.parametrize("case", ["bc_small", "bcql_small", "cpq_small", "bearl_small"])
def test_small_golden(lib_built, case):
    """Engine vs fixtures produced by the UNMODIFIED reference (tests/golden, oracle/make_golden.py)."""
    z, meta = load_golden(case)
    algo, B, steps = meta["algo"], meta["B"], meta["steps"]
    eng = _engine(meta, B)
    init = {k: torch.from_numpy(z["init/" + k]) for k in meta["keys"]}
    eng.load_params(init)
    for s in range(steps):
        batch = {k: z[f"batch{s}/{k}"] for k in ("observations", "next_observations", "actions", "rewards", "costs", "done")}
        noise = {k: z[f"noise{s}/{k}"] for k in eng.noise_layout}
        eng.step(batch, noise)
        _check_stats(eng.stats(), z["stats"][s], meta["stat_keys"], f"{case} step {s}")
    got = eng.read_params()
    worst = max(rel_delta_err(got[k], torch.from_numpy(z["final/" + k]), init[k]) for k in meta["keys"]
                if (torch.from_numpy(z["final/" + k]) - init[k]).abs().max() > 0)
    assert worst <= 10 * RTOL, f"{case}: parameter delta error {worst:.2e}"
    eng.close()


@pytest.mark.parametrize("case", ["bc_full", "bcql_full", "cpq_full", "bearl_full"])
def test_full_size_against_live_oracle(lib_built, case):
    """BASELINE.json layer sizes: live oracle on the same seeds, plus the pinned reference stats."""
    z, meta = load_golden(case)
    algo, B, steps = meta["algo"], meta["B"], meta["steps"]
    orc = make_oracle(algo, meta["cfg"], meta["init_seed"])
    init = {k: v.clone() for k, v in orc.params.items()}
    eng = _engine(meta, B)
    eng.load_params(init)
    rng = np.random.default_rng(meta["data_seed"])
    cfg = meta["cfg"]
    torch.manual_seed(meta["noise_seed"])
    for s in range(steps):
        b = synth.make_batch(rng, B, cfg["state_dim"], cfg["action_dim"])
        ostats = orc.step(*batch_tuple(algo, b))
        eng.step(b, {k: v for k, v in orc.last_noise.items() if k in eng.noise_layout})
        got = eng.stats()
        for k, w in ostats.items():
            assert abs(got[k] - w) <= RTOL * max(abs(w), 1e-3) + 1e-7, f"{case} step {s} {k}: {got[k]} vs oracle {w}"
        # pinned reference numbers (same torch build => same noise stream)
        for k, w in zip(meta["stat_keys"], z["stats"][s]):
            if abs(ostats[k] - w) <= 1e-6 * max(abs(w), 1e-3):
                assert abs(got[k] - w) <= RTOL * max(abs(w), 1e-3) + 1e-7
    got = eng.read_params()
    worst = 0.0
    for k in meta["keys"]:
        if (orc.params[k] - init[k]).abs().max() > 0:
            worst = max(worst, rel_delta_err(got[k], orc.params[k].detach(), init[k]))
    assert worst <= 10 * RTOL, f"{case}: parameter delta error {worst:.2e}"
    eng.close()


def test_gather_bit_exact_and_sampler(lib_built):
    """K0: row gather == dataset[k][idx] bit for bit, including the float32 reward/cost scaling
    (dataset.py:832-842); on-device Philox index draw == the oracle's numpy Philox."""
    from oracle.sampler import philox_indices
    from osrl_b200 import Engine
    data = synth.make_dataset(8, 2, 50, 40, seed=3)
    eng = Engine("bcql", batch_size=64, device=0, seed=1234, state_dim=8, action_dim=2, a_hidden_sizes=[16, 16],
                 c_hidden_sizes=[16, 16], vae_hidden_sizes=16)
    eng.upload_dataset(data, reward_scale=0.1, cost_scale=2.0)
    n = data["observations"].shape[0]
    for idx in (np.array([0, n - 1, 5, 5, 17]), np.random.default_rng(0).integers(0, n, 1000), np.array([3])):
        out = eng.gather(idx)
        assert np.array_equal(out["observations"].cpu().numpy(), data["observations"][idx])
        assert np.array_equal(out["next_observations"].cpu().numpy(), data["next_observations"][idx])
        assert np.array_equal(out["actions"].cpu().numpy(), data["actions"][idx])
        assert np.array_equal(out["rewards"].cpu().numpy(), data["rewards"][idx] * 0.1)
        assert np.array_equal(out["costs"].cpu().numpy(), data["costs"][idx] * 2.0)
        done = np.logical_or(data["terminals"], data["timeouts"]).astype(np.float32)
        assert np.array_equal(out["done"].cpu().numpy(), done[idx])
    for step in range(3):
        eng.steps(1)
        assert np.array_equal(eng.last_indices(), philox_indices(1234, step, 0, 64, n))
    eng.close()


def test_sampled_steps_match_oracle_on_dumped_batch(lib_built):
    """osrl_steps(): on-device sampling + on-device noise; re-run the oracle on the dumped indices
    and noise and compare (full BCQ-Lag config, 2 steps)."""
    z, meta = load_golden("bcql_full")
    cfg, B = meta["cfg"], meta["B"]
    data = synth.make_dataset(cfg["state_dim"], cfg["action_dim"], 300, 60, seed=0)
    orc = make_oracle("bcql", cfg, 0)
    init = {k: v.clone() for k, v in orc.params.items()}
    eng = _engine(meta, B)
    eng.load_params(init)
    eng.upload_dataset(data, reward_scale=0.1, cost_scale=1.0)
    done = np.logical_or(data["terminals"], data["timeouts"]).astype(np.float32)
    for s in range(2):
        eng.steps(1)
        idx = eng.last_indices()
        nz = {k: torch.from_numpy(v) for k, v in eng.last_noise().items()}
        for v in nz.values():
            assert torch.isfinite(v).all() and 0.9 < v.std() < 1.1 and abs(v.mean()) < 0.1
        b = {"observations": data["observations"][idx], "next_observations": data["next_observations"][idx],
             "actions": data["actions"][idx], "rewards": data["rewards"][idx] * np.float32(0.1),
             "costs": data["costs"][idx] * np.float32(1.0), "done": done[idx]}
        ostats = orc.step(*batch_tuple("bcql", b), noise=nz)
        got = eng.stats()
        for k, w in ostats.items():
            assert abs(got[k] - w) <= RTOL * max(abs(w), 1e-3) + 1e-7, f"step {s} {k}: {got[k]} vs {w}"
    eng.close()
