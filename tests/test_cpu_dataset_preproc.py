"""One-time dataset preprocessing restated on flat arrays (osrl_b200/common/dataset.py) pinned against the UNMODIFIED
reference (osrl/common/dataset.py) on seeded synthetic data: process_bc_dataset (all modes incl. the Pareto frontier),
SequenceDataset with the reference's default Pareto-frontier augmentation, random augmentation, pf_only and the
sampling-probability variants.  The reference needs `oapackage` for the frontier; tests/stubs/oapackage.py restates
its two classes (brute-force non-dominated set).  Skipped where /root/reference is absent."""
import os
import random
import sys

import numpy as np
import pytest

from oracle import ref_shim, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not available")


@pytest.fixture(scope="module")
def ref_dataset():
    sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))
    import oapackage  # noqa: F401  (the stub; must be importable before the reference module is)
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]
    osrl = ref_shim.import_reference()
    import osrl.common.dataset as rd
    if not hasattr(rd, "oapackage"):
        rd.oapackage = oapackage
    yield rd
    sys.path.remove(os.path.join(ROOT, "tests", "stubs"))


def _data(seed=3, eps=60, T=37):
    d = synth.make_dataset(6, 2, T, eps, seed=seed)
    rng = np.random.default_rng(seed)
    d["costs"] = (rng.random(d["costs"].shape) < rng.uniform(0.02, 0.5, eps).repeat(T)).astype(np.float32)
    d["rewards"] = (d["rewards"] + rng.uniform(0, 1.5, eps).repeat(T)).astype(np.float32)
    d["timeouts"][-5:] = False
    return d


def test_pareto_front_matches_brute_force():
    from osrl_b200.common.dataset import pareto_front_2d
    rng = np.random.default_rng(1)
    for n in (1, 7, 300):
        c = rng.integers(0, 25, n).astype(float)
        r = (c * 0.7 + rng.integers(0, 12, n)).astype(float)
        brute = [i for i in range(n) if not any((c[j] <= c[i] and r[j] >= r[i]) and (c[j] < c[i] or r[j] > r[i]) for j in range(n))]
        assert brute == pareto_front_2d(c, r)


@pytest.mark.parametrize("mode", ["all", "multi-task", "safe", "risky", "boundary", "frontier"])
@pytest.mark.parametrize("gamma", [1.0, 0.99])
def test_process_bc_dataset(ref_dataset, mode, gamma):
    from osrl_b200.common.dataset import process_bc_dataset
    d = _data()
    a, b = {k: v.copy() for k, v in d.items()}, {k: v.copy() for k, v in d.items()}
    ref_dataset.process_bc_dataset(a, 6, gamma, mode)
    process_bc_dataset(b, 6, gamma, mode)
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (mode, gamma, k)


CASES = {
    "default_augmentation": dict(augment_percent=0.2, deg=2, max_rew_decrease=100.0, max_reward=60.0, min_reward=1.0,
                                 cost_sample=True, cost_transform=lambda x: 70 - x),
    "random_aug": dict(random_aug=0.3, aug_rmin=20, aug_rmax=40, aug_cmin=2, aug_cmax=12, cgap=1, cost_sample=True,
                       cost_transform=lambda x: 70 - x),
    "pf_only": dict(pf_only=True, rmin=10, cost_bins=10, npb=2),
    "pf_sample": dict(augment_percent=0.2, deg=2, max_rew_decrease=100.0, max_reward=60.0, min_reward=1.0, pf_sample=True),
    "start_sampling": dict(start_sampling=True, prob=0.3, cost_sample=True, cost_transform=lambda x: 70 - x),
    "cost_reverse": dict(cost_reverse=True, cost_sample=True, cost_transform=lambda x: 70 - x),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_sequence_dataset_modes(ref_dataset, case):
    """Same seeds -> same trajectories (sources, relabelled return / cost to go), same sampling probabilities, and the
    same windows from the iterator."""
    from osrl_b200.common.dataset import SequenceDataset
    kw = CASES[case]
    d = _data(seed=5, eps=900, T=16)   # enough trajectories for the 10 x 50 outlier grid (bins of <= 2 are dropped)
    np.random.seed(7); random.seed(7)
    ref = ref_dataset.SequenceDataset({k: v.copy() for k, v in d.items()}, seq_len=10, reward_scale=0.1, cost_scale=1.0, **kw)
    np.random.seed(7); random.seed(7)
    mine = SequenceDataset({k: v.copy() for k, v in d.items()}, seq_len=10, reward_scale=0.1, cost_scale=1.0, **kw)
    assert len(ref.dataset) == len(mine)
    for i, t in enumerate(ref.dataset):
        lo, hi, rtg, ctg = mine._traj[i]
        assert np.array_equal(t["observations"], mine._obs[lo:hi]) and np.array_equal(t["actions"], mine._act[lo:hi]), (case, i)
        assert np.array_equal(t["costs"], mine._cost[lo:hi]), (case, i)
        # (relabelled targets: the frontier fit sees the same points in a different order -- 1e-6 relative)
        assert np.allclose(t["returns"], rtg, rtol=2e-6, atol=1e-5) and np.allclose(t["cost_returns"], ctg, rtol=2e-6, atol=1e-5), (case, i)
    if ref.sample_prob is None:
        assert mine.sample_prob is None
    else:
        assert np.allclose(ref.sample_prob, mine.sample_prob, rtol=1e-5, atol=1e-9)
    if kw.get("start_sampling"):
        for a, b in zip(ref.start_idx_sample_prob, mine.start_idx_sample_prob):
            assert np.allclose(a, b, rtol=1e-6)
    np.random.seed(11); random.seed(11)
    ra = [x for _, x in zip(range(5), iter(ref))]
    np.random.seed(11); random.seed(11)
    ma = [x for _, x in zip(range(5), iter(mine))]
    for x, y in zip(ra, ma):
        for u, v in zip(x, y):
            assert np.asarray(u).dtype == np.asarray(v).dtype and np.allclose(u, v, rtol=2e-6, atol=1e-5), case


@pytest.mark.parametrize("cost_reverse", [False, True])
def test_oracle_split_trajectories_is_the_reference(ref_dataset, cost_reverse):
    """oracle/cdt.py: split_trajectories -- the checker of the device preprocessing (tests/test_gpu_preproc.py) --
    against the unmodified process_sequence_dataset (dataset.py:137-183), bit for bit, with fractional costs, a
    trailing unfinished episode and both end flags."""
    from oracle import cdt as ocdt
    d = _data(seed=9, eps=40, T=23)
    rng = np.random.default_rng(2)
    d["costs"] = np.where(rng.random(d["costs"].shape) < 0.5, d["costs"], rng.random(d["costs"].shape)).astype(np.float32)
    d["terminals"][40] = True
    want, _ = ref_dataset.process_sequence_dataset({k: v.copy() for k, v in d.items()}, cost_reverse)
    got = ocdt.split_trajectories(d, cost_reverse)
    assert len(want) == len(got)
    for w, g in zip(want, got):
        for k in ("observations", "actions", "rewards", "costs", "returns", "cost_returns"):
            assert w[k].dtype == g[k].dtype and np.array_equal(w[k], g[k]), k


@pytest.mark.parametrize("state_init", [False, True])
def test_transition_dataset_stream_is_the_reference(ref_dataset, state_init):
    """TransitionDataset (dataset.py:790-847): done / is_init columns, the float32 scaling of rewards and costs, the
    tuple layout and the index stream of __iter__ under the same numpy seed; get_dataset_states for COptiDICE."""
    from osrl_b200.common.dataset import TransitionDataset
    d = _data(seed=4, eps=12, T=9)
    d["terminals"][20] = True
    a = ref_dataset.TransitionDataset({k: v.copy() for k, v in d.items()}, 0.1, 3.0, state_init)
    b = TransitionDataset({k: v.copy() for k, v in d.items()}, 0.1, 3.0, state_init)
    for k in ("done",) + (("is_init",) if state_init else ()):
        assert a.dataset[k].dtype == b.dataset[k].dtype and np.array_equal(a.dataset[k], b.dataset[k]), k
    np.random.seed(2)
    ia = iter(a)
    want = [next(ia) for _ in range(25)]
    np.random.seed(2)
    ib = iter(b)
    got = [next(ib) for _ in range(25)]
    for w, g in zip(want, got):
        assert len(w) == len(g) == (7 if state_init else 6)
        for x, y in zip(w, g):
            assert np.asarray(x).dtype == np.asarray(y).dtype and np.array_equal(x, y)
    if state_init:
        for x, y in zip(a.get_dataset_states(), b.get_dataset_states()):
            assert np.array_equal(np.asarray(x), np.asarray(y))
