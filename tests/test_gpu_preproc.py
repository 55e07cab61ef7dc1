"""process_sequence_dataset on the device (osrl_seq_preprocess) against the oracle restatement of
dataset.py:137-183 / :19-27 -- episode split, reward-to-go and cost-to-go bit for bit, the windows osrl_seq_gather
builds from the device-made buffer, and agreement with the host-preprocessed upload path."""
import numpy as np
import pytest
import torch

from oracle import cdt as ocdt
from osrl_b200 import Engine
from osrl_b200.common.dataset import SequenceDataset

pytestmark = pytest.mark.gpu
O, A, T = 5, 3, 10


def _flat_dataset(seed, n=6000, tail=37):
    """random episode lengths 1..300 ended by terminals or timeouts, fractional and 0/1 costs, an unfinished tail"""
    rng = np.random.default_rng(seed)
    term, tout = np.zeros(n, bool), np.zeros(n, bool)
    i = 0
    while True:
        i += int(rng.integers(1, 300))
        if i >= n - tail:
            break
        (term if rng.random() < 0.4 else tout)[i - 1] = True
    if seed % 2:                                   # both flags on the same transition
        j = int(np.flatnonzero(term | tout)[3]); term[j] = tout[j] = True
    costs = np.where(rng.random(n) < 0.7, (rng.random(n) < 0.2).astype(np.float32), rng.random(n).astype(np.float32))
    return {"observations": rng.standard_normal((n, O)).astype(np.float32),
            "actions": rng.uniform(-1, 1, (n, A)).astype(np.float32),
            "rewards": (rng.standard_normal(n) * 3).astype(np.float32), "costs": costs.astype(np.float32),
            "terminals": term, "timeouts": tout}


def _engine(B=8):
    return Engine("cdt", batch_size=B, device=0, seed=3, state_dim=O, action_dim=A, max_action=1.0, seq_len=T,
                  episode_len=300, embedding_dim=32, num_layers=1, num_heads=4, use_rew=1, use_cost=1, cost_transform=1,
                  stochastic=1, target_entropy=-float(A), learning_rate=1e-3, lr_warmup_steps=4)


@pytest.mark.parametrize("cost_reverse", [False, True])
def test_device_preprocess_bit_exact(lib_built, cost_reverse):
    data = _flat_dataset(7 + int(cost_reverse))
    rs, cs = 0.1, 2.0
    trajs = ocdt.split_trajectories(data, cost_reverse)
    eng = _engine()
    info = eng.preprocess_seq_dataset(data, rs, cs, cost_reverse)
    lens = np.array([t["rewards"].shape[0] for t in trajs])
    assert eng.n_traj == len(trajs) and info["n_used"] == int(lens.sum())
    assert np.array_equal(info["traj_offsets"], np.concatenate([[0], np.cumsum(lens)]))
    assert np.array_equal(info["returns"], np.array([t["returns"][0] for t in trajs], np.float32))
    assert np.array_equal(info["cost_returns"], np.array([t["cost_returns"][0] for t in trajs], np.float32))
    # every window shape the sampler can ask for: start 0, the last transition, a random start
    rng = np.random.default_rng(0)
    ti = np.repeat(np.arange(len(trajs)), 3)
    si = np.stack([np.zeros_like(lens), lens - 1, rng.integers(0, lens)], 1).reshape(-1)
    got = {k: v.cpu().numpy() for k, v in eng.seq_gather(ti, si).items()}
    keys = ("states", "actions", "returns", "costs_return", "time_steps", "mask", None, "costs")
    for row, (t, s0) in enumerate(zip(ti, si)):
        want = ocdt.sequence_sample(trajs, int(t), int(s0), T, rs, cs)
        for k, w in zip(keys, want):
            if k is not None:
                assert np.array_equal(got[k][row], np.asarray(w).astype(got[k].dtype)), (k, t, s0)
    # the host-preprocessed upload path leaves the same buffer
    host = _engine()
    ds = SequenceDataset(data, seq_len=T, reward_scale=rs, cost_scale=cs, cost_reverse=cost_reverse)
    ds.to_engine(host)
    ref = {k: v.cpu().numpy() for k, v in host.seq_gather(ti, si).items()}
    for k in got:
        assert np.array_equal(got[k], ref[k]), k
    eng.close(); host.close()


def test_device_preprocess_sampling_and_training(lib_built):
    data = _flat_dataset(3)
    tf = lambda c: 60.0 - c
    dev, host = _engine(), _engine()
    info = SequenceDataset.device_resident(dev, data, 0.1, 1.0, cost_sample=True, cost_transform=tf)
    ds = SequenceDataset(data, seq_len=T, reward_scale=0.1, cost_scale=1.0, cost_sample=True, cost_transform=tf)
    ds.to_engine(host)
    assert np.allclose(info["sample_prob"], ds.sample_prob, rtol=0, atol=0)
    pa, aa = dev.alias_table(), host.alias_table()
    assert np.array_equal(pa[0], aa[0]) and np.array_equal(pa[1], aa[1])
    # same seed, same resident data, same distribution -> the same windows and the same training steps
    dev.steps(3); host.steps(3)
    assert dev.stats() == host.stats() and all(np.isfinite(v) for v in dev.stats().values())
    a, b = dev.read_params(), host.read_params()
    assert all(torch.equal(a[k], b[k]) for k in a)
    # a second call replaces the buffer (different data) without leaking the first
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(3):
        dev.preprocess_seq_dataset(_flat_dataset(11), 0.1, 1.0)
    assert torch.cuda.mem_get_info()[0] >= free0 - (8 << 20)
    dev.set_seq_sample_prob(None)
    dev.steps(1)
    dev.close(); host.close()
