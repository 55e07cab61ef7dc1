"""Executable counterpart of tests/test_cpu_dropin_surface.py: the training loop of the reference's example scripts
(examples/train/train_{bc,bcql,cpq,bearl,coptidice,cdt}.py: model, trainer, dataset, DataLoader, `.to(device)`,
`train_one_step` x 20, `evaluate`, `{"model_state": state_dict}` checkpoint, reload) written against `osrl.*` -- which
resolves to the alias package compat/osrl -- with the Appendix-D stubs standing in for gym / dsrl / fsrl.  The scripts
themselves are not on the GPU box (no /root/reference there); every call below uses the keywords they use."""
import io
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


@pytest.fixture(scope="module")
def osrl_alias(lib_built):
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]
    paths = [os.path.join(ROOT, "compat"), os.path.join(ROOT, "tests", "stubs")]
    for p in paths:
        sys.path.insert(0, p)
    import osrl.algorithms  # noqa: F401
    import osrl.common  # noqa: F401
    import osrl
    assert "compat" in osrl.__file__
    yield osrl
    for p in paths:
        sys.path.remove(p)
    for m in [m for m in sys.modules if m == "osrl" or m.startswith("osrl.")]:
        del sys.modules[m]


def _env(task):
    import gymnasium as gym
    from dsrl.offline_env import OfflineEnvWrapper, wrap_env
    env = gym.make(task)
    data = env.get_dataset()
    env.set_target_cost(10)
    return OfflineEnvWrapper(wrap_env(env=env, reward_scale=0.1)), data


def _checkpoint_roundtrip(model, fresh):
    buf = io.BytesIO()
    torch.save({"model_state": model.state_dict()}, buf)      # train_bcql.py:108-109
    buf.seek(0)
    fresh.load_state_dict(torch.load(buf)["model_state"])     # eval_bcql.py:45-49
    fresh.to(DEV)
    for (k, a), (_, b) in zip(model.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k


@pytest.mark.parametrize("algo", ["bc", "bcql", "cpq", "bearl", "coptidice"])
def test_transition_algorithms_train_loop(osrl_alias, algo):
    from fsrl.utils import WandbLogger
    from osrl.algorithms import (BC, BCQL, BEARL, CPQ, BCQLTrainer, BCTrainer, BEARLTrainer, COptiDICE, COptiDICETrainer,
                                 CPQTrainer)
    from osrl.common import TransitionDataset
    from osrl.common.dataset import process_bc_dataset
    from osrl.common.exp_util import auto_name, seed_all
    from torch.utils.data import DataLoader
    env, data = _env("OfflineCarCircle-v0")
    logger = WandbLogger({}, "p", "g", auto_name({"a": 1}, {"a": 2}, "BCQL", ""), None)
    seed_all(0)
    o, a, lim = env.observation_space.shape[0], env.action_space.shape[0], env.action_space.high[0]
    common = dict(a_hidden_sizes=[32, 32], c_hidden_sizes=[32, 32], vae_hidden_sizes=48, sample_action_num=10, gamma=0.99,
                  tau=0.005, num_q=2, num_qc=2, cost_limit=10, episode_len=30, device=DEV)
    if algo == "bc":
        process_bc_dataset(data, 10, 1.0, "all")
        make = lambda: BC(state_dim=o, action_dim=a, max_action=lim, a_hidden_sizes=[32, 32], episode_len=30, device=DEV)
        model = make()
        trainer = BCTrainer(model, env, logger=logger, actor_lr=1e-3, bc_mode="all", cost_limit=10, device=DEV)
        dataset = TransitionDataset(data)
    elif algo == "bcql":
        make = lambda: BCQL(state_dim=o, action_dim=a, max_action=lim, PID=[0.1, 0.003, 0.001], lmbda=0.75, beta=0.5,
                            phi=0.05, **common)
        model = make()
        trainer = BCQLTrainer(model, env, logger=logger, actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3, reward_scale=0.1,
                              cost_scale=1, device=DEV)
        dataset = TransitionDataset(data, reward_scale=0.1, cost_scale=1)
    elif algo == "cpq":
        make = lambda: CPQ(state_dim=o, action_dim=a, max_action=lim, beta=0.5, qc_scalar=1.5, **common)
        model = make()
        trainer = CPQTrainer(model, env, logger=logger, actor_lr=1e-4, critic_lr=1e-3, alpha_lr=1e-4, vae_lr=1e-3,
                             reward_scale=0.1, cost_scale=1, device=DEV)
        dataset = TransitionDataset(data, reward_scale=0.1, cost_scale=1)
    elif algo == "bearl":
        make = lambda: BEARL(state_dim=o, action_dim=a, max_action=lim, beta=0.5, lmbda=0.75, mmd_sigma=50,
                             target_mmd_thresh=0.05, start_update_policy_step=0, PID=[0.1, 0.003, 0.001], **common)
        model = make()
        trainer = BEARLTrainer(model, env, logger=logger, actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3, reward_scale=0.1,
                               cost_scale=1, device=DEV)
        dataset = TransitionDataset(data, reward_scale=0.1, cost_scale=1)
    else:
        dataset = TransitionDataset(data, reward_scale=0.1, cost_scale=1, state_init=True)
        p0, osd, asd = dataset.get_dataset_states()
        make = lambda: COptiDICE(state_dim=o, action_dim=a, max_action=lim, f_type="softchi", init_state_propotion=p0,
                                 observations_std=osd, actions_std=asd, a_hidden_sizes=[32, 32], c_hidden_sizes=[32, 32],
                                 gamma=0.99, alpha=0.5, cost_ub_epsilon=0.01, num_nu=2, num_chi=2, cost_limit=10,
                                 episode_len=30, device=DEV)
        model = make()
        trainer = COptiDICETrainer(model, env, logger=logger, actor_lr=1e-3, critic_lr=1e-3, scalar_lr=1e-3,
                                   reward_scale=0.1, cost_scale=1, device=DEV)
    print(f"Total parameters: {sum(p.numel() for p in model.parameters())}")
    logger.setup_checkpoint_fn(lambda: {"model_state": model.state_dict()})
    it = iter(DataLoader(dataset, batch_size=64, pin_memory=True, num_workers=0))
    for step in range(20):
        batch = [b.to(DEV) for b in next(it)]
        if algo == "bc":
            trainer.train_one_step(batch[0], batch[2])            # train_bc.py:121-123
        elif algo == "coptidice":
            trainer.train_one_step(batch)                          # train_coptidice.py:146-148
        else:
            trainer.train_one_step(*batch)                         # train_bcql.py:143-148
        logger.write_without_reset(step)
    assert logger.rows and all(np.isfinite(v) for v in logger.rows[-1].values())
    ret, cost, length = trainer.evaluate(2)                         # train_bcql.py:152
    assert np.isfinite(ret) and np.isfinite(cost) and length == 30
    logger.save_checkpoint()
    assert logger.saved and "model_state" in logger.saved[-1][1]
    _checkpoint_roundtrip(model, make())


def test_cdt_train_loop_default_config(osrl_alias):
    """train_cdt.py:73-187 with the DEFAULT data pipeline of cdt_configs.py (augment_percent=0.2, cost_sample, dropout 0.1)."""
    from fsrl.utils import WandbLogger
    from osrl.algorithms import CDT, CDTTrainer
    from osrl.common import SequenceDataset
    from osrl.common.exp_util import seed_all
    from torch.utils.data import DataLoader
    import gymnasium as gym
    env = gym.make("OfflineCarCircle-v0")
    env.episodes, env.episode_len = 900, 16
    data = env.get_dataset()
    logger = WandbLogger({}, "p", "g", "n", None)
    seed_all(0)
    o, a = env.observation_space.shape[0], env.action_space.shape[0]
    make = lambda: CDT(state_dim=o, action_dim=a, max_action=env.action_space.high[0], embedding_dim=32, seq_len=10,
                       episode_len=30, num_layers=2, num_heads=4, attention_dropout=0.1, residual_dropout=0.1,
                       embedding_dropout=0.1, time_emb=True, use_rew=True, use_cost=True, cost_transform=True,
                       add_cost_feat=False, mul_cost_feat=False, cat_cost_feat=False, action_head_layers=1,
                       cost_prefix=False, stochastic=True, init_temperature=0.1, target_entropy=-a).to(DEV)
    model = make()
    trainer = CDTTrainer(model, env, logger=logger, learning_rate=1e-4, weight_decay=1e-4, betas=(0.9, 0.999),
                         clip_grad=0.25, lr_warmup_steps=500, reward_scale=0.1, cost_scale=1, loss_cost_weight=0.02,
                         loss_state_weight=0, cost_reverse=False, no_entropy=False, device=DEV)
    ct = lambda x: 70 - x
    dataset = SequenceDataset(data, seq_len=10, reward_scale=0.1, cost_scale=1, deg=2, pf_sample=False,
                              max_rew_decrease=100.0, beta=1.0, augment_percent=0.2, cost_reverse=False, max_reward=60.0,
                              min_reward=1.0, pf_only=False, rmin=300, cost_bins=60, npb=5, cost_sample=True,
                              cost_transform=ct, start_sampling=False, prob=0.2, random_aug=0, aug_rmin=400,
                              aug_rmax=500, aug_cmin=-2, aug_cmax=25, cgap=5, rstd=1, cstd=0.2)
    assert dataset.n_augmented > 0
    it = iter(DataLoader(dataset, batch_size=32, pin_memory=True, num_workers=0))
    for step in range(20):
        states, actions, returns, costs_return, time_steps, mask, episode_cost, costs = [b.to(DEV) for b in next(it)]
        trainer.train_one_step(states, actions, returns, costs_return, time_steps, mask, episode_cost, costs)
    assert all(np.isfinite(v) for v in logger.rows[-1].values()) and "act_loss" in logger.rows[-1]
    env.episode_len = 30
    ret, cost, length = trainer.evaluate(2, 45.0 * 0.1, 10 * 1)     # train_cdt.py:199-201
    assert np.isfinite(ret) and np.isfinite(cost) and length == 30
    _checkpoint_roundtrip(model, make())
