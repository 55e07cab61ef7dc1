"""Pin the oracle against the live reference and write tests/golden/*.npz.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden

For every algorithm it (1) seeds torch, builds the UNMODIFIED reference model+trainer
and the oracle, checks the initial parameters are bit-identical; (2) feeds both the
same seeded minibatches with torch's global generator re-seeded identically, so both
consume the same noise stream; (3) asserts per-step stats and final parameters agree;
(4) writes the batches, the consumed noise, the stats and the parameters (small
configs: full tensors; full-size configs: checksums) as fixtures.  Test
infrastructure only.
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import algos, ref_shim, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (algo, oracle cfg, batch, steps, store_full)
    "bc_small": ("bc", algos.BCConfig(28, 2, 1.0, [32, 32], 1e-3), 16, 3, True),
    "bcql_small": ("bcql", algos.BCQLConfig(8, 2, 1.0, [32, 32], [32, 32], 48, 10, num_q=2, num_qc=2,
                                            actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3), 16, 3, True),
    "cpq_small": ("cpq", algos.CPQConfig(11, 3, 1.0, [32, 32], [32, 32], 40, 10, beta=0.5, num_q=2, num_qc=2,
                                         episode_len=200, actor_lr=1e-4, critic_lr=1e-3, alpha_lr=1e-4,
                                         vae_lr=1e-3), 16, 3, True),
    "bearl_small": ("bearl", algos.BEARLConfig(8, 2, 1.0, [32, 32], [32, 32], 48, 10, num_q=2, num_qc=2,
                                               start_update_policy_step=0, actor_lr=1e-3, critic_lr=1e-3,
                                               vae_lr=1e-3), 16, 3, True),
    # BASELINE.json configs[0..2,4] at full layer sizes (checksums only)
    "bc_full": ("bc", algos.BCConfig(28, 2, 1.0, [256, 256], 1e-3), 256, 3, False),
    "bcql_full": ("bcql", algos.BCQLConfig(8, 2, 1.0, [256, 256], [256, 256], 400, 10, num_q=2, num_qc=2,
                                           actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3), 256, 3, False),
    "cpq_full": ("cpq", algos.CPQConfig(33, 8, 1.0, [256, 256], [256, 256], 400, 10, beta=0.5, num_q=2, num_qc=2,
                                        episode_len=200, actor_lr=1e-4, critic_lr=1e-3, alpha_lr=1e-4,
                                        vae_lr=1e-3), 512, 2, False),
    "bearl_full": ("bearl", algos.BEARLConfig(8, 2, 1.0, [256, 256], [256, 256], 400, 10, num_q=2, num_qc=2,
                                              start_update_policy_step=0, actor_lr=1e-3, critic_lr=1e-3,
                                              vae_lr=1e-3), 512, 2, False),
}

ORACLES = {"bc": algos.BCOracle, "bcql": algos.BCQLOracle, "cpq": algos.CPQOracle, "bearl": algos.BEARLOracle}
LR_KEYS = ("actor_lr", "critic_lr", "vae_lr", "alpha_lr")


def build_reference(osrl, algo: str, cfg):
    kw = {k: v for k, v in dataclasses.asdict(cfg).items() if k not in LR_KEYS}
    lrs = {k: v for k, v in dataclasses.asdict(cfg).items() if k in LR_KEYS}
    A = osrl.algorithms
    if algo == "bc":
        model = A.BC(**kw, device="cpu")
        trainer = A.BCTrainer(model, None, logger=ref_shim.NullLogger(), actor_lr=lrs["actor_lr"], device="cpu")
    elif algo == "bcql":
        model = A.BCQL(**kw, device="cpu")
        trainer = A.BCQLTrainer(model, None, logger=ref_shim.NullLogger(), **lrs, device="cpu")
    elif algo == "cpq":
        model = A.CPQ(**kw, device="cpu")
        trainer = A.CPQTrainer(model, None, logger=ref_shim.NullLogger(), **lrs, device="cpu")
    elif algo == "bearl":
        model = A.BEARL(**kw, device="cpu")
        trainer = A.BEARLTrainer(model, None, logger=ref_shim.NullLogger(), **lrs, device="cpu")
    else:
        raise ValueError(algo)
    return model, trainer


def batch_args(algo, b):
    t = {k: torch.from_numpy(v) for k, v in b.items()}
    if algo == "bc":
        return (t["observations"], t["actions"])
    return (t["observations"], t["next_observations"], t["actions"], t["rewards"], t["costs"], t["done"])


def checksum(t: torch.Tensor):
    d = t.double()
    return [float(d.sum()), float(d.abs().sum()), float((d * d).sum())]


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    den = float(b.abs().max()) + 1e-30
    return float((a - b).abs().max()) / den


def run_case(osrl, name, algo, cfg, B, steps, full):
    init_seed, data_seed, noise_seed = 0, 1234, 4321
    torch.manual_seed(init_seed)
    ref_model, ref_trainer = build_reference(osrl, algo, cfg)
    torch.manual_seed(init_seed)
    orc = ORACLES[algo](cfg)
    ref_sd = ref_model.state_dict()
    assert list(ref_sd.keys()) == list(orc.params.keys()), (name, "state_dict keys differ")
    for k, v in ref_sd.items():
        assert torch.equal(v, orc.params[k]), (name, "init differs", k)
    init = {k: v.clone() for k, v in orc.params.items()}

    rng = np.random.default_rng(data_seed)
    batches = [synth.make_batch(rng, B, cfg.state_dim, cfg.action_dim) for _ in range(steps)]

    torch.manual_seed(noise_seed)
    ref_stats = []
    for b in batches:
        n0 = len(ref_trainer.logger.rows)
        ref_trainer.train_one_step(*batch_args(algo, b))
        row = {}
        for r in ref_trainer.logger.rows[n0:]:
            row.update(r)
        ref_stats.append(row)

    torch.manual_seed(noise_seed)
    orc_stats, noises = [], []
    for b in batches:
        orc_stats.append(orc.step(*batch_args(algo, b)))
        noises.append({k: v.clone() for k, v in orc.last_noise.items()})

    worst = 0.0
    for s, (r, o) in enumerate(zip(ref_stats, orc_stats)):
        assert set(r) == set(o), (name, s, set(r) ^ set(o))
        for k in r:
            e = abs(r[k] - o[k]) / (abs(r[k]) + 1e-12)
            worst = max(worst, e)
            assert e < 2e-6, (name, s, k, r[k], o[k])
    perr = 0.0
    for k, v in ref_model.state_dict().items():
        perr = max(perr, rel_err(orc.params[k], v))
    assert perr < 2e-6, (name, "final params differ", perr)
    print(f"[golden] {name}: oracle == reference over {steps} steps (stat rel err {worst:.2e}, param rel err {perr:.2e})")

    out = {"meta": json.dumps({"algo": algo, "cfg": dataclasses.asdict(cfg), "B": B, "steps": steps,
                               "init_seed": init_seed, "data_seed": data_seed, "noise_seed": noise_seed,
                               "full": full, "torch": torch.__version__, "keys": list(init.keys()),
                               "stat_keys": sorted(ref_stats[0].keys())})}
    out["stats"] = np.array([[r[k] for k in sorted(r)] for r in ref_stats], dtype=np.float64)
    if full:
        for k, v in init.items():
            out["init/" + k] = v.numpy()
        for k, v in ref_model.state_dict().items():
            out["final/" + k] = v.numpy()
        for s, (b, nz) in enumerate(zip(batches, noises)):
            for k, v in b.items():
                out[f"batch{s}/{k}"] = v
            for k, v in nz.items():
                out[f"noise{s}/{k}"] = v.numpy()
    else:
        out["init_checksum"] = np.array([checksum(v) for v in init.values()])
        out["final_checksum"] = np.array([checksum(v) for v in ref_model.state_dict().values()])
        out["noise_checksum"] = np.array([[checksum(v) for v in nz.values()] for nz in noises])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    osrl = ref_shim.import_reference()
    only = sys.argv[1:]
    for name, (algo, cfg, B, steps, full) in CASES.items():
        if only and name not in only:
            continue
        run_case(osrl, name, algo, cfg, B, steps, full)


if __name__ == "__main__":
    main()
