"""Data-parallel equivalence worker (run under torch.distributed, gloo on CPU or nccl on GPUs).

N ranks x (B/N rows, rank-partitioned batch and noise)  ==  one rank on the concatenated batch
(SURVEY.md section 8e).  Backend gloo: oracle-with-DataParallel vs single oracle (host-side logic, CPU).
Backend nccl: the CUDA engine with its NCCL gradient all-reduce vs the single oracle.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import algos, synth  # noqa: E402
from tests.helpers import batch_tuple, l2rel, make_oracle  # noqa: E402

CFGS = {
    "bcql": dict(state_dim=8, action_dim=2, max_action=1.0, a_hidden_sizes=[32, 32], c_hidden_sizes=[32, 32],
                 vae_hidden_sizes=48, sample_action_num=10, num_q=2, num_qc=2, actor_lr=1e-3, critic_lr=1e-3,
                 vae_lr=1e-3, cost_limit=0.05),   # low threshold -> the PID multiplier is active
    "bearl": dict(state_dim=8, action_dim=2, max_action=1.0, a_hidden_sizes=[32, 32], c_hidden_sizes=[32, 32],
                  vae_hidden_sizes=48, sample_action_num=10, num_q=2, num_qc=2, actor_lr=1e-3, critic_lr=1e-3,
                  vae_lr=1e-3, start_update_policy_step=0, cost_limit=0.05),
    "bc": dict(state_dim=28, action_dim=2, max_action=1.0, a_hidden_sizes=[32, 32], actor_lr=1e-3),
    # CPQ: the 0.75-quantile of the OOD KL (cpq.py:183) is taken per shard; it only feeds log_alpha and a logged
    # term (no gradient), so the parameters still equal the single-rank step (stats are not compared here)
    "cpq": dict(state_dim=8, action_dim=2, max_action=1.0, a_hidden_sizes=[32, 32], c_hidden_sizes=[32, 32],
                vae_hidden_sizes=48, sample_action_num=10, num_q=2, num_qc=2, actor_lr=1e-3, critic_lr=1e-3,
                vae_lr=1e-3, alpha_lr=1e-3),
}
# noise slot -> rows per batch row (slots are [B*rows, cols] b-major, so a rank takes a contiguous block)
KEYS = ("observations", "next_observations", "actions", "rewards", "costs", "done")


def shard(x, rank, world):
    n = x.shape[0] // world
    return x[rank * n:(rank + 1) * n]


def run(algo: str, backend: str, steps: int = 3, Bg: int = 32, cfg=None, detail: bool = False):
    """detail=True: return {"ok", "worst_ratio", ...} instead of the bool (bench.py prints it as `dp_parity`)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = CFGS[algo] if cfg is None else cfg
    B = Bg // world
    full = make_oracle(algo, cfg, 0)              # single-rank reference on the concatenated batch
    init = {k: v.clone() for k, v in full.params.items()}
    rng = np.random.default_rng(11)
    torch.manual_seed(5)
    if backend == "nccl":
        from osrl_b200 import Engine, comm_unique_id
        dev = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(dev)
        eng = Engine(algo, batch_size=B, device=dev, seed=1, world_size=world, rank=rank, **cfg)
        eng.load_params(init)
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.init_comm(ids[0])
    else:
        part = make_oracle(algo, cfg, 0)
        part.dp = algos.DataParallel(dist)
    worst = 0.0
    for s in range(steps):
        b = synth.make_batch(rng, Bg, cfg["state_dim"], cfg["action_dim"])
        full.step(*batch_tuple(algo, b))
        nz_full = full.last_noise
        bl = {k: shard(v, rank, world) for k, v in b.items()}
        nl = {}
        for k, v in nz_full.items():
            if k.startswith("ood"):                  # CPQ's OOD draws are S-major: [S, B, .] (torch.tile, cpq.py:169-173)
                S = cfg["sample_action_num"]
                w_ = v.reshape(S, Bg, -1)[:, rank * B:(rank + 1) * B]
                nl[k] = w_.reshape(S, B, -1) if v.dim() == 3 else w_.reshape(S * B, -1)
                continue
            flat = v.reshape(Bg, -1) if v.shape[0] == Bg else v.reshape(Bg, -1)
            nl[k] = shard(flat, rank, world).reshape(-1, *v.shape[1:]) if v.dim() > 1 else shard(v, rank, world)
            if v.dim() >= 2 and v.shape[0] != Bg:   # [B*S, L] style: b-major blocks
                per = v.shape[0] // Bg
                nl[k] = v.reshape(Bg, per, *v.shape[1:])[rank * B:(rank + 1) * B].reshape(B * per, *v.shape[1:])
        if backend == "nccl":
            eng.step(bl, {k: v for k, v in nl.items() if k in eng.noise_layout})
            got = eng.read_params()
        else:
            part.step(*batch_tuple(algo, bl), noise=nl)
            got = {k: v.detach() for k, v in part.params.items()}
        for k, ref in full.params.items():
            ref = ref.detach()
            err = float((got[k] - ref).norm())
            bound = 1e-3 * float((ref - init[k]).norm()) + 4e-7 * float(ref.norm()) + 1e-9
            if err > bound and rank == 0:
                print(f"[dp {algo}] step {s} param {k}: err {err:.3e} bound {bound:.3e}", flush=True)
            worst = max(worst, err / bound)
    same = True
    mode = backend
    if backend == "nccl":   # replicas must be bit-identical (every rank applies the same summed gradient)
        mode = eng.dp_mode
        flat = torch.cat([got[k].reshape(-1).float().cuda() for k in sorted(got)])
        ref0 = flat.clone()
        dist.broadcast(ref0, src=0)
        same = bool(torch.equal(flat, ref0))
        if os.environ.get("OSRL_EXPECT_DP"):
            assert mode == os.environ["OSRL_EXPECT_DP"], f"dp mode {mode}, expected {os.environ['OSRL_EXPECT_DP']}"
    ok = torch.tensor([1.0 if worst <= 1.0 and same else 0.0])
    if backend == "nccl":
        ok = ok.cuda()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    wt = torch.tensor([worst], dtype=torch.float64)
    if backend == "nccl":
        wt = wt.cuda()
    dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    out = {"algo": algo, "backend": backend, "dp_mode": mode, "replicas_identical": same, "world": world, "global_batch": Bg,
           "steps": steps,
           "worst_ratio": float(wt.item()), "ok": bool(ok.item()),
           "bound": "final params vs the single-rank oracle on the concatenated batch: |d| <= 1e-3 |delta| + 4e-7 |p| per tensor"}
    if backend == "nccl":
        eng.close()
    if rank == 0 and not detail:
        print(json.dumps(out))
    return out if detail else bool(ok.item())


def run_pipelined(algo: str, steps: int = 5, Bg: int = 32):
    """osrl_steps(k) on a resident dataset shard under data parallelism: the pipelined graphs (VAE branch on its own
    communicator) must leave every rank in exactly the state k sequential single-step graphs do."""
    from osrl_b200 import Engine, comm_unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = CFGS[algo]
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    init = make_oracle(algo, cfg, 0).params
    data = synth.make_dataset(cfg["state_dim"], cfg["action_dim"], 50, 20, seed=100 + rank)   # this rank's shard
    engs = []
    for _ in range(2):
        eng = Engine(algo, batch_size=Bg // world, device=dev, seed=1, world_size=world, rank=rank, **cfg)
        eng.load_params(init)
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.init_comm(ids[0])
        eng.upload_dataset(data, 0.1, 1.0)
        engs.append(eng)
    for _ in range(steps):
        engs[0].steps(1)
    engs[1].steps(steps)
    torch.cuda.synchronize()
    same = True
    for sec in ("param", "target", "adam_m", "adam_v"):
        a, b = engs[0].read_section(sec), engs[1].read_section(sec)
        same = same and all(torch.equal(a[k], b[k]) for k in a)
    ok = torch.tensor([1.0 if same else 0.0]).cuda()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"algo": algo, "mode": "pipelined-vs-sequential", "world": world, "ok": bool(ok.item())}))
    for eng in engs:
        eng.close()
    return bool(ok.item())


def run_cdt(steps: int = 3, Bg: int = 16):
    """CDT under data parallelism: the masked means are means over the GLOBAL batch (valid-token counts differ per
    rank), so N ranks x B/N sequences must equal one oracle on the concatenated batch."""
    from oracle import cdt as ocdt
    from oracle.make_golden import CDT_KEYS, make_seq_batch
    from osrl_b200 import Engine, comm_unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    cfg = ocdt.CDTConfig(5, 3, 1.0, seq_len=10, episode_len=1000, embedding_dim=32, num_layers=2, num_heads=4,
                         learning_rate=1e-3, lr_warmup_steps=4)
    torch.manual_seed(0)
    full = ocdt.CDTOracle(cfg)
    init = {k: v.clone() for k, v in full.params.items()}
    B = Bg // world
    eng = Engine("cdt", batch_size=B, device=dev, seed=3, world_size=world, rank=rank, state_dim=5, action_dim=3,
                 max_action=1.0, seq_len=10, episode_len=1000, embedding_dim=32, num_layers=2, num_heads=4, use_rew=1,
                 use_cost=1, cost_transform=1, stochastic=1, target_entropy=-3.0, learning_rate=1e-3, lr_warmup_steps=4,
                 loss_cost_weight=cfg.loss_cost_weight, loss_state_weight=cfg.loss_state_weight,
                 weight_decay=cfg.weight_decay, betas=cfg.betas, clip_grad=cfg.clip_grad,
                 init_temperature=cfg.init_temperature)
    eng.load_params(init)
    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.init_comm(ids[0])
    rng = np.random.default_rng(3)
    worst = 0.0
    for s in range(steps):
        b = make_seq_batch(rng, Bg, cfg.seq_len, cfg.state_dim, cfg.action_dim)
        args = []
        for k in CDT_KEYS:
            t = torch.from_numpy(np.asarray(b[k]))
            args.append(t if k == "time_steps" else (t.float() if k != "mask" else t.double()))
        ref_stats = full.step(*args)
        eng.step_seq({k: shard(v, rank, world) for k, v in b.items()})
        got = eng.stats()
        for k, w in ref_stats.items():
            r = abs(got[k] - w) / (2e-5 * max(abs(w), 1e-3) + 1e-7)
            if r > 1.0 and rank == 0:
                print(f"[cdt dp] step {s} stat {k}: engine {got[k]} oracle {w} ratio {r:.1f}", flush=True)
            worst = max(worst, r)
    P = eng.read_params()
    for k, ref in full.params.items():
        if "in_proj_bias" in k:
            continue
        ref = ref.detach()
        err = float((P[k] - ref).norm())
        bound = 1e-3 * float((ref - init[k]).norm()) + 4e-7 * float(ref.norm()) + 1e-9
        if err > bound and rank == 0:
            print(f"[cdt dp] param {k}: err {err:.3e} bound {bound:.3e}", flush=True)
        worst = max(worst, err / bound)
    ok = torch.tensor([1.0 if worst <= 1.0 else 0.0]).cuda()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"algo": "cdt", "backend": "nccl", "world": world, "worst_ratio": worst, "ok": bool(ok.item())}))
    eng.close()
    return bool(ok.item())


def run_cdt_gloo(steps: int = 3, Bg: int = 8):
    """CPU restatement of the CDT data-parallel decomposition (oracle with partial sums + all-reduced counts and
    gradients on B/N sequences per rank) == the single oracle on the concatenated batch."""
    from oracle import cdt as ocdt
    from oracle.make_golden import CDT_KEYS, make_seq_batch
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = ocdt.CDTConfig(5, 3, 1.0, seq_len=10, episode_len=1000, embedding_dim=32, num_layers=2, num_heads=4,
                         learning_rate=1e-3, lr_warmup_steps=4)
    torch.manual_seed(0)
    full = ocdt.CDTOracle(cfg)
    torch.manual_seed(0)
    part = ocdt.CDTOracle(cfg)
    part.dp = dist
    init = {k: v.clone() for k, v in full.params.items()}
    rng = np.random.default_rng(3)
    worst = 0.0

    def args_of(b):
        out = []
        for k in CDT_KEYS:
            t = torch.from_numpy(np.asarray(b[k]))
            out.append(t if k == "time_steps" else (t.float() if k != "mask" else t.double()))
        return out

    for s in range(steps):
        b = make_seq_batch(rng, Bg, cfg.seq_len, cfg.state_dim, cfg.action_dim)
        ref = full.step(*args_of(b))
        got = part.step(*args_of({k: shard(v, rank, world) for k, v in b.items()}))
        for k, w in ref.items():
            worst = max(worst, abs(got[k] - w) / (2e-5 * max(abs(w), 1e-3) + 1e-7))
    for k, ref in full.params.items():
        if "in_proj_bias" in k:
            continue
        err = float((part.params[k] - ref).norm())
        bound = 1e-3 * float((ref - init[k]).norm()) + 4e-7 * float(ref.norm()) + 1e-9
        worst = max(worst, err / bound)
    ok = torch.tensor([1.0 if worst <= 1.0 else 0.0])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"algo": "cdt", "backend": "gloo", "world": world, "worst_ratio": worst, "ok": bool(ok.item())}))
    return bool(ok.item())


def _mp_entry(rank, world, algo, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    ok = run_cdt_gloo() if algo == "cdt" else run(algo, "gloo")
    if rank == 0:
        q.put(ok)
    dist.destroy_process_group()


if __name__ == "__main__":   # torchrun entry (GPU): python -m torch.distributed.run ... tests/dp_worker.py bcql
    algo_list = sys.argv[1:] or ["bcql"]
    dist.init_process_group("nccl")
    good = all(run_cdt() if a == "cdt" else (run_pipelined(a[5:]) if a.startswith("pipe:") else run(a, "nccl"))
               for a in algo_list)
    dist.destroy_process_group()
    sys.exit(0 if good else 1)
