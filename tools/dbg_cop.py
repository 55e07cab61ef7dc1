import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses, numpy as np, torch, torch.nn.functional as F
from oracle import coptidice as oc, synth, core as C
from osrl_b200 import Engine
KEYS = ("observations", "next_observations", "actions", "rewards", "costs", "done", "is_init")
cfg = oc.COptiDICEConfig(8, 2, 1.0, f_type="chi2", init_state_propotion=0.2, a_hidden_sizes=[256, 256], c_hidden_sizes=[256, 256], num_nu=2, num_chi=2, actor_lr=1e-3, critic_lr=1e-3, scalar_lr=1e-3)
B = 256
rng = np.random.default_rng(3)
obs_std = rng.uniform(0.5, 1.5, 8).astype(np.float32); act_std = rng.uniform(0.3, 0.8, 2).astype(np.float32)
torch.manual_seed(0)
orc = oc.COptiDICEOracle(cfg, obs_std[None], act_std[None])
os.environ["OSRL_GEMM"] = sys.argv[1] if len(sys.argv) > 1 else "ffma"
eng = Engine("coptidice", batch_size=B, device=0, seed=5, observations_std=obs_std, actions_std=act_std, **dataclasses.asdict(cfg))
eng.load_params({k: v for k, v in orc.params.items() if k not in ("tau", "lmbda")})
torch.manual_seed(11)
for s in range(2):
    b = synth.make_batch(rng, B, 8, 2); b["is_init"] = (rng.random(B) < 0.2).astype(np.float32)
    P0 = eng.read_params()
    for k in P0:
        d = float((P0[k] - orc.params[k]).abs().max())
        if d > 1e-6: print("step", s, "param diff before step", k, d)
    orc.step(*[torch.from_numpy(b[k]) for k in KEYS])
    eng.step(b, {k: v for k, v in orc.last_noise.items() if k in eng.noise_layout})
    G = eng.read_section("grad")
    for k, g in orc.last_grads.items():
        if k in ("tau", "lmbda"): continue
        err = float((G[k] - g).abs().max()) / (float(g.abs().max()) + 1e-30)
        if err > 2e-5: print("step", s, k, "err", err, "gmax", float(g.abs().max()))
    k = "nu_network.q_nets.1.4.weight"
    d = (G[k] - orc.last_grads[k]).flatten()
    print("step", s, k, "diff nonzeros", int((d.abs() > 1e-7).sum()), "of", d.numel(), "max", float(d.abs().max()))
