"""Per-tensor parity diagnostics (engine vs live oracle): grads, params, stats."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from tests.helpers import batch_tuple, load_golden, make_oracle
from osrl_b200 import Engine

case = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
z, meta = load_golden(case)
algo, B = meta["algo"], meta["B"]
orc = make_oracle(algo, meta["cfg"], meta["init_seed"])
init = {k: v.clone() for k, v in orc.params.items()}
eng = Engine(algo, batch_size=B, device=0, seed=7, **meta["cfg"])
eng.load_params(init)
rng = np.random.default_rng(meta["data_seed"])
cfg = meta["cfg"]
torch.manual_seed(meta["noise_seed"])
for s in range(steps):
    b = synth.make_batch(rng, B, cfg["state_dim"], cfg["action_dim"])
    prev = {k: v.clone() for k, v in orc.params.items()}
    ostats = orc.step(*batch_tuple(algo, b))
    eng.step(b, {k: v for k, v in orc.last_noise.items() if k in eng.noise_layout})
    got = eng.stats()
    print(f"--- {case} step {s}")
    for k, w in ostats.items():
        print(f"  stat {k:28s} eng {got[k]: .8e} ref {w: .8e} rel {abs(got[k]-w)/max(abs(w),1e-12):.2e}")
    G = eng.read_section("grad")
    P = eng.read_params()
    for k, g in orc.last_grads.items():
        ge = G[k]
        den = g.abs().max().item() + 1e-30
        gerr = (ge - g).abs().max().item() / den
        d_ref = orc.params[k].detach() - prev[k]
        d_eng = P[k] - prev[k] if s == 0 else None
        perr = (P[k] - orc.params[k].detach()).abs()
        dmax = d_ref.abs().max().item() + 1e-30
        frac = (perr > 1e-5 * dmax).float().mean().item()
        print(f"  {k:34s} |g|max {den:.2e} gerr {gerr:.2e} | dP max {dmax:.2e} perr_max {perr.max().item()/dmax:.2e} "
              f"frac>1e-5 {frac:.4f} l2 {perr.norm().item()/ (d_ref.norm().item()+1e-30):.2e}")
