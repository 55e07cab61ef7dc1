"""single-GPU probe of the dp_worker CPQ config: engine vs oracle gradients after one step, B = 16 and 32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from tests.helpers import batch_tuple, make_oracle, maxrel
from tests.dp_worker import CFGS
from osrl_b200 import Engine

for algo in ("cpq", "bcql"):
    cfg = CFGS[algo]
    for B in (16, 32):
        orc = make_oracle(algo, cfg, 0)
        eng = Engine(algo, batch_size=B, device=0, seed=1, **cfg)
        eng.load_params(orc.params)
        rng = np.random.default_rng(11)
        torch.manual_seed(5)
        b = synth.make_batch(rng, B, cfg["state_dim"], cfg["action_dim"])
        orc.step(*batch_tuple(algo, b))
        eng.step(b, {k: v for k, v in orc.last_noise.items() if k in eng.noise_layout})
        G = eng.read_section("grad")
        bad = [(k, maxrel(G[k], g)) for k, g in orc.last_grads.items() if maxrel(G[k], g) > 1e-4]
        print(algo, B, "noise slots", list(eng.noise_layout), "bad grads:", [(k, f"{v:.2e}") for k, v in bad][:12], flush=True)
        eng.close()
