"""Quick device-side timing of osrl_steps on BASELINE configs[1] (BCQ-Lag B=256); dev aid, not bench.py."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from osrl_b200 import Engine
from oracle import synth
from tests.helpers import make_oracle

algo = sys.argv[1] if len(sys.argv) > 1 else "bcql"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 500
cfg = dict(state_dim=8, action_dim=2, max_action=1.0, a_hidden_sizes=[256, 256], c_hidden_sizes=[256, 256],
           vae_hidden_sizes=400, sample_action_num=10, num_q=2, num_qc=2, actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3)
if algo == "bc":
    cfg = dict(state_dim=28, action_dim=2, max_action=1.0, a_hidden_sizes=[256, 256], actor_lr=1e-3)
orc = make_oracle(algo, cfg, 0)
eng = Engine(algo, batch_size=B, device=0, seed=1, **cfg)
eng.load_params(orc.params)
data = synth.make_dataset(cfg["state_dim"], cfg["action_dim"], 300, 200, seed=0)
eng.upload_dataset(data, 0.1, 1.0)
eng.steps(20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.steps(steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"algo": algo, "B": B, "steps": steps, "ms_per_step": ms / steps, "steps_per_s": steps / ms * 1e3,
                  "launches_per_step": eng.launches_per_step, "stats": eng.stats()}))
