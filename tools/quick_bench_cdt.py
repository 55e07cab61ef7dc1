"""Device-side timing of the CDT step at BASELINE configs[3] (B=2048, T=10); dev aid."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from osrl_b200 import Engine
from oracle import synth, cdt as ocdt
from osrl_b200.common.dataset import SequenceDataset
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1     # cdt_configs.py: 0.1 at all three dropout sites
torch.manual_seed(0)
orc = ocdt.CDTOracle(ocdt.CDTConfig(17, 6, 1.0))
eng = Engine("cdt", batch_size=B, device=0, seed=1, state_dim=17, action_dim=6, max_action=1.0, seq_len=10, episode_len=1000,
             embedding_dim=128, num_layers=3, num_heads=8, use_rew=1, use_cost=1, cost_transform=1, stochastic=1,
             target_entropy=-6.0, learning_rate=1e-4, lr_warmup_steps=500, loss_cost_weight=0.02,
             attention_dropout=drop, residual_dropout=drop, embedding_dropout=drop)
eng.load_params(orc.params)
d = synth.make_dataset(17, 6, 1000, 200, seed=0)
d["costs"] = (np.random.default_rng(1).random(d["costs"].shape[0]) < 0.03).astype(np.float32)  # ~30 per episode (< 70)
SequenceDataset(d, seq_len=10, reward_scale=0.1, cost_scale=1.0, cost_sample=True, cost_transform=lambda x: 70 - x).to_engine(eng)
eng.steps(5); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.steps(steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"algo": "cdt", "B": B, "dropout": drop, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "tflops": 305.8e9 * (B / 2048) / (ms * 1e-3) / 1e12,
                  "launches_per_step": eng.launches_per_step, "stats": eng.stats()}))
prof = eng.profile(3)
import collections
agg = collections.defaultdict(float)
for n, m, by, fl in prof: agg[n.split("<")[0]] += m
tot = sum(agg.values())
for n, m in sorted(agg.items(), key=lambda x: -x[1])[:8]: print(f"{n:24s} {m*1e3:10.1f} us {100*m/tot:5.1f}%")
if os.environ.get("OPS"):
    for i, (n, m, by, fl) in enumerate(prof):
        print(f"{i:3d} {n:34s} {m*1e3:9.1f} us  {fl/1e9:9.2f} GFLOP  {fl/max(m,1e-9)/1e9:8.1f} TFLOP/s" if fl else f"{i:3d} {n:34s} {m*1e3:9.1f} us")
