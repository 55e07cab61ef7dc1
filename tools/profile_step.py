"""Per-launch timing of one engine step (eager launches + CUDA events, warm caches)."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from osrl_b200 import Engine
from oracle import synth
from tests.helpers import make_oracle
algo = sys.argv[1] if len(sys.argv) > 1 else "bcql"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = dict(state_dim=8, action_dim=2, max_action=1.0, a_hidden_sizes=[256, 256], c_hidden_sizes=[256, 256],
           vae_hidden_sizes=400, sample_action_num=10, num_q=2, num_qc=2, actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3)
orc = make_oracle(algo, cfg, 0)
eng = Engine(algo, batch_size=B, device=0, seed=1, **cfg)
eng.load_params(orc.params)
eng.upload_dataset(synth.make_dataset(8, 2, 300, 200, seed=0), 0.1, 1.0)
eng.steps(20); torch.cuda.synchronize()
prof = eng.profile(30)
tot = sum(p[1] for p in prof)
print(f"total eager {tot*1e3:.1f} us over {len(prof)} launches")
for i, (n, ms, by, fl) in enumerate(prof):
    print(f"{i:3d} {n:34s} {ms*1e3:8.2f} us  {fl/ms/1e9 if ms>0 else 0:8.1f} GFLOP/s  {by/ms/1e6 if ms>0 else 0:8.1f} GB/s  flops {fl:.3g} bytes {by:.3g}")
agg = collections.defaultdict(float)
for n, ms, by, fl in prof: agg[n] += ms
for n, ms in sorted(agg.items(), key=lambda x: -x[1]): print(f"{n:36s} {ms*1e3:9.1f} us {100*ms/tot:5.1f}%")
