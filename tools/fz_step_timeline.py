"""clock64 timeline of the slowest CTA of every fused launch of one BCQ-Lag step (OSRL_FZ_DBG; tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OSRL_FZ_DBG"] = "1"
os.environ["OSRL_PIPELINE"] = "0"
import ctypes as C
import numpy as np, torch
from osrl_b200 import Engine
from oracle import synth
from tests.helpers import make_oracle
cfg = dict(state_dim=8, action_dim=2, max_action=1.0, a_hidden_sizes=[256, 256], c_hidden_sizes=[256, 256],
           vae_hidden_sizes=400, sample_action_num=10, num_q=2, num_qc=2, actor_lr=1e-3, critic_lr=1e-3, vae_lr=1e-3)
orc = make_oracle("bcql", cfg, 0)
eng = Engine("bcql", batch_size=256, device=0, seed=1, **cfg)
eng.load_params(orc.params)
eng.upload_dataset(synth.make_dataset(8, 2, 300, 200, seed=0), 0.1, 1.0)
eng.steps(5); torch.cuda.synchronize()
eng.lib.osrl_debug_fz_timelines.argtypes = [C.c_void_p]
eng.lib.osrl_debug_fz_timelines(eng.h)
