"""Engine-vs-engine diff between two GEMM modes (OSRL_GEMM is read when the engine's program is built)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from tests.helpers import load_golden, make_oracle
from osrl_b200 import Engine

case = sys.argv[1] if len(sys.argv) > 1 else "bcql_full"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
modes = (sys.argv[3] if len(sys.argv) > 3 else "mma,tc5").split(",")
z, meta = load_golden(case)
algo, B, cfg = meta["algo"], meta["B"], meta["cfg"]
orc = make_oracle(algo, cfg, meta["init_seed"])
data = synth.make_dataset(cfg["state_dim"], cfg["action_dim"], 300, 200, seed=0)
engs = []
for m in modes:
    if m.startswith("pack"):          # "pack0" / "pack1": tc5 with the packed-image path off / on
        os.environ["OSRL_GEMM"] = "tc5"; os.environ["OSRL_PACK"] = m[4:]
    else:
        os.environ["OSRL_GEMM"] = m
    e = Engine(algo, batch_size=B, device=0, seed=7, **cfg)
    e.load_params(orc.params)
    e.upload_dataset(data, 0.1, 1.0)
    engs.append(e)
for s in range(steps):
    for e in engs:
        e.steps(1)
    print(f"--- step {s}")
    for sec in ("grad", "param"):
        X = [e.read_section(sec) if sec == "grad" else e.read_params() for e in engs]
        for k in X[0]:
            a, b = X[0][k].double(), X[1][k].double()
            den = a.abs().max().item() + 1e-30
            err = (a - b).abs().max().item() / den
            if err > 2e-5:
                print(f"  {sec:5s} {k:36s} max {den:.2e} rel diff {err:.2e}")
    st = [e.stats() for e in engs]
    for k in st[0]:
        d = abs(st[0][k] - st[1][k]) / max(abs(st[0][k]), 1e-12)
        if d > 1e-5: print(f"  stat {k} {st[0][k]} {st[1][k]}")
