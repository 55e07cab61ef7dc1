"""Per-CTA clock64 timeline of the fused tcgen05 kernel on a few shapes (OSRL_FZ_DBG; kernel tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OSRL_FZ_DBG"] = "1"
os.environ.setdefault("OSRL_DEBUG_TIME", "50")
import torch
from osrl_b200 import Engine
eng = Engine("bc", batch_size=8, device=0, state_dim=4, action_dim=2, a_hidden_sizes=[8, 8])
g = torch.Generator().manual_seed(0)
for (M, N, K, a_kc, b_kc) in [(256, 256, 256, 1, 1), (256, 256, 256, 0, 0), (256, 256, 256, 1, 0), (2560, 256, 256, 1, 1)]:
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    print(f"--- {M}x{N}x{K} a_kc={a_kc} b_kc={b_kc}", file=sys.stderr)
    eng.debug_gemm("fz", A if a_kc else A.T.contiguous(), B if b_kc else B.T.contiguous(), bool(a_kc), bool(b_kc),
                   colsum=not a_kc and not b_kc)
