"""Back-to-back device time of single GEMM launches through osrl_debug_gemm (kernel tuning aid).
usage: OSRL_DEBUG_TIME=200 python tools/gemm_bench.py [impl]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OSRL_DEBUG_TIME", "200")
import torch
from osrl_b200 import Engine
impl = sys.argv[1] if len(sys.argv) > 1 else "tc5"
SHAPES = [  # (M, N, K, a_kc, b_kc) -- the BCQ-Lag step's GEMM shapes
    (256, 400, 10, 1, 1), (256, 400, 400, 1, 1), (256, 8, 400, 1, 1), (256, 2, 400, 1, 1), (2560, 400, 12, 1, 1),
    (2560, 400, 400, 1, 1), (2560, 2, 400, 1, 1), (2560, 2048, 12, 1, 1), (2560, 256, 256, 1, 1), (2560, 1, 256, 1, 1),
    (256, 400, 400, 1, 0), (400, 400, 256, 0, 0), (400, 10, 256, 0, 0), (8, 400, 256, 0, 0), (256, 400, 8, 1, 0),
    (256, 256, 256, 1, 1), (256, 256, 256, 1, 0), (256, 256, 256, 0, 0), (256, 12, 256, 1, 0),
]
eng = Engine("bc", batch_size=8, device=0, state_dim=4, action_dim=2, a_hidden_sizes=[8, 8])
g = torch.Generator().manual_seed(0)
for (M, N, K, a_kc, b_kc) in SHAPES:
    A = torch.randn(M, K, generator=g); B = torch.randn(N, K, generator=g)
    eng.debug_gemm(impl, A if a_kc else A.T.contiguous(), B if b_kc else B.T.contiguous(), bool(a_kc), bool(b_kc),
                   colsum=not a_kc and not b_kc)
