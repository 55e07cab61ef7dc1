"""Unit test of the three GEMM kernel families against a plain PyTorch reference of the same op:
CUDA-core FFMA, 3xTF32 mma.sync, 3xTF32 tcgen05/TMEM.  fp64 matmul is the yardstick; the fp32 torch
result's own error sets the scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(2560, 256, 256), (5120, 400, 400), (640, 128, 64), (1024, 1024, 128), (515, 72, 100), (256, 400, 12)]


@pytest.mark.parametrize("impl", ["ffma", "mma", "tc5"])
def test_linear_matches_torch(lib_built, impl):
    from osrl_b200 import Engine
    eng = Engine("bc", batch_size=8, device=0, state_dim=4, action_dim=2, a_hidden_sizes=[8, 8])
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in SHAPES:
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g)
        for act in (0, 1, 2):
            pre64 = A.double() @ W.double().T + b.double()
            want64 = {0: pre64, 1: pre64.relu(), 2: pre64.tanh()}[act]
            got = eng.debug_linear(impl, A, W, b, act).double()
            scale = float(pre64.abs().max())     # errors are made on the pre-activation (ReLU / Tanh are 1-Lipschitz)
            err = float((got - want64).abs().max()) / scale
            ref32 = torch.nn.functional.linear(A, W, b)
            ref32 = {0: ref32, 1: ref32.relu(), 2: ref32.tanh()}[act].double()
            err32 = float((ref32 - want64).abs().max()) / scale
            # tcgen05 accumulates all K/8*3 MMAs of a tile in TMEM (truncating adder, no intermediate flush)
            tol = 2e-6
            print(f"{impl} {M}x{N}x{K} act {act}: err {err:.2e} torch-fp32 {err32:.2e}")
            assert err <= max(tol, 4 * err32), f"{impl} {M}x{N}x{K} act {act}: err {err:.2e} (torch fp32 {err32:.2e})"
    eng.close()
