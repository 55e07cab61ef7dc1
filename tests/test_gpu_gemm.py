"""Unit test of the three GEMM kernel families against a plain PyTorch reference of the same op:
CUDA-core FFMA, 3xTF32 mma.sync, 3xTF32 tcgen05/TMEM.  fp64 matmul is the yardstick; the fp32 torch
result's own error sets the scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(2560, 256, 256), (5120, 400, 400), (640, 128, 64), (1024, 1024, 128), (515, 72, 100), (256, 400, 12),
          (256, 256, 256), (16, 32, 32), (300, 70, 41)]


@pytest.mark.parametrize("impl", ["ffma", "mma", "tc5", "fz"])
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


# (M, N, K, a_kc, b_kc): the thin shapes of an OSRL MLP (gemm_thin.cuh) next to tiled ones in every operand layout
LAYOUT_SHAPES = [
    (2560, 2048, 12, 1, 1),   # stacked first layer of 8 Q networks on 2560 rows (thin K)
    (256, 400, 10, 1, 1),
    (2560, 400, 2, 1, 0),     # last-layer dgrad: dH = dY[M,2] W[2,400] (thin K, B n-contiguous)
    (2560, 1, 256, 1, 1),     # Q head (thin N)
    (256, 8, 400, 1, 1),      # VAE mean/log-std heads
    (300, 12, 256, 1, 0),     # first-layer dgrad wrt the action input (thin N, B n-contiguous)
    (400, 10, 256, 0, 0),     # first-layer weight gradient: dW[H,in] = dH^T X (batch reduction, wide M)
    (2, 400, 2560, 0, 0),     # last-layer weight gradient: dW[out,H] = dY^T H (batch reduction, wide N)
    (16, 750, 300, 0, 0),
    (333, 16, 1000, 0, 0),
    (256, 256, 256, 1, 0),    # tiled dgrad / wgrad for comparison
    (400, 400, 256, 0, 0),
    (256, 256, 256, 0, 0),    # (fz: batch-256 weight gradient of a 256x256 layer, colsum = bias gradient)
    (32, 32, 16, 0, 0),
    (130, 70, 100, 0, 1),
    (100, 48, 300, 1, 0),
]


@pytest.mark.parametrize("impl", ["ffma", "mma", "tc5", "fz"])
def test_gemm_layouts_match_torch(lib_built, impl):
    from osrl_b200 import Engine
    eng = Engine("bc", batch_size=8, device=0, state_dim=4, action_dim=2, a_hidden_sizes=[8, 8])
    g = torch.Generator().manual_seed(1)
    for (M, N, K, a_kc, b_kc) in LAYOUT_SHAPES:
        A = torch.randn(M, K, generator=g)
        B = torch.randn(N, K, generator=g)
        want = A.double() @ B.double().T
        wcs = A.double().sum(1)
        Ain = A if a_kc else A.T.contiguous()
        Bin = B if b_kc else B.T.contiguous()
        use_cs = not a_kc and not b_kc
        got = eng.debug_gemm(impl, Ain, Bin, bool(a_kc), bool(b_kc), colsum=use_cs)
        if use_cs:
            got, cs = got
            errc = float((cs.double() - wcs).abs().max()) / float(A.abs().sum(1).max())
            assert errc <= 2e-6, f"{impl} {M}x{N}x{K} colsum err {errc:.2e}"
        err = float((got.double() - want).abs().max()) / float(want.abs().max())
        err32 = float(((A @ B.T).double() - want).abs().max()) / float(want.abs().max())
        print(f"{impl} {M}x{N}x{K} a_kc={a_kc} b_kc={b_kc}: err {err:.2e} torch-fp32 {err32:.2e}")
        tol = 2e-6 * max(1.0, K / 1024) if impl == "ffma" else 2e-6   # ffma: one serial fp32 chain over K
        assert err <= max(tol, 4 * err32), f"{impl} {M}x{N}x{K} ({a_kc},{b_kc}): err {err:.2e} (torch fp32 {err32:.2e})"
    eng.close()
