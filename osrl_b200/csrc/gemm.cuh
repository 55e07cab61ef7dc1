// Multi-task fp32 GEMM for the MLP layers of the OSRL step (sm_100a).
//
// One launch executes a list of independent C[M,N] = op(A)[M,K] * op(B)[K,N] problems
// ("tasks"): the forward of an ensemble layer, or the dgrad + wgrad of one layer, or the
// same layer of several networks.  Every task carries its own fused epilogue (bias,
// ReLU/Tanh, scale, residual add, clamp, activation-derivative mask, bias-gradient column
// sum) so no elementwise kernel runs between layers.  Replaces nn.Linear + activation in
// the reference's mlp() (osrl/common/net.py:12-30) and their autograd backward.
//
// Arithmetic: fp32 FFMA with fp32 accumulation -- the parity mode (1e-5 vs the reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace osrl {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct GemmTask {
  const float* A;
  const float* B;
  float* C;
  const float* bias;      // [N], added per output column (forward)
  const float* resid;     // [M, ldr] added after act*scale (perturbation actor: + act)
  const float* dact_src;  // [M, ld_dact] stored activation whose derivative masks C (dgrad)
  float* aux;             // [M, ldaux] optional copy of act(acc + bias) before scale/resid
  float* colsum;          // [M] sum_k A[i,k] (bias gradient when A = dY^T), written by tn==0 tiles
  int M, N, K;
  int lda, ldb, ldc, ldr, ld_dact, ldaux;
  int a_kc;               // 1: A[i*lda + k] (k contiguous); 0: A[k*lda + i]
  int b_kc;               // 1: B[j*ldb + k];                 0: B[k*ldb + j]
  int act;                // Act applied to acc + bias
  float scale;            // multiplies the activated value
  int clamp;              // clamp final value to [lo, hi]
  float lo, hi;
  int dact;               // 0 none, ACT_RELU: *= (src > 0), ACT_TANH: *= 1 - src^2
  int tile0;              // index of this task's first tile in the launch
  int tiles_n;            // tiles along N
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gemm_tasks(const GemmTask* __restrict__ tasks, int ntasks) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int EA = BM * BK / NT;
  constexpr int EB = BN * BK / NT;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  __shared__ GemmTask ts;

  const int tid = threadIdx.x;
  if (tid == 0) {
    int ti = 0;
    const int tile = blockIdx.x;
    while (ti + 1 < ntasks && tasks[ti + 1].tile0 <= tile) ++ti;
    ts = tasks[ti];
  }
  __syncthreads();
  const GemmTask& t = ts;
  const int lt = blockIdx.x - t.tile0;
  const int m0 = (lt / t.tiles_n) * BM;
  const int n0 = (lt % t.tiles_n) * BN;
  const int M = t.M, N = t.N, K = t.K;
  const float* __restrict__ A = t.A;
  const float* __restrict__ B = t.B;
  const int lda = t.lda, ldb = t.ldb;
  const bool akc = t.a_kc != 0, bkc = t.b_kc != 0;

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float rs[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) rs[i] = 0.f;
  const bool want_colsum = (t.colsum != nullptr) && (n0 == 0) && (tx == 0);

  float ra[EA], rb[EB];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int r = 0; r < EA; ++r) {
      const int e = tid + r * NT;
      int i, k;
      if (akc) { i = e / BK; k = e % BK; } else { i = e % BM; k = e / BM; }
      const int gi = m0 + i, gk = k0 + k;
      float v = 0.f;
      if (gi < M && gk < K) v = akc ? A[(size_t)gi * lda + gk] : A[(size_t)gk * lda + gi];
      ra[r] = v;
    }
#pragma unroll
    for (int r = 0; r < EB; ++r) {
      const int e = tid + r * NT;
      int j, k;
      if (bkc) { j = e / BK; k = e % BK; } else { j = e % BN; k = e / BN; }
      const int gj = n0 + j, gk = k0 + k;
      float v = 0.f;
      if (gj < N && gk < K) v = bkc ? B[(size_t)gj * ldb + gk] : B[(size_t)gk * ldb + gj];
      rb[r] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int r = 0; r < EA; ++r) {
      const int e = tid + r * NT;
      int i, k;
      if (akc) { i = e / BK; k = e % BK; } else { i = e % BM; k = e / BM; }
      As[buf][k][i] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < EB; ++r) {
      const int e = tid + r * NT;
      int j, k;
      if (bkc) { j = e / BK; k = e % BK; } else { j = e % BN; k = e / BN; }
      Bs[buf][k][j] = rb[r];
    }
  };

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[cur][k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[cur][k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (want_colsum) {
#pragma unroll
        for (int i = 0; i < TM; ++i) rs[i] += a[i];
      }
    }
    if (kt + 1 < nk) {
      store_tiles(cur ^ 1);
      __syncthreads();
    }
  }

  // ---- fused epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gi = m0 + ty * TM + i;
    if (gi >= M) continue;
    if (want_colsum) t.colsum[gi] = rs[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gj = n0 + tx * TN + j;
      if (gj >= N) continue;
      float v = acc[i][j];
      if (t.bias) v += t.bias[gj];
      v = apply_act(v, t.act);
      if (t.aux) t.aux[(size_t)gi * t.ldaux + gj] = v;
      v *= t.scale;
      if (t.resid) v += t.resid[(size_t)gi * t.ldr + gj];
      if (t.clamp) v = fminf(fmaxf(v, t.lo), t.hi);
      if (t.dact) {
        const float s = t.dact_src[(size_t)gi * t.ld_dact + gj];
        v = (t.dact == ACT_RELU) ? (s > 0.f ? v : 0.f) : v * (1.f - s * s);
      }
      t.C[(size_t)gi * t.ldc + gj] = v;
    }
  }
}

}  // namespace osrl
