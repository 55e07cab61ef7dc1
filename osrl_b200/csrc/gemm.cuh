// Multi-task fp32 GEMM for the MLP layers of the OSRL step (sm_100a).
//
// One launch executes a list of independent C[M,N] = op(A)[M,K] * op(B)[K,N] problems
// ("tasks"): the forward of an ensemble layer, or the dgrad + wgrad of one layer, or the
// same layer of several networks.  Every task carries its own fused epilogue (bias,
// ReLU/Tanh, scale, residual add, clamp, activation-derivative mask, bias-gradient column
// sum) so no elementwise kernel runs between layers.  Replaces nn.Linear + activation in
// the reference's mlp() (osrl/common/net.py:12-30) and their autograd backward.
//
// The layers of this workload are small (M = 256..5120, N,K <= 1024) and every operand is
// L2-resident, so a tile's k-loop is bound by load latency, not bandwidth: operands are
// staged global->shared with cp.async (LDGSTS) through an NSTAGE-deep ring so several
// k-slabs are in flight per CTA, in their native layout (no transpose on the way in):
//   k-contiguous operand  X[i*ld + k]  ->  smem [rows][BK+4]   (float4 fragments along k)
//   mn-contiguous operand X[k*ld + i]  ->  smem [BK][rows+4]   (float4 fragments along m/n)
// Arithmetic: fp32 FFMA with fp32 accumulation -- the parity mode (1e-5 vs the reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace osrl {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_GELU = 3 };

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
  int a_vec, b_vec;       // 16-byte cp.async allowed (base, ld and contiguous extent 4-float aligned)
  int act;                // Act applied to acc + bias
  float scale;            // multiplies the activated value
  int clamp;              // clamp final value to [lo, hi]
  float lo, hi;
  int dact;               // 0 none, ACT_RELU: *= (src > 0), ACT_TANH: *= 1 - src^2
  int tile0;              // index of this task's first tile in the launch
  int tiles_n;            // tiles along N
  int tiles_mn;           // tiles of one k-split (tiles_m * tiles_n)
  int ksplit;             // >1: split-K -- each split atomically adds its partial into a pre-zeroed C / colsum
  int klen;               // k extent of one split (multiple of every BK)
  int thin;               // ThinKind (gemm_thin.cuh): 0 = tiled kernels, else which thin body runs the task
  // ---- packed tf32 hi/lo operands (gemm_tc5.cuh, "A-packed" path).  A producer task with pk_hi set also writes
  // its output, split into hi = tf32(x) and lo = tf32(x - hi), as ready-made SWIZZLE_128B shared-memory images:
  // blocks of 64 rows x 32 k (8 KB), block (rb, ks) at float offset (rb * pk_ks + ks) * 2048, one image set per
  // group of pk_gcols output columns (= one consumer network of a stacked ensemble layer).  The consumer GEMM then
  // fetches its A tiles with cp.async.bulk instead of loading, splitting and storing them in every column tile.
  float* pk_hi; float* pk_lo;
  const float* a_hi; const float* a_lo;   // consumer side: packed images of A (null: load + split A from t.A)
  int pk_gcols;           // producer: columns per group (the consumer's K); 0 = no packed output wanted
  int pk_gstride;         // floats between the image sets of consecutive groups
  int pk_ks;              // k-slabs (of 32) per row block = ceil(pk_gcols / 32); consumer: same value
  int c_dead;             // producer hint: nothing reads C except through the packed images -> skip the fp32 store
  const float* mmask;     // [M, ldmm] optional multiplier applied after act*scale, before the residual add (dropout)
  int ldmm;
  int fz_pending;         // > 0: marker for Engine::fz_pending[fz_pending - 1] -- this task only carries the last layer's
                          // epilogue of a fused network (gemm_fz.cuh); emit_gemm completes and launches the fused task
};
__host__ __device__ __forceinline__ size_t pk_offset(int row, int col, int ks_per_rb) {   // float offset in an image set
  const int rb = row >> 6, r = row & 63, ks = col >> 5, c = col & 31;
  return ((size_t)rb * ks_per_rb + ks) * 2048 + (size_t)((r >> 3) * 256 + (r & 7) * 32 + ((((c >> 2) ^ (r & 7))) << 2) + (c & 3));
}

// A launch's task list travels as a kernel parameter (constant bank), not through global memory: finding the
// owner of a tile and reading its fields then costs no L2 round trip and no CTA barrier -- on the thin / small
// problems of this workload the old lookup (scan tile0 in global memory, copy the descriptor to shared memory)
// was 1-2 us of a 5-10 us kernel.  Fields are read straight from the parameter (uniform loads the compiler may
// hoist, since parameter space is read-only).  Longer lists are split over several launches by the host.
constexpr int PACK_MAX = 16;
struct TaskPack { GemmTask t[PACK_MAX]; };
static_assert(sizeof(TaskPack) <= 4000, "task pack must fit the 4 KB kernel parameter space");

// tile0 is increasing over the list: the owner of `tile` is the last task with tile0 <= tile
__device__ __forceinline__ int find_task(const TaskPack& P, int ntasks, int tile) {
  int ti = 0;
#pragma unroll
  for (int i = 1; i < PACK_MAX; ++i)
    if (i < ntasks && P.t[i].tile0 <= tile) ti = i;
  return ti;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

// ---- "FULL" epilogue extras (CDT): exact-erf GELU and its derivative, kept out of line and compiled only into
// the FULL kernel variants -- inlined into the common epilogue they get if-converted and cost the ReLU/Tanh
// layers ~12 % (measured on B200).
static __device__ __noinline__ float gelu_fwd(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
static __device__ __noinline__ float gelu_bwd(float s) {
  return 0.5f * (1.f + erff(s * 0.70710678118654752f)) + s * 0.3989422804014327f * expf(-0.5f * s * s);
}

// The fused epilogue.  A task's fields live in the kernel's parameter space behind a runtime task index, so
// reading them per output element costs a dozen indexed constant loads and branches per element -- measured at
// half of the tcgen05 kernel's samples.  Every thread therefore copies what the epilogue needs into registers
// once (Epi) and loads the bias of its columns once; the per-element path is then a handful of predicated ops.
// FULL adds GELU (aux keeps the PRE-activation, the dgrad mask reads it back) and split-K accumulation.
struct Epi {
  float* C; float* aux; const float* resid; const float* dsrc; const float* bias; const float* mmask;
  int ldc, ldaux, ldr, ldd, ldmm, act, clamp, dact, ksplit;
  float scale, lo, hi;
};
__device__ __forceinline__ Epi make_epi(const GemmTask& t) {
  Epi e;
  e.C = t.C; e.aux = t.aux; e.resid = t.resid; e.dsrc = t.dact_src; e.bias = t.bias; e.mmask = t.mmask;
  e.ldc = t.ldc; e.ldaux = t.ldaux; e.ldr = t.ldr; e.ldd = t.ld_dact; e.ldmm = t.ldmm;
  e.act = t.act; e.clamp = t.clamp; e.dact = t.dact; e.ksplit = t.ksplit;
  e.scale = t.scale; e.lo = t.lo; e.hi = t.hi;
  return e;
}
__device__ __forceinline__ float epi_bias(const Epi& e, int gj, int N) { return (e.bias && gj < N) ? e.bias[gj] : 0.f; }
// final value of one output element (stores aux on the way); `bias` = epi_bias of the element's column
template <bool FULL>
__device__ __forceinline__ float epi_value(const Epi& e, float bias, int gi, int gj, float v) {
  v += bias;
  if (FULL && e.act == ACT_GELU) {
    if (e.aux) e.aux[(size_t)gi * e.ldaux + gj] = v;
    v = gelu_fwd(v);
  } else {
    v = apply_act(v, e.act);
    if (e.aux) e.aux[(size_t)gi * e.ldaux + gj] = v;
  }
  v *= e.scale;
  if (FULL && e.mmask) v *= e.mmask[(size_t)gi * e.ldmm + gj];
  if (e.resid) v += e.resid[(size_t)gi * e.ldr + gj];
  if (e.clamp) v = fminf(fmaxf(v, e.lo), e.hi);
  if (e.dact) {
    const float s = e.dsrc[(size_t)gi * e.ldd + gj];
    if (FULL && e.dact == ACT_GELU) v *= gelu_bwd(s);
    else v = (e.dact == ACT_RELU) ? (s > 0.f ? v : 0.f) : v * (1.f - s * s);
  }
  return v;
}
template <bool FULL>
__device__ __forceinline__ void epi_store(const Epi& e, float bias, int gi, int gj, float v) {
  if (FULL && e.ksplit > 1) { atomicAdd(&e.C[(size_t)gi * e.ldc + gj], v); return; }
  e.C[(size_t)gi * e.ldc + gj] = epi_value<FULL>(e, bias, gi, gj, v);
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// load T consecutive floats from shared memory (T in {1,2,4,8}), address aligned to min(T,4) floats
template <int T>
__device__ __forceinline__ void lds_vec(const float* p, float* out) {
  if constexpr (T == 1) {
    out[0] = p[0];
  } else if constexpr (T == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    out[0] = v.x; out[1] = v.y;
  } else {
#pragma unroll
    for (int q = 0; q < T / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
      out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN, int NSTAGE>
struct GemmCfg {
  static constexpr int NT = (BM / TM) * (BN / TN);
  static constexpr int A_KC = BM * (BK + 4), A_MC = BK * (BM + 4);
  static constexpr int B_KC = BN * (BK + 4), B_MC = BK * (BN + 4);
  static constexpr int A_STAGE = A_KC > A_MC ? A_KC : A_MC;
  static constexpr int B_STAGE = B_KC > B_MC ? B_KC : B_MC;
  static constexpr int SMEM_BYTES = NSTAGE * (A_STAGE + B_STAGE) * (int)sizeof(float);
};

// stage one operand slab: ROWS x BK elements starting at (r0, k0)
template <int ROWS, int BK, int NT>
__device__ __forceinline__ void stage_operand(float* __restrict__ s, const float* __restrict__ G, int ld, bool kc,
                                              bool vec, int r0, int k0, int R, int K, int tid) {
  if (kc) {  // G[r*ld + k] -> s[r*(BK+4) + k]
    if (vec) {
      constexpr int CH = ROWS * (BK / 4);
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        const int r = c / (BK / 4), kq = (c % (BK / 4)) * 4;
        const bool ok = (r0 + r < R) && (k0 + kq < K);
        cp_async16(s + r * (BK + 4) + kq, ok ? G + (size_t)(r0 + r) * ld + k0 + kq : G, ok);
      }
    } else {
      constexpr int EL = ROWS * BK;
#pragma unroll
      for (int e = tid; e < EL; e += NT) {
        const int r = e / BK, k = e % BK;
        const bool ok = (r0 + r < R) && (k0 + k < K);
        cp_async4(s + r * (BK + 4) + k, ok ? G + (size_t)(r0 + r) * ld + k0 + k : G, ok);
      }
    }
  } else {  // G[k*ld + r] -> s[k*(ROWS+4) + r]
    if (vec) {
      constexpr int CH = BK * (ROWS / 4);
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        const int k = c / (ROWS / 4), rq = (c % (ROWS / 4)) * 4;
        const bool ok = (k0 + k < K) && (r0 + rq < R);
        cp_async16(s + k * (ROWS + 4) + rq, ok ? G + (size_t)(k0 + k) * ld + r0 + rq : G, ok);
      }
    } else {
      constexpr int EL = ROWS * BK;
#pragma unroll
      for (int e = tid; e < EL; e += NT) {
        const int k = e / ROWS, r = e % ROWS;
        const bool ok = (k0 + k < K) && (r0 + r < R);
        cp_async4(s + k * (ROWS + 4) + r, ok ? G + (size_t)(k0 + k) * ld + r0 + r : G, ok);
      }
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN, int NSTAGE, bool FULL>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gemm_tasks(const __grid_constant__ TaskPack P, int ntasks) {
  using Cfg = GemmCfg<BM, BN, BK, TM, TN, NSTAGE>;
  constexpr int NT = Cfg::NT;
  constexpr int TXN = BN / TN;  // threads along N
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + NSTAGE * Cfg::A_STAGE;

  const int tid = threadIdx.x;
  const GemmTask& t = P.t[find_task(P, ntasks, blockIdx.x)];
  int lt = blockIdx.x - t.tile0, kbeg = 0;
  if constexpr (FULL) { kbeg = (lt / t.tiles_mn) * t.klen; lt %= t.tiles_mn; }
  const int m0 = (lt / t.tiles_n) * BM;
  const int n0 = (lt % t.tiles_n) * BN;
  const int M = t.M, N = t.N, K = FULL ? min(t.K, kbeg + t.klen) : t.K;   // K = end of this CTA's k range
  const float* __restrict__ A = t.A;
  const float* __restrict__ B = t.B;
  const int lda = t.lda, ldb = t.ldb;
  const bool akc = t.a_kc != 0, bkc = t.b_kc != 0;
  const bool avec = t.a_vec != 0, bvec = t.b_vec != 0;

  const int tx = tid % TXN;
  const int ty = tid / TXN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float rs[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) rs[i] = 0.f;
  const bool want_colsum = (t.colsum != nullptr) && (n0 == 0) && (tx == 0);

  const int nk = (K - kbeg + BK - 1) / BK;
  // prologue: NSTAGE-1 slabs in flight
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nk) {
      stage_operand<BM, BK, NT>(As + s * Cfg::A_STAGE, A, lda, akc, avec, m0, kbeg + s * BK, M, K, tid);
      stage_operand<BN, BK, NT>(Bs + s * Cfg::B_STAGE, B, ldb, bkc, bvec, n0, kbeg + s * BK, N, K, tid);
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();  // slab kt visible to all; slab kt-1's buffer is free
    {
      const int nx = kt + NSTAGE - 1;
      if (nx < nk) {
        const int sb = nx % NSTAGE;
        stage_operand<BM, BK, NT>(As + sb * Cfg::A_STAGE, A, lda, akc, avec, m0, kbeg + nx * BK, M, K, tid);
        stage_operand<BN, BK, NT>(Bs + sb * Cfg::B_STAGE, B, ldb, bkc, bvec, n0, kbeg + nx * BK, N, K, tid);
      }
      cp_async_commit();
    }
    const float* __restrict__ as = As + (kt % NSTAGE) * Cfg::A_STAGE;
    const float* __restrict__ bs = Bs + (kt % NSTAGE) * Cfg::B_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      float a[4][TM], b[4][TN];
      if (akc) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(as + (ty * TM + i) * (BK + 4) + k4);
          a[0][i] = v.x; a[1][i] = v.y; a[2][i] = v.z; a[3][i] = v.w;
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) lds_vec<TM>(as + (k4 + kk) * (BM + 4) + ty * TM, a[kk]);
      }
      if (bkc) {  // strided column ownership: col = tx + j*TXN (bank-conflict-free float4 reads along k)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(bs + (tx + j * TXN) * (BK + 4) + k4);
          b[0][j] = v.x; b[1][j] = v.y; b[2][j] = v.z; b[3][j] = v.w;
        }
      } else {    // consecutive column ownership: col = tx*TN + j
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) lds_vec<TN>(bs + (k4 + kk) * (BN + 4) + tx * TN, b[kk]);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[kk][i], b[kk][j], acc[i][j]);
        if (want_colsum) {
#pragma unroll
          for (int i = 0; i < TM; ++i) rs[i] += a[kk][i];
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---- fused epilogue
  const Epi ep = make_epi(t);
  float bj[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bj[j] = epi_bias(ep, n0 + (bkc ? tx + j * TXN : tx * TN + j), N);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gi = m0 + ty * TM + i;
    if (gi >= M) continue;
    if (want_colsum) {
      if (FULL && t.ksplit > 1) atomicAdd(&t.colsum[gi], rs[i]);
      else t.colsum[gi] = rs[i];
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gj = n0 + (bkc ? tx + j * TXN : tx * TN + j);
      if (gj >= N) continue;
      epi_store<FULL>(ep, bj[j], gi, gj, acc[i][j]);
    }
  }
}

}  // namespace osrl
