// "Thin" members of the multi-task GEMM family: problems where one extent is <= 16 (sm_100a).
//
// An OSRL MLP has a first layer with K = obs_dim + act_dim (10..40 here, often <= 16 after the VAE latent), and a
// last layer with 1..8 outputs, so about half of the GEMM launches of a step are not matrix-shaped at all: they
// are bandwidth / latency bound outer products, row dot-products and batch reductions.  On a 32x32 tensor-core
// tile they spend 6..40 us per launch waiting on a 1..13-slab dependent chain; here they are plain fp32 FFMA
// loops with coalesced traffic and a CTA count that scales with the long extent.  Same task list and fused
// epilogue (epilogue_store<false>) as gemm.cuh, exact fp32 arithmetic.
//
//   THIN_K  (K <= 16, any layouts)           C[M,N]: thread per column, 16 rows per CTA, A tile in smem,
//                                            B row in registers.  first layers (forward), last-layer dgrad.
//   THIN_N  (N <= 16, A k-contiguous)        warp per row, lanes stride k, shuffle reduction.
//                                            last layers (forward), first-layer dgrad (d loss / d action).
//   THIN_R  (M or N <= 16, both operands     batch reduction  C[m,n] = sum_k A[k,m] B[k,n]: thread per wide index,
//            mn-contiguous)                  16 k-groups per CTA, fixed-order smem reduction (deterministic).
//                                            weight gradients of first and last layers (+ bias gradient colsum).
#pragma once
#include "gemm.cuh"

namespace osrl {

enum ThinKind { THIN_NONE = 0, THIN_K = 1, THIN_N = 2, THIN_R_WIDE_M = 3, THIN_R_WIDE_N = 4 };
constexpr int THIN_THREADS = 256;
constexpr int THIN_K_ROWS = 16, THIN_K_COLS = THIN_THREADS;   // THIN_K tile
constexpr int THIN_SMEM_FLOATS = 16 * 17 * 16;                // shared scratch: THIN_R reduction / THIN_N B tile / THIN_K A tile
constexpr int THIN_R_W = 16, THIN_R_KG = THIN_THREADS / THIN_R_W;   // THIN_R: 16 wide indices x 16 k-groups

// number of CTAs ("tiles") a task needs, by kind; tiles_n is what the kernel divides the local tile index by
static inline int thin_tiles(const GemmTask& t, int kind, int* tiles_n) {
  if (kind == THIN_K) {
    *tiles_n = (t.N + THIN_K_COLS - 1) / THIN_K_COLS;
    return ((t.M + THIN_K_ROWS - 1) / THIN_K_ROWS) * *tiles_n;
  }
  *tiles_n = 1;
  if (kind == THIN_N) return (t.M + t.klen - 1) / t.klen;   // klen = rows per CTA (set by the host)
  if (kind == THIN_R_WIDE_M) return (t.M + THIN_R_W - 1) / THIN_R_W;
  return (t.N + THIN_R_W - 1) / THIN_R_W;
}

__device__ __forceinline__ float thin_rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float ld_a(const GemmTask& t, int i, int k) {
  return t.a_kc ? t.A[(size_t)i * t.lda + k] : t.A[(size_t)k * t.lda + i];
}
__device__ __forceinline__ float ld_b(const GemmTask& t, int j, int k) {
  return t.b_kc ? t.B[(size_t)j * t.ldb + k] : t.B[(size_t)k * t.ldb + j];
}

// ---------------------------------------------------------------- K <= 16
__device__ __forceinline__ void thin_k_body(const GemmTask& t, int lt, float* smem) {
  const int tid = threadIdx.x;
  const int m0 = (lt / t.tiles_n) * THIN_K_ROWS, n0 = (lt % t.tiles_n) * THIN_K_COLS;
  float(*As)[17] = reinterpret_cast<float(*)[17]>(smem);
  for (int e = tid; e < THIN_K_ROWS * 16; e += THIN_THREADS) {
    int r, k;
    if (t.a_kc) { r = e / 16; k = e % 16; } else { k = e / THIN_K_ROWS; r = e % THIN_K_ROWS; }
    As[r][k] = (m0 + r < t.M && k < t.K) ? ld_a(t, m0 + r, k) : 0.f;
  }
  const int j = n0 + tid;
  float b[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) b[k] = (j < t.N && k < t.K) ? ld_b(t, j, k) : 0.f;
  const float bias = (t.bias && j < t.N) ? t.bias[j] : 0.f;
  const Epi ep = make_epi(t);
  const int rows = min(THIN_K_ROWS, t.M - m0);
  __syncthreads();
  if (j >= t.N) return;
  // the common first layer: ReLU(x W^T + b) and nothing else.  Full 16-row tiles are unrolled so the row part of
  // every address is an immediate; the generic epilogue costs more instructions than the 16 FMAs it follows
  // (the 21 MB stacked first layer of the 8 target Q networks was issue-bound on it).
  const bool plain = ep.act == ACT_RELU && !ep.aux && !ep.resid && !ep.clamp && !ep.dact && ep.scale == 1.f &&
                     rows == THIN_K_ROWS;
  if (plain && t.pk_hi && t.c_dead) {
    const int grp = j / t.pk_gcols, col = j - grp * t.pk_gcols;
    // m0 is a multiple of 16: rows m0..m0+15 lie in one 64-row block, 8-row groups g0 and g0+1
    const size_t base = ((size_t)(m0 >> 6) * t.pk_ks + (col >> 5)) * 2048 + (size_t)(((m0 & 63) >> 3) * 256) + (col & 3);
    float* __restrict__ phi = t.pk_hi + (size_t)grp * t.pk_gstride + base;
    float* __restrict__ plo = t.pk_lo + (size_t)grp * t.pk_gstride + base;
    const int ck = (col & 31) >> 2;
#pragma unroll
    for (int r = 0; r < THIN_K_ROWS; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(As[r][k], b[k], acc);
      const float v = fmaxf(acc + bias, 0.f);   // (bias last, like the generic epilogue: keeps the ReLU kinks where they were)
      const int off = (r >> 3) * 256 + (r & 7) * 32 + ((ck ^ (r & 7)) << 2);
      const float hi = thin_rna_tf32(v);
      phi[off] = hi;
      plo[off] = thin_rna_tf32(v - hi);
    }
    return;
  }
  if (plain && !t.pk_hi) {
    float* __restrict__ crow = ep.C + (size_t)m0 * ep.ldc + j;
#pragma unroll
    for (int r = 0; r < THIN_K_ROWS; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(As[r][k], b[k], acc);
      crow[(size_t)r * ep.ldc] = fmaxf(acc + bias, 0.f);
    }
    return;
  }
  if (t.pk_hi) {   // packed tf32 hi/lo images for a tcgen05 consumer (see GemmTask), optionally instead of C
    const int grp = j / t.pk_gcols, col = j - grp * t.pk_gcols, ksr = t.pk_ks;
    float* __restrict__ phi = t.pk_hi + (size_t)grp * t.pk_gstride;
    float* __restrict__ plo = t.pk_lo + (size_t)grp * t.pk_gstride;
    const bool dead = t.c_dead != 0;
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(As[r][k], b[k], acc);
      const float v = epi_value<false>(ep, bias, m0 + r, j, acc);
      if (!dead) ep.C[(size_t)(m0 + r) * ep.ldc + j] = v;
      const size_t off = pk_offset(m0 + r, col, ksr);
      const float hi = thin_rna_tf32(v);
      phi[off] = hi;
      plo[off] = thin_rna_tf32(v - hi);
    }
    return;
  }
  if (rows == THIN_K_ROWS) {
#pragma unroll 4
    for (int r = 0; r < THIN_K_ROWS; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(As[r][k], b[k], acc);
      epi_store<false>(ep, bias, m0 + r, j, acc);
    }
  } else {
    for (int r = 0; r < rows; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc = fmaf(As[r][k], b[k], acc);
      epi_store<false>(ep, bias, m0 + r, j, acc);
    }
  }
}

// ---------------------------------------------------------------- N <= 16, rows of A contiguous
// One warp per row, lanes stride k; B (N x K, a few KB) is read through L1, where every warp after the first
// finds it.  (Measured: holding a whole A row in registers, or staging B through shared memory, is slower --
// the 4-deep unrolled loop below already keeps enough loads in flight and stays under 64 registers.)
// 5..16 outputs per row with k-contiguous, 16-byte aligned operands (the VAE's mean / log-std heads): lane =
// (k-quarter, output n) -- every lane owns ONE output and a slice of k read as float4, so nothing but one
// accumulator lives in registers (the lanes-stride-k form needs N accumulators and N loads per k in flight and ran
// 3-5x slower for N = 8 at the 64-register cap); the k slices are combined with one or two shuffles.
template <int NL>   // lanes per row across n: 8 or 16
__device__ __forceinline__ void thin_n_split_body(const GemmTask& t, int lt) {
  constexpr int KQ = 32 / NL;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = lane % NL, kq = lane / NL;
  const int N = t.N, K = t.K, M = t.M, lda = t.lda, ldb = t.ldb;
  const float* __restrict__ A = t.A;
  const float* __restrict__ B = t.B;
  const int rows_per_tile = t.klen;
  const int row1 = min(M, (lt + 1) * rows_per_tile);
  const Epi ep = make_epi(t);
  const float bias = (t.bias && n < N) ? t.bias[n] : 0.f;
  const int k4tot = K / 4, per = (k4tot + KQ - 1) / KQ;
  const int q0 = kq * per, q1 = min(k4tot, q0 + per);
  const float4* __restrict__ b4 = reinterpret_cast<const float4*>(B + (size_t)min(n, N - 1) * ldb);
  for (int row = lt * rows_per_tile + warp; row < row1; row += THIN_THREADS / 32) {
    const float4* __restrict__ a4 = reinterpret_cast<const float4*>(A + (size_t)row * lda);
    float acc = 0.f;
#pragma unroll 5
    for (int q = q0; q < q1; ++q) {
      const float4 a = a4[q], b = b4[q];
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
#pragma unroll
    for (int o = NL; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (kq == 0 && n < N) epi_store<false>(ep, bias, row, n, acc);
  }
}

template <int NMAX>
__device__ __forceinline__ void thin_n_body(const GemmTask& t, int lt) {
  if constexpr (NMAX >= 8) {
    if (t.b_kc && t.a_vec && t.b_vec) {   // (set by the host: bases, leading dimensions and K 4-float aligned)
      thin_n_split_body<NMAX>(t, lt);
      return;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = t.N, K = t.K, M = t.M, lda = t.lda, ldb = t.ldb;
  const bool bkc = t.b_kc != 0;
  const float* __restrict__ A = t.A;
  const float* __restrict__ B = t.B;
  const int rows_per_tile = t.klen;            // host: 8 (one row per warp), 16 for very long M
  const int row1 = min(M, (lt + 1) * rows_per_tile);
  const Epi ep = make_epi(t);
  const float bias = (t.bias && lane < N) ? t.bias[lane] : 0.f;   // lane n finishes output column n
  for (int row = lt * rows_per_tile + warp; row < row1; row += THIN_THREADS / 32) {
    const float* __restrict__ a = A + (size_t)row * lda;
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    if (bkc) {
#pragma unroll 4
      for (int k = lane; k < K; k += 32) {
        const float av = a[k];
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
          if (n < N) acc[n] = fmaf(av, B[(size_t)n * ldb + k], acc[n]);
      }
    } else {
#pragma unroll 4
      for (int k = lane; k < K; k += 32) {
        const float av = a[k];
        const float* __restrict__ bp = B + (size_t)k * ldb;
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
          if (n < N) acc[n] = fmaf(av, bp[n], acc[n]);
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      float v = acc[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == n) mine = v;
    }
    if (lane < N) epi_store<false>(ep, bias, row, lane, mine);
  }
}

// ---------------------------------------------------------------- batch reduction, thin extent <= 16
// wide operand Xw[k*ldw + w], thin operand Xt[k*ldt + q];  out(w, q) = sum_k Xw[k,w] Xt[k,q]
template <int TMAX, bool WIDE_M>
__device__ __forceinline__ void thin_r_body(const GemmTask& t, int lt, float* smem) {
  const int tid = threadIdx.x, wl = tid % THIN_R_W, kg = tid / THIN_R_W;
  const int W = WIDE_M ? t.M : t.N, T = WIDE_M ? t.N : t.M, K = t.K;
  const float* __restrict__ Xw = WIDE_M ? t.A : t.B;
  const float* __restrict__ Xt = WIDE_M ? t.B : t.A;
  const int ldw = WIDE_M ? t.lda : t.ldb, ldt = WIDE_M ? t.ldb : t.lda;
  const int w = lt * THIN_R_W + wl;
  const bool ok = w < W;
  float acc[TMAX], cs[WIDE_M ? 1 : TMAX];
#pragma unroll
  for (int q = 0; q < TMAX; ++q) acc[q] = 0.f;
#pragma unroll
  for (int q = 0; q < (WIDE_M ? 1 : TMAX); ++q) cs[q] = 0.f;
  const bool want_cs = t.colsum != nullptr;
  const Epi ep = make_epi(t);
  constexpr int UNR = TMAX <= 2 ? 8 : (TMAX <= 4 ? 4 : (TMAX <= 8 ? 2 : 1));   // loads in flight per thread
#pragma unroll UNR
  for (int k = kg; k < K; k += THIN_R_KG) {
    const float xw = ok ? Xw[(size_t)k * ldw + w] : 0.f;
    const float* __restrict__ tp = Xt + (size_t)k * ldt;
    if (WIDE_M) cs[0] += xw;
#pragma unroll
    for (int q = 0; q < TMAX; ++q)
      if (q < T) {
        const float xt = tp[q];
        acc[q] = fmaf(xw, xt, acc[q]);
        if (!WIDE_M) cs[q] += xt;
      }
  }
  // fixed-order reduction over the k-groups: red[kg][q][wl]
  float* red = smem;
  constexpr int QS = TMAX + 1;   // slot TMAX carries the wide colsum
#pragma unroll
  for (int q = 0; q < TMAX; ++q) red[(kg * QS + q) * THIN_R_W + wl] = acc[q];
  red[(kg * QS + TMAX) * THIN_R_W + wl] = WIDE_M ? cs[0] : 0.f;
  __syncthreads();
  // thread (wl, q) sums the 16 partials of one output; slot q == TMAX is the wide colsum
  for (int q = kg; q <= TMAX; q += THIN_R_KG) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < THIN_R_KG; ++g) v += red[(g * QS + q) * THIN_R_W + wl];
    if (!ok) continue;
    if (q < TMAX) {
      if (q < T) {   // (weight gradients: no bias)
        if (WIDE_M) epi_store<false>(ep, 0.f, w, q, v);
        else epi_store<false>(ep, 0.f, q, w, v);
      }
    } else if (WIDE_M && want_cs) {
      t.colsum[w] = v;
    }
  }
  if (!WIDE_M && want_cs && lt == 0) {   // thin colsum: sum_k Xt[k, q]; every wl column holds the same partials
    __syncthreads();
    if (wl == 0) {
#pragma unroll
      for (int q = 0; q < TMAX; ++q) red[kg * TMAX + q] = cs[q];
    }
    __syncthreads();
    if (tid < T) {
      float v = 0.f;
      for (int g = 0; g < THIN_R_KG; ++g) v += red[g * TMAX + tid];
      t.colsum[tid] = v;
    }
  }
}

template <int TMAX>
__device__ __forceinline__ void thin_dispatch(const GemmTask& t, int lt, float* smem) {
  if (t.thin == THIN_N) thin_n_body<TMAX>(t, lt);
  else if (t.thin == THIN_R_WIDE_M) thin_r_body<TMAX, true>(t, lt, smem);
  else thin_r_body<TMAX, false>(t, lt, smem);
}

static __global__ void __launch_bounds__(THIN_THREADS, 4) k_gemm_thin(const __grid_constant__ TaskPack P, int ntasks) {
  __shared__ float smem[THIN_SMEM_FLOATS];
  const GemmTask& t = P.t[find_task(P, ntasks, blockIdx.x)];
  const int lt = blockIdx.x - t.tile0;
  if (t.thin == THIN_K) { thin_k_body(t, lt, smem); return; }
  const int T = t.thin == THIN_R_WIDE_N ? t.M : t.N;
  if (T <= 1) thin_dispatch<1>(t, lt, smem);
  else if (T <= 2) thin_dispatch<2>(t, lt, smem);
  else if (T <= 4) thin_dispatch<4>(t, lt, smem);
  else if (T <= 8) thin_dispatch<8>(t, lt, smem);
  else thin_dispatch<16>(t, lt, smem);
}

}  // namespace osrl
