// Tensor-core variant of the multi-task GEMM: fp32-accurate 3xTF32 on mma.sync (sm_100a).
//
// Same task list, staging ring (cp.async, native layouts) and fused epilogue as gemm.cuh, but the
// inner product runs on the tensor cores: every fp32 operand x is split in registers into
//   x_hi = tf32(x),  x_lo = tf32(x - x_hi)
// and  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (three m16n8k8 TF32 MMAs, fp32 accumulate), which
// keeps ~21 mantissa bits per operand -- the parity mode (1e-5 vs the fp32 reference) holds.
// Shared-memory paddings are chosen so the fragment reads (thread (g,t) = (lane/4, lane%4)) are
// bank-conflict-free in both layouts:
//   k-contiguous  [rows][BK+4] : bank = (g*(BK+4) + t) mod 32  -> 32 distinct for BK in {16,32}
//   mn-contiguous [BK][rows+8] : bank = (t*(rows+8) + g) mod 32 = 8t+g -> 32 distinct
#pragma once
#include "gemm.cuh"

namespace osrl {

// KG > 1: "k-groups" -- KG copies of the WM x WN warp grid share a tile, group g taking the k8-steps g, g+KG, ...
// of every slab, and the partial tiles are summed through shared memory in a fixed order at the end.  The small
// layers of this workload (256 rows) are one 32x32 tile per SM with a 13-slab dependent chain: splitting the
// chain four ways cuts the kernel's critical path, which is all that matters there.
template <int BM, int BN, int BK, int WM_, int WN_, int NSTAGE, int KG_ = 1>
struct MmaCfg {
  static constexpr int WARPS_M = WM_, WARPS_N = WN_, KG = KG_;
  static constexpr int NT = WARPS_M * WARPS_N * 32 * KG;
  static constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;  // warp tile
  static constexpr int MT = WTM / 16, NTL = WTN / 8;            // mma tiles per warp
  static constexpr int A_KC = BM * (BK + 4), A_MC = BK * (BM + 8);
  static constexpr int B_KC = BN * (BK + 4), B_MC = BK * (BN + 8);
  static constexpr int A_STAGE = A_KC > A_MC ? A_KC : A_MC;
  static constexpr int B_STAGE = B_KC > B_MC ? B_KC : B_MC;
  static constexpr int SMEM_BYTES = NSTAGE * (A_STAGE + B_STAGE) * (int)sizeof(float);
  static_assert(WTM % 16 == 0 && WTN % 8 == 0 && BK % 8 == 0, "bad mma tiling");
  static_assert(BK % (8 * KG) == 0, "every k-group needs a k8-step in every slab");
  static_assert(KG == 1 || (KG * BM * (BN + 1) + KG * BM) * (int)sizeof(float) <= SMEM_BYTES, "reduction scratch");
};

// stage ROWS x BK elements of one operand (layout fixed at compile time, 16B/4B chosen per task)
template <int ROWS, int BK, int NT, bool KC>
__device__ __forceinline__ void stage_op(float* __restrict__ s, const float* __restrict__ G, int ld, bool vec, int r0,
                                         int k0, int R, int K, int tid) {
  if constexpr (KC) {  // G[r*ld + k] -> s[r*(BK+4) + k]
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
#pragma unroll 4
      for (int e = tid; e < EL; e += NT) {
        const int r = e / BK, k = e % BK;
        const bool ok = (r0 + r < R) && (k0 + k < K);
        cp_async4(s + r * (BK + 4) + k, ok ? G + (size_t)(r0 + r) * ld + k0 + k : G, ok);
      }
    }
  } else {  // G[k*ld + r] -> s[k*(ROWS+8) + r]
    if (vec) {
      constexpr int CH = BK * (ROWS / 4);
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        const int k = c / (ROWS / 4), rq = (c % (ROWS / 4)) * 4;
        const bool ok = (k0 + k < K) && (r0 + rq < R);
        cp_async16(s + k * (ROWS + 8) + rq, ok ? G + (size_t)(k0 + k) * ld + r0 + rq : G, ok);
      }
    } else {
      constexpr int EL = ROWS * BK;
#pragma unroll 4
      for (int e = tid; e < EL; e += NT) {
        const int k = e / ROWS, r = e % ROWS;
        const bool ok = (k0 + k < K) && (r0 + r < R);
        cp_async4(s + k * (ROWS + 8) + r, ok ? G + (size_t)(k0 + k) * ld + r0 + r : G, ok);
      }
    }
  }
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <class Cfg, int BM, int BN, int BK, int NSTAGE, bool AKC, bool BKC, bool FULL>
__device__ __forceinline__ void gemm_mma_body(const GemmTask& t, float* __restrict__ As, float* __restrict__ Bs,
                                              int m0, int n0, int kbeg) {
  constexpr int NT = Cfg::NT, MT = Cfg::MT, NTL = Cfg::NTL;
  constexpr int KG = Cfg::KG;
  const int tid = threadIdx.x, lane = tid & 31;
  const int kgrp = (tid >> 5) / (Cfg::WARPS_M * Cfg::WARPS_N), warp = (tid >> 5) % (Cfg::WARPS_M * Cfg::WARPS_N);
  const int g = lane >> 2, tq = lane & 3;
  const int wm = (warp / Cfg::WARPS_N) * Cfg::WTM;
  const int wn = (warp % Cfg::WARPS_N) * Cfg::WTN;
  const int M = t.M, N = t.N, K = FULL ? min(t.K, kbeg + t.klen) : t.K;   // K = end of this CTA's k range
  const float* __restrict__ A = t.A;
  const float* __restrict__ B = t.B;
  const int lda = t.lda, ldb = t.ldb;
  const bool avec = t.a_vec != 0, bvec = t.b_vec != 0;

  float acc[MT][NTL][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  float rs[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i) rs[i][0] = rs[i][1] = 0.f;
  const bool want_colsum = (t.colsum != nullptr) && (n0 == 0) && (wn == 0);

  const int nk = (K - kbeg + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < NSTAGE - 1; ++s) {
    if (s < nk) {
      stage_op<BM, BK, NT, AKC>(As + s * Cfg::A_STAGE, A, lda, avec, m0, kbeg + s * BK, M, K, tid);
      stage_op<BN, BK, NT, BKC>(Bs + s * Cfg::B_STAGE, B, ldb, bvec, n0, kbeg + s * BK, N, K, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    {
      const int nx = kt + NSTAGE - 1;
      if (nx < nk) {
        const int sb = nx % NSTAGE;
        stage_op<BM, BK, NT, AKC>(As + sb * Cfg::A_STAGE, A, lda, avec, m0, kbeg + nx * BK, M, K, tid);
        stage_op<BN, BK, NT, BKC>(Bs + sb * Cfg::B_STAGE, B, ldb, bvec, n0, kbeg + nx * BK, N, K, tid);
      }
      cp_async_commit();
    }
    const float* __restrict__ as = As + (kt % NSTAGE) * Cfg::A_STAGE;
    const float* __restrict__ bs = Bs + (kt % NSTAGE) * Cfg::B_STAGE;
    // per-slab tensor-core partials: the MMA unit accumulates with truncation, so only BK/8 MMAs are
    // chained per partial and the slabs are added with round-to-nearest FADDs (keeps the bias < 1e-6)
    float part[MT][NTL][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) part[i][j][q] = 0.f;
#pragma unroll
    for (int kk = 8 * kgrp; kk < BK; kk += 8 * KG) {
      uint32_t ah[MT][4], al[MT][4], bh[NTL][2], bl[NTL][2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int r = wm + i * 16 + g;
        float v[4];
        if constexpr (AKC) {
          v[0] = as[r * (BK + 4) + kk + tq];
          v[1] = as[(r + 8) * (BK + 4) + kk + tq];
          v[2] = as[r * (BK + 4) + kk + tq + 4];
          v[3] = as[(r + 8) * (BK + 4) + kk + tq + 4];
        } else {
          v[0] = as[(kk + tq) * (BM + 8) + r];
          v[1] = as[(kk + tq) * (BM + 8) + r + 8];
          v[2] = as[(kk + tq + 4) * (BM + 8) + r];
          v[3] = as[(kk + tq + 4) * (BM + 8) + r + 8];
        }
        if (want_colsum) { rs[i][0] += v[0] + v[2]; rs[i][1] += v[1] + v[3]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) split_tf32(v[q], ah[i][q], al[i][q]);
      }
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        const int c = wn + j * 8 + g;
        float v[2];
        if constexpr (BKC) {
          v[0] = bs[c * (BK + 4) + kk + tq];
          v[1] = bs[c * (BK + 4) + kk + tq + 4];
        } else {
          v[0] = bs[(kk + tq) * (BN + 8) + c];
          v[1] = bs[(kk + tq + 4) * (BN + 8) + c];
        }
        split_tf32(v[0], bh[j][0], bl[j][0]);
        split_tf32(v[1], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
          mma_tf32(part[i][j], al[i], bh[j]);   // small terms first
          mma_tf32(part[i][j], ah[i], bl[j]);
          mma_tf32(part[i][j], ah[i], bh[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] += part[i][j][q];
  }
  cp_async_wait<0>();

  if constexpr (KG > 1) {
    // ---- sum the k-groups' partial tiles through shared memory (the staging ring is idle now), fixed order
    __syncthreads();
    float* red = As;                             // [KG][BM][BN+1]
    float* redcs = As + KG * BM * (BN + 1);      // [KG][BM]
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NTL; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = wm + i * 16 + g + (q >> 1) * 8, c = wn + j * 8 + 2 * tq + (q & 1);
          red[(kgrp * BM + r) * (BN + 1) + c] = acc[i][j][q];
        }
    if (want_colsum) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v = rs[i][h];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (tq == 0) redcs[kgrp * BM + wm + i * 16 + g + h * 8] = v;
        }
    }
    __syncthreads();
    if (t.colsum != nullptr && n0 == 0 && tid < BM && m0 + tid < M) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < KG; ++k) v += redcs[k * BM + tid];
      if (FULL && t.ksplit > 1) atomicAdd(&t.colsum[m0 + tid], v);
      else t.colsum[m0 + tid] = v;
    }
    const Epi ep = make_epi(t);
    static_assert(NT % BN == 0, "a thread keeps its column across the reduction loop");
    const int c = tid % BN, gj = n0 + c;         // consecutive threads -> consecutive columns: coalesced epilogue
    const float bias = epi_bias(ep, gj, N);
    if (gj < N) {
      for (int r = tid / BN; r < BM; r += NT / BN) {
        const int gi = m0 + r;
        if (gi >= M) break;
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < KG; ++k) v += red[(k * BM + r) * (BN + 1) + c];
        epi_store<FULL>(ep, bias, gi, gj, v);
      }
    }
    return;
  }
  // ---- bias gradient: rows g / g+8 of each m-tile; the 4 lanes of a quad hold different k
  if (want_colsum) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = rs[i][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int gi = m0 + wm + i * 16 + g + h * 8;
        if (tq == 0 && gi < M) {
          if (FULL && t.ksplit > 1) atomicAdd(&t.colsum[gi], v);
          else t.colsum[gi] = v;
        }
      }
  }
  // ---- fused epilogue: c0,c1 -> (row g, cols 2t,2t+1); c2,c3 -> (row g+8, same cols)
  const Epi ep = make_epi(t);
  float bj[NTL][2];
#pragma unroll
  for (int j = 0; j < NTL; ++j) {
    bj[j][0] = epi_bias(ep, n0 + wn + j * 8 + 2 * tq, N);
    bj[j][1] = epi_bias(ep, n0 + wn + j * 8 + 2 * tq + 1, N);
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gi = m0 + wm + i * 16 + g + (q >> 1) * 8;
        const int gj = n0 + wn + j * 8 + 2 * tq + (q & 1);
        if (gi >= M || gj >= N) continue;
        epi_store<FULL>(ep, bj[j][q & 1], gi, gj, acc[i][j][q]);
      }
}

template <int BM, int BN, int BK, int WM_, int WN_, int NSTAGE, int KG, bool FULL>
__global__ void __launch_bounds__(WM_* WN_ * 32 * KG, (WM_ * WN_ * 32 * KG <= 256 ? 2 : 1))   // <= 128 registers: 2 CTAs/SM
k_gemm_mma(const __grid_constant__ TaskPack P, int ntasks) {
  using Cfg = MmaCfg<BM, BN, BK, WM_, WN_, NSTAGE, KG>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + NSTAGE * Cfg::A_STAGE;
  const GemmTask& t = P.t[find_task(P, ntasks, blockIdx.x)];
  int lt = blockIdx.x - t.tile0, kbeg = 0;
  if constexpr (FULL) { kbeg = (lt / t.tiles_mn) * t.klen; lt %= t.tiles_mn; }
  const int m0 = (lt / t.tiles_n) * BM;
  const int n0 = (lt % t.tiles_n) * BN;
  // CTA-uniform dispatch on the operand layouts: each body is fully specialised (no layout branches
  // in the k-loop).  forward: A,B k-contiguous; dgrad: A k-contiguous, B n-contiguous; wgrad: both mn.
  if (t.a_kc && t.b_kc) gemm_mma_body<Cfg, BM, BN, BK, NSTAGE, true, true, FULL>(t, As, Bs, m0, n0, kbeg);
  else if (t.a_kc && !t.b_kc) gemm_mma_body<Cfg, BM, BN, BK, NSTAGE, true, false, FULL>(t, As, Bs, m0, n0, kbeg);
  else if (!t.a_kc && !t.b_kc) gemm_mma_body<Cfg, BM, BN, BK, NSTAGE, false, false, FULL>(t, As, Bs, m0, n0, kbeg);
  else gemm_mma_body<Cfg, BM, BN, BK, NSTAGE, false, true, FULL>(t, As, Bs, m0, n0, kbeg);
}

}  // namespace osrl
