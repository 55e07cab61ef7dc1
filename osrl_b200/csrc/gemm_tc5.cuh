// tcgen05 / TMEM variant of the multi-task GEMM for the large forward layers (sm_100a).
//
// C[M,N] = A[M,K] * B[N,K]^T (both k-contiguous: activations x nn.Linear weights), fp32-accurate via
// 3xTF32: every operand tile is written to shared memory twice, as hi = x with the low 13 mantissa bits
// cleared and lo = x - hi (exact), and each k-step issues three tcgen05.mma.kind::tf32
// (hi*hi + lo*hi + hi*lo) into a 128 x BN fp32 accumulator that lives in TMEM.
//
//   * CTA tile 128 x 128, BK = 32 floats = one 128-byte swizzle row; 3-stage ring of
//     {A_hi, A_lo, B_hi, B_lo} (64 KB per stage) in the canonical K-major SWIZZLE_128B layout:
//         byte(r, c16) = (r/8)*1024 + (r%8)*128 + ((c16 ^ (r%8)) * 16)        (c16 = 16-byte chunk along k)
//     so coalesced global reads (8 lanes = one 128-byte row) become conflict-free shared stores.
//   * warps 0-7: producers (global -> registers -> hi/lo -> shared, fence.proxy.async, mbarrier arrive)
//     and, after the k-loop, the epilogue (tcgen05.ld 32x32b from TMEM -> fused epilogue -> global);
//     warp 8: one elected lane issues the MMAs and tcgen05.commit's the stage-empty / accumulator-full
//     mbarriers.  One CTA per SM (192 KB of shared memory).
#pragma once
#include "gemm.cuh"

namespace osrl {
namespace tc5 {

constexpr int BM = 128, BK = 32;
constexpr int PRODUCERS = 256, THREADS = PRODUCERS + 32;
constexpr int A_TILE_BYTES = BM * BK * 4;               // 16 KB
// Two shapes: 128x128 tiles, 3 stages, register-prefetched loads, one CTA per SM (192 KB); and 128x64 tiles,
// 2 stages, two CTAs per SM (2 x 97 KB) so one CTA's prologue / epilogue overlaps the other's k-loop and the
// tail wave is half as long.
template <int BN_, int NSTAGE_>
struct Shape {
  static constexpr int BN = BN_, NSTAGE = NSTAGE_;
  static constexpr int B_TILE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // A_hi, A_lo, B_hi, B_lo
  static constexpr int EPI_BYTES = BM * (BN + 4) * 4;                        // epilogue tile (reuses the ring)
  static constexpr int RING_BYTES = NSTAGE * STAGE_BYTES > EPI_BYTES ? NSTAGE * STAGE_BYTES : EPI_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + 1024;                       // + alignment slack
  // three accumulators of BN columns: hi*hi terms alternate between two of them (even / odd k8) and the small
  // lo*hi + hi*lo corrections get their own, so the tensor core's truncating fp32 accumulation is applied to a
  // third as many large-magnitude additions (measured: ~3e-6 -> ~1e-6 of max|C| at K = 400)
  static constexpr int TMEM_COLS = BN == 64 ? 256 : 512;                     // power of two >= 3 * BN
  static constexpr uint32_t IDESC =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO (ignored for swizzled K-major) = 1, SBO = 1024 B between 8-row groups, version 1, layout type 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// Shape::IDESC = cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10), K-major A and B,
// N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// split one 16-byte chunk into hi = tf32(x) and lo = tf32(x - hi), both round-to-nearest: kind::tf32 would
// otherwise TRUNCATE the low 13 mantissa bits of what it reads, which costs ~2 bits on the lo term
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = rna_tf32(v.x); hi.y = rna_tf32(v.y); hi.z = rna_tf32(v.z); hi.w = rna_tf32(v.w);
  lo.x = rna_tf32(v.x - hi.x); lo.y = rna_tf32(v.y - hi.y); lo.z = rna_tf32(v.z - hi.z); lo.w = rna_tf32(v.w - hi.w);
}

// APACK: the A operand arrives as ready-made hi/lo shared-memory images (GemmTask::a_hi / a_lo, written by the
// producer of the activations): one thread fetches a stage's four 8 KB blocks with cp.async.bulk (completion
// counted on the stage's full barrier) and the producer warps only load / split / store the 64-row B tile --
// a third of the shared-memory stores and ALU work of the generic path, which bounds this kernel.
template <int BN_, int NSTAGE_, int MINB, bool FULL, bool APACK = false>
__global__ void __launch_bounds__(THREADS, MINB) k_gemm_tc5(const __grid_constant__ TaskPack P, int ntasks) {
  using S = Shape<BN_, NSTAGE_>;
  constexpr int BN = S::BN, NSTAGE = S::NSTAGE, STAGE_BYTES = S::STAGE_BYTES, TMEM_COLS = S::TMEM_COLS;
  constexpr int A_T = A_TILE_BYTES, B_T = S::B_TILE_BYTES;
  constexpr bool PREFETCH = (MINB == 1) || APACK;   // APACK: only the small B tile passes through registers
  constexpr int BCH = BN * 8 / PRODUCERS;   // 16-byte chunks of the B tile per producer thread
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], acc_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1 KB alignment

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], PRODUCERS / 32 + (APACK ? 1 : 0));   // + the bulk-copy issuer's expect_tx arrival
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM allocation (one warp), address published through shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const GemmTask& t = P.t[find_task(P, ntasks, blockIdx.x)];
  const int lt = blockIdx.x - t.tile0;
  const int m0 = (lt / t.tiles_n) * BM, n0 = (lt % t.tiles_n) * BN;
  const int M = t.M, N = t.N, K = t.K;
  const int nk = (K + BK - 1) / BK;

  if (warp < PRODUCERS / 32) {
    // ------------------------------------------------ producers
    const float* __restrict__ A = t.A;
    const float* __restrict__ B = t.B;
    const int lda = t.lda, ldb = t.ldb;
    // register double buffer: the global loads of slab kt+1 are in flight while slab kt is split and stored,
    // so a k-step costs the shared-store time, not a global round trip
    auto load_slab = [&](int kt, float4 (&va)[4], float4 (&vb)[BCH]) {
      const int k0 = kt * BK;
      if constexpr (!APACK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // 1024 chunks of the A tile, 4 per thread; 8 lanes = one 128-byte row
          const int c = tid + PRODUCERS * i, r = c >> 3, ck = c & 7;
          const int gk = k0 + ck * 4;
          va[i] = (m0 + r < M && gk < K) ? *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + gk)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < BCH; ++i) {
        const int c = tid + PRODUCERS * i, r = c >> 3, ck = c & 7;
        const int gk = k0 + ck * 4;
        vb[i] = (n0 + r < N && gk < K) ? *reinterpret_cast<const float4*>(B + (size_t)(n0 + r) * ldb + gk)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_slab = [&](int kt, const float4 (&va)[4], const float4 (&vb)[BCH]) {
      const int s = kt % NSTAGE;
      if (kt >= NSTAGE) mbar_wait(&empty_bar[s], ((kt / NSTAGE) - 1) & 1);
      uint8_t* st = smem + (size_t)s * STAGE_BYTES;
      if constexpr (APACK) {
        if (tid == 0) {   // A_hi | A_lo of this stage = two 64-row blocks each, 8 KB apiece, already swizzled
          const uint32_t bar = smem_u32(&full_bar[s]);
          asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar),
                       "r"(2 * A_T)
                       : "memory");
          const int rb = m0 >> 6, ksr = t.pk_ks;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float* src = (h ? t.a_lo : t.a_hi);
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
              const float* g = src + ((size_t)(rb + b2) * ksr + kt) * 2048;
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      smem_u32(st + h * A_T + b2 * 8192)),
                  "l"(g), "r"(8192), "r"(bar)
                  : "memory");
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = tid + PRODUCERS * i, r = c >> 3, ck = c & 7;
          const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((ck ^ (r & 7)) << 4);
          float4 hi, lo;
          split4(va[i], hi, lo);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + A_T + off) = lo;
        }
      }
#pragma unroll
      for (int i = 0; i < BCH; ++i) {
        const int c = tid + PRODUCERS * i, r = c >> 3, ck = c & 7;
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((ck ^ (r & 7)) << 4);
        float4 hi, lo;
        split4(vb[i], hi, lo);
        *reinterpret_cast<float4*>(st + 2 * A_T + off) = hi;
        *reinterpret_cast<float4*>(st + 2 * A_T + B_T + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    };
    if constexpr (PREFETCH) {
      float4 a0[4], b0[BCH], a1[4], b1[BCH];
      load_slab(0, a0, b0);
      for (int kt = 0; kt < nk; kt += 2) {
        if (kt + 1 < nk) load_slab(kt + 1, a1, b1);
        store_slab(kt, a0, b0);
        if (kt + 1 < nk) {
          if (kt + 2 < nk) load_slab(kt + 2, a0, b0);
          store_slab(kt + 1, a1, b1);
        }
      }
    } else {   // two CTAs per SM: the co-resident CTA covers this one's load latency
      float4 a0[4], b0[BCH];
      for (int kt = 0; kt < nk; ++kt) {
        load_slab(kt, a0, b0);
        store_slab(kt, a0, b0);
      }
    }
    // ------------------------------------------------ epilogue: TMEM -> registers -> fused epilogue -> global
    mbar_wait(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, half = warp >> 2;          // TMEM lane quarter of this warp / column half
    // phase 1: accumulator rows -> shared tile [128][TP] (the stage ring is free: every MMA has completed).
    // A thread owns one accumulator ROW, so storing straight to global would touch 32 sectors per request
    // (ncu: 8x write amplification); the tile is re-read row-wise so that lanes hold consecutive columns.
    constexpr int TP = BN + 4;
    float* tile = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int cb = 0; cb < BN / 2; cb += 16) {
      const int col = half * (BN / 2) + cb;
      uint32_t v[3][16];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + col);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[a][0]), "=r"(v[a][1]), "=r"(v[a][2]), "=r"(v[a][3]), "=r"(v[a][4]), "=r"(v[a][5]), "=r"(v[a][6]),
              "=r"(v[a][7]), "=r"(v[a][8]), "=r"(v[a][9]), "=r"(v[a][10]), "=r"(v[a][11]), "=r"(v[a][12]),
              "=r"(v[a][13]), "=r"(v[a][14]), "=r"(v[a][15])
            : "r"(taddr));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = tile + (q * 32 + lane) * TP + col;
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        f[j] = (__uint_as_float(v[0][j]) + __uint_as_float(v[1][j])) + __uint_as_float(v[2][j]);
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(PRODUCERS) : "memory");   // the 8 epilogue warps only
    // phase 2: fused epilogue + coalesced stores (32 lanes = 32 consecutive columns of one row)
    const Epi ep = make_epi(t);
    float bj[BN / 32];
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) bj[j] = epi_bias(ep, n0 + j * 32 + lane, N);
    for (int r = warp; r < BM; r += PRODUCERS / 32) {
      const int gi = m0 + r;
      if (gi >= M) break;
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) {
        const int gj = n0 + j * 32 + lane;
        if (gj < N) epi_store<FULL>(ep, bj[j], gi, gj, tile[r * TP + j * 32 + lane]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (lane == 0) {
    // ------------------------------------------------ MMA issuer (one thread)
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % NSTAGE;
      mbar_wait(&full_bar[s], (kt / NSTAGE) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb = smem_u32(smem + (size_t)s * STAGE_BYTES);
      const uint64_t a_hi = make_desc(sb), a_lo = make_desc(sb + A_T);
      const uint64_t b_hi = make_desc(sb + 2 * A_T), b_lo = make_desc(sb + 2 * A_T + B_T);
#pragma unroll
      for (int k8 = 0; k8 < BK / 8; ++k8) {          // one MMA consumes K = 8 tf32 = 32 bytes of every row
        const uint64_t adv = (uint64_t)((k8 * 32) >> 4);
        mma_tf32_ss(tmem_base + 2 * BN, a_lo + adv, b_hi + adv, S::IDESC, (kt | k8) != 0);
        mma_tf32_ss(tmem_base + 2 * BN, a_hi + adv, b_lo + adv, S::IDESC, 1);
        mma_tf32_ss(tmem_base + (k8 & 1) * BN, a_hi + adv, b_hi + adv, S::IDESC, kt != 0 || k8 >= 2);
      }
      commit(&empty_bar[s]);            // arrives when the MMAs above have finished reading this stage
    }
    commit(&acc_bar);                   // accumulator complete -> epilogue
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace tc5
}  // namespace osrl
