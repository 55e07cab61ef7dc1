// Data-parallel gradient exchange over NVLink peer memory, replacing ncclAllReduce + k_adam
// on the step's critical path.  One process per GPU; every rank maps every peer's gradient section and flag block
// (cudaIpc, exchanged once in osrl_comm_init).  Per optimiser group and step:
//
//   k_dp_adam:  block 0 tells every peer "my gradients of this group are final" (release store of the round number
//               into the peer's flag block), every block waits until all peers said so, then each thread sums the W
//               gradient copies of its elements IN RANK ORDER (identical bits on every rank, so the replicas never
//               drift) straight out of peer memory and applies Adam (+ Polyak) to the local replica; the last block
//               tells every peer "I am done reading you".
//   k_prologue: before a rank overwrites a gradient range in the next step it waits for those "done" marks.
//   k_dp_scalar: the cross-batch scalars (PID error mean net.py:380, BEAR's mean MMD bearl.py:261) as a one-warp
//               all-gather + ordered sum.
//
// Everything is a plain kernel node: it replays inside the step graphs, round numbers live in device memory.  Waits are
// bounded (DP_SPIN_NS); a timeout raises DevState-independent `timeout` in the local flag block, checked by the host.
#pragma once
#include <cstdint>

namespace osrl {

constexpr int DP_MAX_WORLD = 8;
constexpr int DP_MAX_SLOT = 32;
constexpr int DP_SCAL_N = 8;
constexpr unsigned long long DP_SPIN_NS = 4000000000ull;   // 4 s

struct DpFlags {   // one per rank, in its own cudaMalloc so that peers can map it
  unsigned ready[DP_MAX_SLOT][DP_MAX_WORLD];              // written by peer r: round in which its gradients became final
  unsigned done[DP_MAX_SLOT][DP_MAX_WORLD];               // written by peer r: round whose reads of MY gradients it finished
  unsigned scal_flag[2][DP_MAX_SLOT][DP_MAX_WORLD];
  float scal[2][DP_MAX_SLOT][DP_MAX_WORLD][DP_SCAL_N];
  // local only
  unsigned round[DP_MAX_SLOT];                            // completed rounds of each slot on this rank
  unsigned arrive[DP_MAX_SLOT];                           // blocks that finished the current round
  unsigned scal_round[DP_MAX_SLOT];
  unsigned timeout;
  unsigned magic;                                          // set-up check: 0xD9000000 | rank
};
struct DpPeers {
  const float* G[DP_MAX_WORLD];   // every rank's gradient section (own rank: local pointer)
  DpFlags* flags[DP_MAX_WORLD];
  int world, rank;
};

__device__ __forceinline__ void dp_store_release(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned dp_load_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dp_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want (rounds only grow); false on timeout
__device__ __forceinline__ bool dp_wait(const unsigned* p, unsigned want, unsigned* timeout_flag) {
  if ((int)(dp_load_acquire(p) - want) >= 0) return true;
  const unsigned long long t0 = dp_now();
  while ((int)(dp_load_acquire(p) - want) < 0) {
    __nanosleep(64);
    if (dp_now() - t0 > DP_SPIN_NS) { *timeout_flag = 1u; return false; }
  }
  return true;
}
__device__ __forceinline__ float4 dp_ld_peer(const float* p) {   // L2-only: peer lines must not linger in L1
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// gradient all-reduce (ordered sum out of peer memory) fused with Adam (+ Polyak): same arithmetic as k_adam on the
// summed gradient.  `off` = float offset of the range inside the gradient section.  grid <= 148 (see engine.cu).
static __global__ void __launch_bounds__(256) k_dp_adam(const __grid_constant__ DpPeers pr, int slot, int64_t off,
                                                        float* __restrict__ P, float* __restrict__ Mm,
                                                        float* __restrict__ Vv, float* __restrict__ T, int64_t n4,
                                                        const DevState* ds, int group, float beta1, float beta2, float w1,
                                                        float w2, float eps, float weight_decay, float tau, int polyak) {
  DpFlags* mine = pr.flags[pr.rank];
  const int W = pr.world;
  const unsigned round = mine->round[slot] + 1u;   // stable during the launch: only the last block advances it
  if (blockIdx.x == 0 && (int)threadIdx.x < W && (int)threadIdx.x != pr.rank)
    dp_store_release(&pr.flags[threadIdx.x]->ready[slot][pr.rank], round);
  if ((int)threadIdx.x < W && (int)threadIdx.x != pr.rank) dp_wait(&mine->ready[slot][threadIdx.x], round, &mine->timeout);
  __syncthreads();
  const float step_size = ds->adam_step_size[group];
  const float bc2s = ds->adam_bc2_sqrt[group];
  const float lr = ds->adam_lr[group];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 gr[DP_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
      if (r < W) gr[r] = dp_ld_peer(pr.G[r] + off + 4 * i);
    float gg[4] = {gr[0].x, gr[0].y, gr[0].z, gr[0].w};
#pragma unroll
    for (int r = 1; r < DP_MAX_WORLD; ++r)
      if (r < W) { gg[0] += gr[r].x; gg[1] += gr[r].y; gg[2] += gr[r].z; gg[3] += gr[r].w; }
    float4 p = reinterpret_cast<float4*>(P)[i];
    float4 m = reinterpret_cast<float4*>(Mm)[i];
    float4 v = reinterpret_cast<float4*>(Vv)[i];
    float pp[4] = {p.x, p.y, p.z, p.w};
    float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k];
      if (weight_decay != 0.f) pp[k] = pp[k] * (1.f - lr * weight_decay);
      mm[k] = fmaf(w1, gk - mm[k], mm[k]);
      vv[k] = vv[k] * beta2 + (w2 * gk) * gk;
      const float denom = sqrtf(vv[k]) / bc2s + eps;
      pp[k] = pp[k] + (-step_size * mm[k]) / denom;
    }
    reinterpret_cast<float4*>(P)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(Mm)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(Vv)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (polyak) {
      float4 t = reinterpret_cast<float4*>(T)[i];
      t.x = tau * pp[0] + (1.f - tau) * t.x;
      t.y = tau * pp[1] + (1.f - tau) * t.y;
      t.z = tau * pp[2] + (1.f - tau) * t.z;
      t.w = tau * pp[3] + (1.f - tau) * t.w;
      reinterpret_cast<float4*>(T)[i] = t;
    }
  }
  // last block: this rank has read everything -> release the peers' gradient ranges, advance the round
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) last = (atomicAdd(&mine->arrive[slot], 1u) == gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) { mine->arrive[slot] = 0u; mine->round[slot] = round; }
    if ((int)threadIdx.x < W && (int)threadIdx.x != pr.rank) dp_store_release(&pr.flags[threadIdx.x]->done[slot][pr.rank], round);
  }
}

// in-place sum of n <= DP_SCAL_N floats over the ranks, in rank order (one warp)
static __global__ void k_dp_scalar(const __grid_constant__ DpPeers pr, int slot, float* v, int n) {
  DpFlags* mine = pr.flags[pr.rank];
  const int W = pr.world, t = threadIdx.x;
  const unsigned round = mine->scal_round[slot] + 1u;
  const int par = (int)(round & 1u);
  if (t < W) {
    DpFlags* dst = pr.flags[t];
    for (int j = 0; j < n; ++j) dst->scal[par][slot][pr.rank][j] = v[j];
    __threadfence_system();
    dp_store_release(&dst->scal_flag[par][slot][pr.rank], round);
    dp_wait(&mine->scal_flag[par][slot][t], round, &mine->timeout);
  }
  __syncwarp();
  if (t < n) {
    float s = 0.f;
    for (int r = 0; r < W; ++r) {
      float x;
      asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(x) : "l"(&mine->scal[par][slot][r][t]));
      s = r == 0 ? x : s + x;
    }
    v[t] = s;
  }
  if (t == 0) mine->scal_round[slot] = round;
}

// wait until every peer has finished reading my gradients of the slots in `slot_mask` (called from k_prologue)
__device__ __forceinline__ void dp_wait_done(const DpPeers& pr, unsigned slot_mask) {
  DpFlags* mine = pr.flags[pr.rank];
  for (int idx = threadIdx.x; idx < DP_MAX_SLOT * pr.world; idx += blockDim.x) {
    const int slot = idx / pr.world, r = idx % pr.world;
    if (!((slot_mask >> slot) & 1u) || r == pr.rank) continue;
    dp_wait(&mine->done[slot][r], mine->round[slot], &mine->timeout);
  }
}

// before a rank frees its gradient section: every peer must have finished its last read of it
static __global__ void k_dp_drain(const DpPeers* pr) { dp_wait_done(*pr, 0xffffffffu); }

}  // namespace osrl
