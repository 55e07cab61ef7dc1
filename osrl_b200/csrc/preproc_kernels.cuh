// One-time trajectory preprocessing on the device: what process_sequence_dataset
// (dataset.py:137-183) does with a Python loop over every transition -- split the flat DSRL arrays into episodes at
// terminals | timeouts, reward-to-go and cost-to-go per episode (discounted_cumsum with gamma = 1, dataset.py:19-27) --
// producing directly the packed trajectory buffer k_seq_gather reads.  Integer work is exact; the two running sums are
// evaluated per episode from its last transition backwards with one round-to-nearest fp32 add per step, the same
// operations in the same order as the reference's loop, so the result is bit-identical.
#pragma once
#include <cstdint>

namespace osrl {

struct EpisodeCount { long long n_traj, n_used; };

// offsets[e + 1] = index after the e-th transition with terminals | timeouts; n_used = offsets[n_traj]: a trailing
// episode without an end flag is dropped, as the reference's loop never appends it.  One CTA walks the array in
// 1024-wide chunks (block scan of the end flags, running total carried in a register): order-preserving, exact.
static __global__ void __launch_bounds__(1024) k_episode_offsets(const uint8_t* __restrict__ term,
                                                                 const uint8_t* __restrict__ tout, long long n,
                                                                 long long* __restrict__ offsets, EpisodeCount* out) {
  __shared__ int warp_tot[32];
  __shared__ int chunk_tot;
  long long running = 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) offsets[0] = 0;
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + tid;
    const int f = (i < n && ((term && term[i]) || (tout && tout[i]))) ? 1 : 0;
    int incl = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t += v;
      }
      warp_tot[lane] = t;   // inclusive totals of the warps
      if (lane == 31) chunk_tot = t;
    }
    __syncthreads();
    const int before = (w ? warp_tot[w - 1] : 0) + incl - f;   // end flags before transition i inside the chunk
    if (f) offsets[running + before + 1] = i + 1;
    running += chunk_tot;
    __syncthreads();
  }
  if (tid == 0) { out->n_traj = running; }
}
static __global__ void k_episode_finish(const long long* __restrict__ offsets, EpisodeCount* out) {
  out->n_used = offsets[out->n_traj];
}

// packed rows [obs | act | return-to-go | cost-to-go | cost | pad]; the two sums are filled by k_episode_suffix
static __global__ void k_seq_pack(const float* __restrict__ obs, const float* __restrict__ act, const float* __restrict__ cost,
                                  long long n_used, int o, int a, int stride, int cost_reverse, float* __restrict__ rows) {
  const long long total = n_used * stride;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / stride;
    const int c = (int)(e - i * stride);
    float v = 0.f;
    if (c < o) v = obs[i * o + c];
    else if (c < o + a) v = act[i * a + (c - o)];
    else if (c == o + a + 2) v = cost_reverse ? __fsub_rn(1.0f, cost[i]) : cost[i];   // dataset.py:165: 1.0 - cost (float32)
    rows[e] = v;
  }
}

// one thread per episode, backwards: acc = x[t] + acc (dataset.py:24-26 with gamma = 1).  The buffer keeps
// return * reward_scale / cost_return * cost_scale (dataset.py:762-763: float32 * weak python float); the unscaled
// first elements (what cost_sample / pf_sample / the augmentation read, dataset.py:452-458) go to first_*.
static __global__ void k_episode_suffix(const float* __restrict__ rew, const long long* __restrict__ offsets,
                                        long long n_traj, int o, int a, int stride, float reward_scale, float cost_scale,
                                        float* __restrict__ rows, float* __restrict__ first_ret,
                                        float* __restrict__ first_cret) {
  const long long ep = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (ep >= n_traj) return;
  const long long lo = offsets[ep], hi = offsets[ep + 1];
  float r = 0.f, c = 0.f;
  for (long long t = hi - 1; t >= lo; --t) {
    float* row = rows + t * stride + o + a;
    const float x = rew[t], y = row[2];
    r = (t == hi - 1) ? x : __fadd_rn(x, r);
    c = (t == hi - 1) ? y : __fadd_rn(y, c);
    row[0] = __fmul_rn(r, reward_scale);
    row[1] = __fmul_rn(c, cost_scale);
  }
  first_ret[ep] = r;
  first_cret[ep] = c;
}

}  // namespace osrl
