// Elementwise / reduction / sampling kernels of the OSRL step (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace osrl {

// ------------------------------------------------------------------ device-resident scalar state
#define OSRL_MAX_GROUPS 8
struct DevState {
  unsigned long long step;            // completed steps (Philox counter word)
  unsigned long long vae_step;        // completed VAE updates: equals `step` between osrl_steps() calls; inside a
                                      // pipelined run the VAE update of step s+1 overlaps the rest of step s and
                                      // draws its batch / noise with this counter (engine.cu, pipelined graphs)
  int adam_t[OSRL_MAX_GROUPS];        // per-group Adam step count
  float adam_lr[OSRL_MAX_GROUPS];     // lr used for the current step (after schedule)
  float adam_step_size[OSRL_MAX_GROUPS];  // lr / (1 - beta1^t)
  float adam_bc2_sqrt[OSRL_MAX_GROUPS];   // sqrt(1 - beta2^t)
  float pid_e_old, pid_e_int;         // LagrangianPIDController state (net.py:373-374)
  float log_alpha;                    // CPQ / BEAR dual variable (cpq.py:93, bearl.py:112)
  int n_train_steps;                  // BEAR policy-update gate (bearl.py:249)
  double log_temperature;             // CDT: float64 leaf outside state_dict (cdt.py:144)
  double temp_m, temp_v;              // Adam state of log_temperature
  float scratch[16];
  // COptiDICE dual variables (coptidice.py:104-105): raw values (softplus applied in the step), their Adam state
  float cop_tau, cop_lmbda;
  float cop_m[2], cop_v[2];           // [0] tau, [1] lmbda
  int cop_t[2];
  float cop_lm_old;                   // softplus(lmbda) of the step's first phase, reused by the policy phase (:204-206)
};

// ------------------------------------------------------------------ Philox4x32-10
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c[0];
    const uint64_t p1 = (uint64_t)M1 * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

enum { STREAM_INDEX = 1, STREAM_NOISE0 = 16 };

// minibatch index draw: idx_i = floor(u32 * N / 2^32), u32 = philox(ctr = (i, step, STREAM_INDEX, rank))[0]
__host__ __device__ __forceinline__ int64_t draw_index(uint64_t seed, uint64_t step, uint32_t rank, uint32_t i,
                                                       int64_t n) {
  uint32_t c[4] = {i, (uint32_t)step, (uint32_t)STREAM_INDEX | ((uint32_t)(step >> 32) << 8), rank};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (int64_t)(((uint64_t)c[0] * (uint64_t)n) >> 32);
}

// ------------------------------------------------------------------ block reduction (deterministic)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// all threads get the block total; blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ float block_sum(float v, float* sh /*[33]*/) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.f;
    x = warp_sum(x);
    if (l == 0) sh[32] = x;
  }
  __syncthreads();
  return sh[32];
}

// ------------------------------------------------------------------ step prologue: counters + Adam scalars
struct AdamGroupCfg {
  double lr, beta1, beta2;
  int warmup;  // >0: lr * min((t)/warmup, 1) with t = step count after increment (LambdaLR, cdt.py:327-330)
};
}  // namespace osrl
#include "dp_peer.cuh"
namespace osrl {

static __global__ void k_prologue(DevState* ds, const AdamGroupCfg* g, int ngroups, unsigned mask, const DpPeers* dp = nullptr,
                                  unsigned dp_slot_mask = 0u) {
  if (blockIdx.x != 0) return;
  const int i = threadIdx.x;           // one thread per optimiser group (the double-precision pow()s are ~1.5 us each)
  if (i < ngroups && ((mask >> i) & 1u)) {   // pipelined graphs advance the VAE group and the others separately
    const int t = ds->adam_t[i] + 1;
    ds->adam_t[i] = t;
    double lr = (double)g[i].lr;
    if (g[i].warmup > 0) {
      double f = (double)t / (double)g[i].warmup;
      lr = lr * (f < 1.0 ? f : 1.0);
    }
    const double bc1 = 1.0 - pow(g[i].beta1, (double)t);
    const double bc2 = 1.0 - pow(g[i].beta2, (double)t);
    ds->adam_lr[i] = (float)lr;
    ds->adam_step_size[i] = (float)(lr / bc1);
    ds->adam_bc2_sqrt[i] = (float)sqrt(bc2);
  }
  // data parallel over peer memory: the step is about to overwrite gradient ranges the peers summed last step
  if (dp && dp->world > 1) dp_wait_done(*dp, dp_slot_mask);
}
// mode 0: a whole step finished (both counters); 1: everything but the VAE update; 2: the VAE update only
// Host-batch queue of osrl_steps_host: `ring` holds the call's k packed minibatches and `st_*` its per-step stat
// rows, all in pinned host memory mapped into the device address space; `side` / `main` = how many batches the VAE
// branch / the rest of the step have consumed in this call.
struct HostQueue {
  unsigned side, main, pad0, pad1;
  const float* ring;
  float* st_side;
  float* st_main;
  long long slot_floats;
};
constexpr int HQ_STAT_LD = 16;

// mode 0: a whole step ended; 1: the main branch of a pipelined step; 2: its VAE branch.  With a host queue the
// branch also posts its stat row to the host (zero-copy store) and advances its queue position.
static __global__ void k_epilogue(DevState* ds, int mode, HostQueue* q = nullptr, const float* stats = nullptr, int nstats = 0) {
  if (blockIdx.x != 0) return;
  if (q) {
    float* row = mode == 2 ? q->st_side + (size_t)q->side * HQ_STAT_LD : q->st_main + (size_t)q->main * HQ_STAT_LD;
    if ((int)threadIdx.x < nstats) row[threadIdx.x] = stats[threadIdx.x];
    __syncwarp();
  }
  if (threadIdx.x != 0) return;
  if (mode == 0) { ds->step += 1ull; ds->vae_step = ds->step; }
  else if (mode == 1) ds->step += 1ull;
  else ds->vae_step += 1ull;
  if (q) {
    if (mode != 2) q->main += 1u;
    if (mode != 1) q->side += 1u;
  }
}

// Minibatch `q->side` (which = 1) or `q->main` (which = 0) of the host queue -> the step's input buffers.  The slot is
// [obs B*o | next_obs B*o | act B*a | rew B | cost B | done B | is_init B]; the loads go over PCIe (mapped pinned
// memory), 22 KB for a CarCircle batch of 256.
static __global__ void k_unpack_batch(const HostQueue* q, int which, int B, int o, int a, float* obs, float* nobs, float* act,
                                      float* rew, float* cost, float* done, float* init) {
  const long long n = q->slot_floats;
  const float* src = q->ring + (size_t)(which ? q->side : q->main) * n;
  const long long e0 = (long long)B * o, e1 = 2 * e0, e2 = e1 + (long long)B * a, e3 = e2 + B, e4 = e3 + B, e5 = e4 + B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float* d; long long j;
    if (i < e0) { d = obs; j = i; }
    else if (i < e1) { d = nobs; j = i - e0; }
    else if (i < e2) { d = act; j = i - e1; }
    else if (i < e3) { d = rew; j = i - e2; }
    else if (i < e4) { d = cost; j = i - e3; }
    else if (i < e5) { d = done; j = i - e4; }
    else { d = init; j = i - e5; }
    if (d) d[j] = src[i];
  }
}

// timestamp node of the in-graph profile (osrl_profile): nanoseconds of the global timer
static __global__ void k_stamp(unsigned long long* out, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  out[slot] = t;
}

// ------------------------------------------------------------------ Adam (+ Polyak target) over a flat range
// torch.optim.Adam single-tensor semantics (bias-corrected, eps after sqrt); optional decoupled
// weight decay (AdamW) and global-norm clip factor; optional fused Polyak update of the target copy
// (bcql.py:114-120) -- valid because no later sub-update of the same step reads that target.
static __global__ void k_adam(float* __restrict__ P, const float* __restrict__ G, float* __restrict__ Mm,
                       float* __restrict__ Vv, float* __restrict__ T, int64_t n4, const DevState* ds, int group,
                       float beta1, float beta2, float w1, float w2, float eps, float weight_decay, float tau,
                       int polyak, float grad_mul, const float* clip_coef) {
  const float step_size = ds->adam_step_size[group];
  const float bc2s = ds->adam_bc2_sqrt[group];
  const float lr = ds->adam_lr[group];
  const float gm = clip_coef ? grad_mul * (*clip_coef) : grad_mul;
  // w1 = float(1 - beta1), w2 = float(1 - beta2) are rounded from the double difference on the host, as torch
  // does (a float 1.f - 0.999f would be off by 1.3e-5 relative)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(P)[i];
    float4 g = reinterpret_cast<const float4*>(G)[i];
    float4 m = reinterpret_cast<float4*>(Mm)[i];
    float4 v = reinterpret_cast<float4*>(Vv)[i];
    float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g.x, g.y, g.z, g.w};
    float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * gm;
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
}

static __global__ void k_polyak(const float* __restrict__ P, float* __restrict__ T, int64_t n4, float tau) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = reinterpret_cast<const float4*>(P)[i];
    float4 t = reinterpret_cast<float4*>(T)[i];
    t.x = tau * p.x + (1.f - tau) * t.x;
    t.y = tau * p.y + (1.f - tau) * t.y;
    t.z = tau * p.z + (1.f - tau) * t.z;
    t.w = tau * p.w + (1.f - tau) * t.w;
    reinterpret_cast<float4*>(T)[i] = t;
  }
}

// ------------------------------------------------------------------ noise: raw N(0,1) via Philox + Box-Muller
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}
// drop_p > 0: the slot holds dropout multipliers (0 with probability p, else 1/(1-p)) instead of normals
struct NoiseSlot { float* dst; long long count; int stream; int enabled; float drop_p; };
static __global__ void k_noise_fill(const NoiseSlot* slots, int nslots, uint64_t seed,
                                    const unsigned long long* __restrict__ step_ctr, uint32_t rank) {
  const NoiseSlot s = slots[blockIdx.y];
  if (!s.enabled) return;
  const uint64_t step = *step_ctr;
  const long long n4 = (s.count + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)step,
                     (uint32_t)(STREAM_NOISE0 + s.stream) | ((uint32_t)(step >> 32) << 8), rank};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
    if (s.drop_p > 0.f) {
      const float keep = 1.f / (1.f - s.drop_p);
#pragma unroll
      for (int k = 0; k < 4; ++k) z[k] = ((float)(c[k] >> 8) * (1.f / 16777216.f) < s.drop_p) ? 0.f : keep;
    } else {
      box_muller(c[0], c[1], z[0], z[1]);
      box_muller(c[2], c[3], z[2], z[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (q * 4 + k < s.count) s.dst[q * 4 + k] = z[k];
  }
}

// ------------------------------------------------------------------ resident dataset: sample + gather
// Packed row: [obs(o) | next_obs(o) | act(a) | r | c | done | pad] ; stride multiple of 4 floats.
// Replaces TransitionDataset.__iter__/__prepare_sample + collate + .to(device)
// (dataset.py:832-847, train_bcql.py:143-146).  One warp per sampled row.
static __global__ void k_sample_gather(const float* __restrict__ ds_rows, int64_t n, int stride, int o, int a,
                                const int64_t* __restrict__ idx_in, uint64_t seed,
                                const unsigned long long* __restrict__ step_ctr, uint32_t rank,
                                int rows, float* obs, float* nobs, float* act, float* rew, float* cost, float* done,
                                int64_t* idx_out, float* init = nullptr) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  int64_t idx;
  if (idx_in) idx = idx_in[w];
  else idx = draw_index(seed, *step_ctr, rank, (uint32_t)w, n);
  if (idx_out && lane == 0) idx_out[w] = idx;
  const float* __restrict__ row = ds_rows + idx * (int64_t)stride;
  for (int c = lane; c < o; c += 32) {
    if (obs) obs[(size_t)w * o + c] = row[c];
    if (nobs) nobs[(size_t)w * o + c] = row[o + c];
  }
  for (int c = lane; c < a; c += 32)
    if (act) act[(size_t)w * a + c] = row[2 * o + c];
  if (lane == 0) {
    if (rew) rew[w] = row[2 * o + a];
    if (cost) cost[w] = row[2 * o + a + 1];
    if (done) done[w] = row[2 * o + a + 2];
    if (init) init[w] = row[2 * o + a + 3];   // (COptiDICE rows carry is_init, dataset.py:817-820)
  }
}

// ------------------------------------------------------------------ column copies (concat / repeat / clamp)
struct CopyTask {
  float* dst; int ldd;
  const float* src; int lds;
  int rows, cols;
  int row_div, row_mod;   // src_row = (r / row_div) % row_mod
  int clamp; float lo, hi;
  float add, mul;          // value = src*mul + add  (then clamp)
};
static __global__ void k_copy_tasks(const CopyTask* tasks) {
  const CopyTask t = tasks[blockIdx.y];
  const long long total = (long long)t.rows * t.cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / t.cols), c = (int)(e % t.cols);
    const int sr = (r / t.row_div) % t.row_mod;
    float v = t.src[(size_t)sr * t.lds + c] * t.mul + t.add;
    if (t.clamp) v = fminf(fmaxf(v, t.lo), t.hi);
    t.dst[(size_t)r * t.ldd + c] = v;
  }
}

// ------------------------------------------------------------------ BC loss (bc.py:45-52)
// u = act_lim * tanh(pre) [B,a] -> loss = mean((u-a)^2); dpre = 2(u-a)/(B a) * act_lim * (1 - (u/act_lim)^2)
static __global__ void k_bc_loss(const float* __restrict__ u, const float* __restrict__ act, int n, float lim,
                          float* __restrict__ dpre, float* stat, float inv_world) {
  __shared__ float sh[33];
  float s = 0.f;
  const float inv = 1.f / (float)n;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const float d = u[e] - act[e];
    s += d * d;
    const float t = u[e] / lim;
    dpre[e] = 2.f * d * inv * inv_world * lim * (1.f - t * t);
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) stat[0] = s * inv;
}

// ------------------------------------------------------------------ VAE (net.py:319-339, bcql.py:122-132)
// ml = [mean | raw_log_std] [B, 2L]; z = mean + exp(clamp(raw,-4,15)) * eps, written into dec_in[:, o:o+L]
static __global__ void k_vae_reparam(const float* __restrict__ ml, const float* __restrict__ eps, int B, int L,
                              float* __restrict__ stdv, float* __restrict__ dec_in, int ldd, int col0) {
  const int n = B * L;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int b = e / L, j = e % L;
    const float mean = ml[(size_t)b * 2 * L + j];
    const float raw = ml[(size_t)b * 2 * L + L + j];
    const float sd = expf(fminf(fmaxf(raw, -4.f), 15.f));
    stdv[e] = sd;
    dec_in[(size_t)b * ldd + col0 + j] = mean + sd * eps[e];
  }
}
// loss_vae = mse(u, act) + beta * KL; dpre3 = d loss / d (d3 pre-activation)
static __global__ void k_vae_loss(const float* __restrict__ u, const float* __restrict__ act, int B, int a, float lim,
                           const float* __restrict__ ml, const float* __restrict__ stdv, int L, float beta,
                           float* __restrict__ dpre3, float* stat, float inv_world) {
  __shared__ float sh[33];
  float s = 0.f;
  const int n = B * a;
  const float inv = 1.f / (float)n;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const float d = u[e] - act[e];
    s += d * d;
    const float t = u[e] / lim;
    dpre3[e] = 2.f * d * inv * inv_world * lim * (1.f - t * t);
  }
  const float recon = block_sum(s, sh) * inv;
  float k = 0.f;
  const int nl = B * L;
  for (int e = threadIdx.x; e < nl; e += blockDim.x) {
    const int b = e / L, j = e % L;
    const float mean = ml[(size_t)b * 2 * L + j];
    const float sd = stdv[e];
    k += 1.f + logf(sd * sd) - mean * mean - sd * sd;
  }
  const float kl = -0.5f * block_sum(k, sh) / (float)nl;
  if (threadIdx.x == 0) stat[0] = recon + beta * kl;
}
// dz = d loss / d dec_in[:, o:o+L]  ->  dml = [dmean | draw_log_std]
static __global__ void k_vae_reparam_bwd(const float* __restrict__ ddec_in, int ldd, int col0, const float* __restrict__ ml,
                                  const float* __restrict__ stdv, const float* __restrict__ eps, int B, int L,
                                  float beta, float* __restrict__ dml, float inv_world) {
  const int n = B * L;
  const float c = beta / (float)n * inv_world;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int b = e / L, j = e % L;
    const float dz = ddec_in[(size_t)b * ldd + col0 + j];
    const float mean = ml[(size_t)b * 2 * L + j];
    const float raw = ml[(size_t)b * 2 * L + L + j];
    const float sd = stdv[e];
    const float dmean = dz + c * mean;
    const float dstd = dz * eps[e] + c * (sd - 1.f / sd);
    const float dls = dstd * sd;
    dml[(size_t)b * 2 * L + j] = dmean;
    dml[(size_t)b * 2 * L + L + j] = (raw >= -4.f && raw <= 15.f) ? dls : 0.f;
  }
}

// ------------------------------------------------------------------ Bellman backup from sampled target Qs
// q [R, n] with R = B*S rows (b-major), n = 2*num nets (first half = q1 list, second = q2 list).
// lambda*min(q1,q2) + (1-lambda)*max(q1,q2), max over the S samples, then
// backup = r + gamma * (1-done)^use_done * max     (bcql.py:143-148, 166-172)
static __global__ void k_q_backup(const float* __restrict__ q, int B, int S, int n, float lmbda, float gamma,
                           const float* __restrict__ r, const float* __restrict__ done, int use_done,
                           float* __restrict__ backup) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int h = n / 2;
  float best = -INFINITY;
  for (int s = 0; s < S; ++s) {
    const float* row = q + ((size_t)b * S + s) * n;
    float q1 = row[0], q2 = row[h];
    for (int i = 1; i < h; ++i) { q1 = fminf(q1, row[i]); q2 = fminf(q2, row[h + i]); }
    const float v = lmbda * fminf(q1, q2) + (1.f - lmbda) * fmaxf(q1, q2);
    best = fmaxf(best, v);
  }
  const float nd = use_done ? (1.f - done[b]) : 1.f;
  backup[b] = r[b] + gamma * nd * best;
}

// ensemble MSE against the backup: loss = sum_i mean_b (q[b,i]-y[b])^2 ; dq = 2 (q-y) / B  (net.py:285-287)
static __global__ void k_critic_loss(const float* __restrict__ q, const float* __restrict__ y, int B, int n,
                              float* __restrict__ dq, float* stat, float inv_world, float extra_const_ptr_mul,
                              const float* extra) {
  __shared__ float sh[33];
  float s = 0.f;
  const float inv = 1.f / (float)B;
  for (int e = threadIdx.x; e < B * n; e += blockDim.x) {
    const float d = q[e] - y[e / n];
    s += d * d;
    dq[e] = 2.f * d * inv * inv_world;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) stat[0] = s * inv + (extra ? extra_const_ptr_mul * extra[0] : 0.f);
}

// Both critics' backup + loss in one launch (block 0: reward critic, block 1: cost critic): the four kernels above are
// four dependent 3-5 us launches on 256 rows otherwise.  Same arithmetic, same order per element.
struct BackupLossArgs {
  const float* tq; int n;              // target ensemble outputs [B*S, n]
  const float* r; int use_done;        // reward / cost; (1 - done) factor only for the reward critic (bcql.py:150,174)
  float* y;                            // [B] backup
  const float* q; float* dq; float* stat;
};
static __global__ void k_backup_critic_loss2(BackupLossArgs a0, BackupLossArgs a1, const float* __restrict__ done, int B,
                                             int S, float lmbda, float gamma, float inv_world) {
  __shared__ float sh[33];
  const BackupLossArgs a = blockIdx.x ? a1 : a0;
  const int n = a.n, h = n / 2;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float best = -INFINITY;
    for (int s = 0; s < S; ++s) {
      const float* row = a.tq + ((size_t)b * S + s) * n;
      float q1 = row[0], q2 = row[h];
      for (int i = 1; i < h; ++i) { q1 = fminf(q1, row[i]); q2 = fminf(q2, row[h + i]); }
      const float v = lmbda * fminf(q1, q2) + (1.f - lmbda) * fmaxf(q1, q2);
      best = fmaxf(best, v);
    }
    const float nd = a.use_done ? (1.f - done[b]) : 1.f;
    a.y[b] = a.r[b] + gamma * nd * best;
  }
  __syncthreads();
  float s = 0.f;
  const float inv = 1.f / (float)B;
  for (int e = threadIdx.x; e < B * n; e += blockDim.x) {
    const float d = a.q[e] - a.y[e / n];
    s += d * d;
    a.dq[e] = 2.f * d * inv * inv_world;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) a.stat[0] = s * inv;
}

// ------------------------------------------------------------------ data-parallel partial means (summed by NCCL)
// out[0] = scale * mean_b( min_i q[b,i] - sub )      (PID error, net.py:380, over the GLOBAL batch)
static __global__ void k_rowmin_mean(const float* __restrict__ q, int n, int B, float sub, float scale, float* out) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float m = q[(size_t)b * n];
    for (int i = 1; i < n; ++i) m = fminf(m, q[(size_t)b * n + i]);
    s += m - sub;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = scale * s / (float)B;
}
static __global__ void k_mean_sub(const float* __restrict__ x, int B, float sub, float scale, float* out) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += x[b] - sub;
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = scale * s / (float)B;
}

// ------------------------------------------------------------------ BCQ-Lag actor loss + PID (bcql.py:181-208, net.py:376-387)
// q [B,nq], qc [B,nqc] (all nets of each double critic).  q_pi = min over all nets; gradient flows to the argmin.
static __global__ void k_bcql_actor_loss(const float* __restrict__ q, int nq, const float* __restrict__ qc, int nqc, int B,
                                  float qc_thres, float kp, float ki, float kd, DevState* ds,
                                  float* __restrict__ dq, float* __restrict__ dqc, float* stat /*[3]*/,
                                  float inv_world, const float* qc_mean_global) {
  __shared__ float sh[33];
  float sq = 0.f, sc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float m = q[(size_t)b * nq];
    for (int i = 1; i < nq; ++i) m = fminf(m, q[(size_t)b * nq + i]);
    float c = qc[(size_t)b * nqc];
    for (int i = 1; i < nqc; ++i) c = fminf(c, qc[(size_t)b * nqc + i]);
    sq += m;
    sc += c - qc_thres;
  }
  const float q_mean = block_sum(sq, sh) / (float)B;
  float e_new = block_sum(sc, sh) / (float)B;
  if (qc_mean_global) e_new = *qc_mean_global;  // data-parallel: mean over the global batch
  __shared__ float mult_s;
  if (threadIdx.x == 0) {
    const float e_diff = fmaxf(e_new - ds->pid_e_old, 0.f);
    const float e_int = fmaxf(ds->pid_e_int + e_new, 0.f);
    ds->pid_e_int = e_int;
    ds->pid_e_old = e_new;
    mult_s = fmaxf(kp * fmaxf(e_new, 0.f) + ki * e_int + kd * e_diff, 0.f);
  }
  __syncthreads();
  const float mult = mult_s;
  const float invB = 1.f / (float)B * inv_world;
  float pen = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int am = 0; float m = q[(size_t)b * nq];
    for (int i = 1; i < nq; ++i) { const float v = q[(size_t)b * nq + i]; if (v < m) { m = v; am = i; } }
    for (int i = 0; i < nq; ++i) dq[(size_t)b * nq + i] = (i == am) ? -invB : 0.f;
    int ac = 0; float c = qc[(size_t)b * nqc];
    for (int i = 1; i < nqc; ++i) { const float v = qc[(size_t)b * nqc + i]; if (v < c) { c = v; ac = i; } }
    for (int i = 0; i < nqc; ++i) dqc[(size_t)b * nqc + i] = (i == ac) ? mult * invB : 0.f;
    pen += (c - qc_thres) * mult;
  }
  const float qc_pen = block_sum(pen, sh) / (float)B;
  if (threadIdx.x == 0) {
    stat[0] = -q_mean + qc_pen;  // loss/actor_loss
    stat[1] = qc_pen;            // loss/qc_penalty
    stat[2] = mult;              // loss/lagrangian
  }
}

// a = clamp(phi*lim*t + a_vae, +-lim), t = tanh(l3): d l3pre = (da_q + da_qc) * [|.|<=lim] * phi*lim*(1-t^2)
static __global__ void k_perturb_bwd(const float* __restrict__ da1, const float* __restrict__ da2, int ld_da,
                              const float* __restrict__ t, const float* __restrict__ avae, int ld_av, int B, int a,
                              float philim, float lim, float* __restrict__ dpre) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * a; e += gridDim.x * blockDim.x) {
    const int b = e / a, j = e % a;
    const float tv = t[e];
    const float pre = philim * tv + avae[(size_t)b * ld_av + j];
    const float g = da1[(size_t)b * ld_da + j] + (da2 ? da2[(size_t)b * ld_da + j] : 0.f);
    dpre[e] = (pre >= -lim && pre <= lim) ? g * philim * (1.f - tv * tv) : 0.f;
  }
}

}  // namespace osrl

// ====================================================================== CPQ / BEAR-Lag kernels
namespace osrl {

// ------------------------------------------------------------------ squashed-Gaussian sampling (net.py:169-205)
// mh = [mu | raw_log_std] [*, 2a]; u = mu + exp(clamp(ls,-20,2)) * eps; out = do_tanh ? scale*tanh(u) : u
struct SquashTask {
  const float* mh; int row_div, row_mod;  // source row of (mu, log_std) = (r / row_div) % row_mod
  const float* eps;                       // [rows, a]
  float* dst; int ldd;                    // dst[r*ldd + j]
  float* u_out;                           // optional raw u [rows, a]
  int rows, a, do_tanh;
  float scale;
};
static __global__ void k_squash_tasks(const SquashTask* tasks) {
  const SquashTask t = tasks[blockIdx.y];
  const int n = t.rows * t.a;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int r = e / t.a, j = e % t.a;
    const int sr = (r / t.row_div) % t.row_mod;
    const float mu = t.mh[(size_t)sr * 2 * t.a + j];
    const float ls = fminf(fmaxf(t.mh[(size_t)sr * 2 * t.a + t.a + j], -20.f), 2.f);
    const float u = mu + expf(ls) * t.eps[e];
    if (t.u_out) t.u_out[e] = u;
    t.dst[(size_t)r * t.ldd + j] = t.do_tanh ? t.scale * tanhf(u) : u;
  }
}

// ------------------------------------------------------------------ CPQ Bellman backups (cpq.py:140-146, 158-161)
__device__ __forceinline__ float row_min(const float* p, int n) {
  float m = p[0];
  for (int i = 1; i < n; ++i) m = fminf(m, p[i]);
  return m;
}
static __global__ void k_cpq_backup(const float* __restrict__ tq, int nq, const float* __restrict__ tqc1,
                                    const float* __restrict__ tqc2, int nqc, int B, float gamma, float q_thres,
                                    const float* __restrict__ r, const float* __restrict__ c,
                                    const float* __restrict__ done, float* __restrict__ yq, float* __restrict__ yc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float q = row_min(tq + (size_t)b * nq, nq);
  const float qc1 = row_min(tqc1 + (size_t)b * nqc, nqc);
  const float qc2 = row_min(tqc2 + (size_t)b * nqc, nqc);
  yq[b] = r[b] + gamma * (1.f - done[b]) * (qc1 <= q_thres ? 1.f : 0.f) * q;
  yc[b] = c[b] + gamma * qc2;
}

// per-row KL of the OOD samples (cpq.py:178-182) and ensemble-min cost Q
static __global__ void k_cpq_kl_rows(const float* __restrict__ ml, int L, const float* __restrict__ qc, int nqc,
                                     int rows, float* __restrict__ kl, float* __restrict__ qcmin) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int j = 0; j < L; ++j) {
    const float mean = ml[(size_t)r * 2 * L + j];
    const float sd = expf(fminf(fmaxf(ml[(size_t)r * 2 * L + L + j], -4.f), 15.f));
    s += 1.f + logf(sd * sd) - mean * mean - sd * sd;
  }
  kl[r] = -0.5f * (s / (float)L);
  qcmin[r] = row_min(qc + (size_t)r * nqc, nqc);
}

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// k-th smallest (0-based) of v[0..n) by 4-pass radix select; single CTA; sh: 256+2 uints
static __device__ uint32_t block_select(const float* __restrict__ v, int n, int kth, uint32_t* hist) {
  uint32_t prefix = 0, mask = 0;
  int k = kth;
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = f2key(v[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, d = 0;
      for (; d < 256; ++d) {
        if (acc + (int)hist[d] > k) break;
        acc += hist[d];
      }
      hist[256] = (uint32_t)d;
      hist[257] = (uint32_t)(k - acc);
    }
    __syncthreads();
    prefix |= hist[256] << shift;
    mask |= 255u << shift;
    k = (int)hist[257];
    __syncthreads();
  }
  return prefix;
}
// torch.quantile(kl, 0.75) (linear interpolation) -> qc_ood -> dual step on log_alpha (cpq.py:183-195)
// kl, qcmin are [S, B] (S-major).  stat_extra[0] = exp(log_alpha_old) * (mean qc_ood - qc_thres)
static __global__ void k_cpq_ood(const float* __restrict__ kl, const float* __restrict__ qcmin, int S, int B,
                                 float qfrac, float qc_thres, float alpha_lr, DevState* ds, float* stat_extra,
                                 float* stat_alpha) {
  __shared__ uint32_t hist[258];
  __shared__ float sh[33];
  const int n = S * B;
  const double pos = (double)qfrac * (double)(n - 1);
  const int lo = (int)floor(pos), hi = (int)ceil(pos);
  const float w = (float)(pos - (double)lo);
  const float vlo = key2f(block_select(kl, n, lo, hist));
  const float vhi = (hi == lo) ? vlo : key2f(block_select(kl, n, hi, hist));
  const float quant = vlo + w * (vhi - vlo);   // torch lerp, weight < 0.5 form; (>=0.5 differs by <=1ulp)
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (kl[i] >= quant) ? qcmin[i] : 0.f;
  const float tot = block_sum(s, sh);
  if (threadIdx.x == 0) {
    const float mean_ood = tot / (float)S / (float)B;  // mean over s, then mean over b
    const float la = ds->log_alpha;
    stat_extra[0] = expf(la) * (mean_ood - qc_thres);
    float nla = la + alpha_lr * expf(la) * (qc_thres - mean_ood);
    nla = fminf(fmaxf(nla, -5.f), 5.f);
    ds->log_alpha = nla;
    stat_alpha[0] = expf(nla);
  }
}

// CPQ actor loss (cpq.py:203-222): loss = -mean(1[qc<=q_thres] * q); grad only through q's argmin net
static __global__ void k_cpq_actor_loss(const float* __restrict__ q, int nq, const float* __restrict__ qc, int nqc,
                                        int B, float q_thres, float* __restrict__ dq, float* stat, float inv_world) {
  __shared__ float sh[33];
  float s = 0.f;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int am = 0; float m = q[(size_t)b * nq];
    for (int i = 1; i < nq; ++i) { const float v = q[(size_t)b * nq + i]; if (v < m) { m = v; am = i; } }
    const float gate = row_min(qc + (size_t)b * nqc, nqc) <= q_thres ? 1.f : 0.f;
    s += gate * m;
    for (int i = 0; i < nq; ++i) dq[(size_t)b * nq + i] = (i == am) ? -gate * invB * inv_world : 0.f;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) stat[0] = -s * invB;
}

// a = scale*tanh(u), u = mu + std*eps: d[mu | raw_log_std] from da (one sample per row)
static __global__ void k_squash_bwd(const float* __restrict__ da, int ld_da, const float* __restrict__ u,
                                    const float* __restrict__ mh, const float* __restrict__ eps, int B, int a,
                                    float scale, float* __restrict__ dmh) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * a; e += gridDim.x * blockDim.x) {
    const int b = e / a, j = e % a;
    const float t = tanhf(u[e]);
    const float du = da[(size_t)b * ld_da + j] * scale * (1.f - t * t);
    const float raw = mh[(size_t)b * 2 * a + a + j];
    const float sd = expf(fminf(fmaxf(raw, -20.f), 2.f));
    dmh[(size_t)b * 2 * a + j] = du;
    dmh[(size_t)b * 2 * a + a + j] = (raw >= -20.f && raw <= 2.f) ? du * eps[e] * sd : 0.f;
  }
}

// ------------------------------------------------------------------ BEAR-Lag MMD (bearl.py:283-318)
#define OSRL_MMD_MAXN 32
#define OSRL_MMD_MAXA 16
__device__ __forceinline__ float mmd_k(const float* x, const float* y, int a, float inv2s, int laplacian) {
  float d = 0.f;
  for (int j = 0; j < a; ++j) {
    const float t = x[j] - y[j];
    d += laplacian ? fabsf(t) : t * t;
  }
  return expf(-d * inv2s);
}
// x = raw VAE decodes [B,N,a]; y = raw actor samples u [B,N,a]; mmd [B]
static __global__ void k_mmd_fwd(const float* __restrict__ x, const float* __restrict__ y, int B, int N, int a,
                                 float sigma, int laplacian, float* __restrict__ mmd) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float inv2s = 1.f / (2.f * sigma);
  const float* xb = x + (size_t)b * N * a;
  const float* yb = y + (size_t)b * N * a;
  // the three kernel means are O(1) and their combination is O(1e-3): accumulate in double so the
  // cancellation does not amplify fp32 summation error (torch's pairwise mean is similarly accurate)
  double kxx = 0.0, kyy = 0.0, kxy = 0.0;
  for (int i = 0; i < N; ++i)
    for (int m = 0; m < N; ++m) {
      kxx += (double)mmd_k(xb + i * a, xb + m * a, a, inv2s, laplacian);
      kyy += (double)mmd_k(yb + i * a, yb + m * a, a, inv2s, laplacian);
      kxy += (double)mmd_k(xb + i * a, yb + m * a, a, inv2s, laplacian);
    }
  const double nn = (double)(N * N);
  mmd[b] = sqrtf((float)(kxx / nn + kyy / nn - 2.0 * (kxy / nn) + 1e-6));
}

// BEAR actor loss + PID + dual step (bearl.py:240-262).  q/qc: all nets of each double critic.
static __global__ void k_bear_actor_loss(const float* __restrict__ q, int nq, const float* __restrict__ qc, int nqc,
                                         const float* __restrict__ mmd, int B, float qc_thres, float kp, float ki,
                                         float kd, float mmd_thresh, float alpha_lr, int start_step, DevState* ds,
                                         float* __restrict__ dq, float* __restrict__ dqc, float* stat /*[5]*/,
                                         float* mmd_coef, float inv_world, const float* global2) {
  __shared__ float sh[33];
  __shared__ float mult_s, gate_s, ealpha_s;
  float sq = 0.f, sc = 0.f, sm = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    sq += row_min(q + (size_t)b * nq, nq);
    sc += row_min(qc + (size_t)b * nqc, nqc) - qc_thres;
    sm += mmd[b] - mmd_thresh;
  }
  const float q_mean = block_sum(sq, sh) / (float)B;
  float e_new = block_sum(sc, sh) / (float)B;
  float mmd_mean_c = block_sum(sm, sh) / (float)B;  // mean(mmd - thresh)
  const float mmd_mean_local = mmd_mean_c;
  if (global2) { e_new = global2[0]; mmd_mean_c = global2[1]; }  // data-parallel: means over the global batch
  if (threadIdx.x == 0) {
    const float e_diff = fmaxf(e_new - ds->pid_e_old, 0.f);
    const float e_int = fmaxf(ds->pid_e_int + e_new, 0.f);
    ds->pid_e_int = e_int;
    ds->pid_e_old = e_new;
    mult_s = fmaxf(kp * fmaxf(e_new, 0.f) + ki * e_int + kd * e_diff, 0.f);
    gate_s = ds->n_train_steps >= start_step ? 1.f : 0.f;
    ealpha_s = expf(ds->log_alpha);
  }
  __syncthreads();
  const float mult = mult_s, gate = gate_s, ealpha = ealpha_s;
  const float invB = 1.f / (float)B * inv_world;
  float pen = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int am = 0; float m = q[(size_t)b * nq];
    for (int i = 1; i < nq; ++i) { const float v = q[(size_t)b * nq + i]; if (v < m) { m = v; am = i; } }
    for (int i = 0; i < nq; ++i) dq[(size_t)b * nq + i] = (i == am) ? -gate * invB : 0.f;
    int ac = 0; float c = qc[(size_t)b * nqc];
    for (int i = 1; i < nqc; ++i) { const float v = qc[(size_t)b * nqc + i]; if (v < c) { c = v; ac = i; } }
    for (int i = 0; i < nqc; ++i) dqc[(size_t)b * nqc + i] = (i == ac) ? mult * invB : 0.f;
    pen += (c - qc_thres) * mult;
  }
  const float qc_pen = block_sum(pen, sh) / (float)B;
  if (threadIdx.x == 0) {
    stat[0] = -gate * q_mean + ealpha * mmd_mean_local + qc_pen;  // loss/actor_loss
    stat[1] = mmd_mean_local + mmd_thresh;                     // loss/mmd_loss
    stat[2] = qc_pen;                                          // loss/qc_penalty
    stat[3] = mult;                                            // loss/lagrangian
    mmd_coef[0] = ealpha * invB;                               // d loss / d mmd[b]
    float nla = ds->log_alpha + alpha_lr * ealpha * mmd_mean_c;
    nla = fminf(fmaxf(nla, -5.f), 5.f);
    ds->log_alpha = nla;
    ds->n_train_steps += 1;
    stat[4] = expf(nla);                                       // loss/alpha_value
  }
}

// d loss / d [mu | raw_log_std] of the actor from (i) the MMD term on all N samples and
// (ii) the Q terms on sample 0 (da = d loss / d tanh(u[b,0]))
static __global__ void k_bear_actor_bwd(const float* __restrict__ x, const float* __restrict__ y,
                                        const float* __restrict__ mmd, const float* __restrict__ mmd_coef,
                                        const float* __restrict__ da1, const float* __restrict__ da2, int ld_da,
                                        const float* __restrict__ mh, const float* __restrict__ eps, int B, int N,
                                        int a, float sigma, int laplacian, float* __restrict__ dmh) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float inv2s = 1.f / (2.f * sigma);
  const float* xb = x + (size_t)b * N * a;
  const float* yb = y + (size_t)b * N * a;
  const float nn = (float)(N * N);
  const float gm = mmd_coef[0] / (2.f * mmd[b]);  // d loss / d (mmd^2 argument)
  double dmu[OSRL_MMD_MAXA], dls[OSRL_MMD_MAXA];  // double: the k(y,y) and k(x,y) terms nearly cancel
  for (int j = 0; j < a; ++j) { dmu[j] = 0.0; dls[j] = 0.0; }
  for (int n = 0; n < N; ++n) {
    double du[OSRL_MMD_MAXA];
    for (int j = 0; j < a; ++j) du[j] = 0.0;
    for (int m = 0; m < N; ++m) {
      const float kyy = mmd_k(yb + n * a, yb + m * a, a, inv2s, laplacian);
      const float kxy = mmd_k(xb + m * a, yb + n * a, a, inv2s, laplacian);
      for (int j = 0; j < a; ++j) {
        const float dyy = yb[n * a + j] - yb[m * a + j];
        const float dxy = xb[m * a + j] - yb[n * a + j];
        float gyy, gxy;  // d k / d y_n[j]
        if (laplacian) {
          gyy = -kyy * inv2s * (dyy > 0.f ? 1.f : (dyy < 0.f ? -1.f : 0.f));
          gxy = kxy * inv2s * (dxy > 0.f ? 1.f : (dxy < 0.f ? -1.f : 0.f));
        } else {
          gyy = -kyy * 2.f * inv2s * dyy;
          gxy = kxy * 2.f * inv2s * dxy;
        }
        // mean k(y,y): y_n appears in row n and column n -> factor 2;  -2 * mean k(x,y)
        du[j] += (double)gm * (2.0 * (double)gyy - 2.0 * (double)gxy) / (double)nn;
      }
    }
    for (int j = 0; j < a; ++j) {
      if (n == 0) {
        const float t = tanhf(yb[j]);
        du[j] += (double)((da1[(size_t)b * ld_da + j] + da2[(size_t)b * ld_da + j]) * (1.f - t * t));
      }
      dmu[j] += du[j];
      dls[j] += du[j] * (double)eps[((size_t)b * N + n) * a + j];
    }
  }
  for (int j = 0; j < a; ++j) {
    const float raw = mh[(size_t)b * 2 * a + a + j];
    const float sd = expf(fminf(fmaxf(raw, -20.f), 2.f));
    dmh[(size_t)b * 2 * a + j] = (float)dmu[j];
    dmh[(size_t)b * 2 * a + a + j] = (raw >= -20.f && raw <= 2.f) ? (float)(dls[j] * (double)sd) : 0.f;
  }
}

}  // namespace osrl

namespace osrl {

// ====================================================================== COptiDICE (coptidice.py:125-227)
// f-divergence pieces (get_f_div_fn, coptidice.py:15-38): f_type 0 chi2, 1 softchi, 2 kl
__device__ __forceinline__ float cop_fprime_inv(float x, int ft) {
  if (ft == 0) return x + 1.f;
  if (ft == 1) return x < 0.f ? expf(fminf(x, 0.f)) : x + 1.f;
  return expf(x - 1.f);
}
__device__ __forceinline__ float cop_f(float w, int ft) {
  if (ft == 0) return 0.5f * (w - 1.f) * (w - 1.f);
  if (ft == 1) return w < 1.f ? w * (logf(w + 1e-10f) - 1.f) + 1.f : 0.5f * (w - 1.f) * (w - 1.f);
  return w * logf(w + 1e-10f);
}
__device__ __forceinline__ float cop_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // F.softplus
__device__ __forceinline__ float block_max(float v, float* sh /*[33]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = (l < (int)(blockDim.x >> 5)) ? sh[l] : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    if (l == 0) sh[32] = x;
  }
  __syncthreads();
  return sh[32];
}
// one scalar Adam step with torch's constants (bias corrections in double, like k_prologue)
__device__ __forceinline__ float cop_scalar_adam(float p, float g, float& m, float& v, int& t, float lr) {
  t += 1;
  m = m + 0.1f * (g - m);                       // torch: m.lerp_(g, 1 - beta1)
  v = 0.999f * v + 0.001f * g * g;              // v.mul_(beta2).addcmul_(g, g, 1 - beta2)
  const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
  const float step = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
  return p - step * (m / (sqrtf(v) / bc2s + 1e-8f));
}

struct CopArgs {
  const float *q_nu, *q_chi;     // [2B, n]: rows 0..B-1 on observations, B..2B-1 on next_observations
  int n_nu, n_chi, B;
  const float *rew, *cost, *done, *init;
  float gamma, alpha, eps, p0, thres, scalar_lr;
  int ftype;
  float *dq_nu, *dq_chi;         // [2B, n] gradients wrt the ensemble outputs (zero except at each row's arg-min)
  float *e_buf, *w_buf, *ell_buf;  // [B] scratch
  float* stats;                  // chi_loss, tau_loss, D_kl, Df, td_error, nu_loss, lmbda_loss, actor_loss, tau, lmbda
};
__device__ __forceinline__ float cop_min(const float* __restrict__ q, int n, int& arg) {
  float m = q[0];
  arg = 0;
  for (int i = 1; i < n; ++i)
    if (q[i] < m) { m = q[i]; arg = i; }   // first minimum, like torch.min(dim) (net.py:235-238)
  return m;
}
// nu / chi / tau / lambda phase: every loss, every gradient wrt the network outputs, both scalar Adam steps.
// One CTA: the chi weights are a softmax over the BATCH (:167-168).  d(nu_loss)/de_b = w_b / B: the derivative of
// w e - alpha f(w) through w vanishes because f'(w) = e / alpha wherever w > 0 (w = relu(f'^-1(e / alpha))).
static __global__ void __launch_bounds__(1024) k_cop_main(CopArgs a, DevState* ds) {
  __shared__ float sh[33];
  const int B = a.B, tid = threadIdx.x;
  const float lm = cop_softplus(ds->cop_lmbda), tau = cop_softplus(ds->cop_tau);
  const float invB = 1.f / (float)B, g1 = 1.f - a.gamma;
  float s_f = 0.f, s_e2 = 0.f, s_init = 0.f, s_main = 0.f, s_wc = 0.f;
  for (int i = tid; i < 2 * B * a.n_nu; i += blockDim.x) a.dq_nu[i] = 0.f;
  if (a.eps != 0.f)
    for (int i = tid; i < 2 * B * a.n_chi; i += blockDim.x) a.dq_chi[i] = 0.f;
  __syncthreads();
  float lmax = -3.0e38f;
  for (int b = tid; b < B; b += blockDim.x) {
    int as, an;
    const float nu_s = cop_min(a.q_nu + (size_t)b * a.n_nu, a.n_nu, as);
    const float nu_n = cop_min(a.q_nu + (size_t)(B + b) * a.n_nu, a.n_nu, an);
    const float nd = a.gamma * (1.f - a.done[b]);
    const float e = (a.rew[b] - lm * a.cost[b]) + nd * nu_n - nu_s;
    const float w = fmaxf(cop_fprime_inv(e / a.alpha, a.ftype), 0.f);
    const float fw = cop_f(w, a.ftype), ini = a.init[b] / a.p0;
    a.e_buf[b] = e; a.w_buf[b] = w;
    s_f += fw; s_e2 += e * e; s_init += nu_s * ini; s_main += w * e - a.alpha * fw; s_wc += w * a.cost[b];
    a.dq_nu[(size_t)b * a.n_nu + as] = (g1 * ini - w) * invB;
    a.dq_nu[(size_t)(B + b) * a.n_nu + an] = w * nd * invB;
    if (a.eps != 0.f) {
      int cs, cn;
      const float chi_s = cop_min(a.q_chi + (size_t)b * a.n_chi, a.n_chi, cs);
      const float chi_n = cop_min(a.q_chi + (size_t)(B + b) * a.n_chi, a.n_chi, cn);
      const float ell = g1 * (chi_s * ini) + w * (a.cost[b] + nd * chi_n - chi_s);
      a.ell_buf[b] = ell;
      lmax = fmaxf(lmax, ell / tau);
    }
  }
  const float Df = block_sum(s_f, sh) * invB, td = block_sum(s_e2, sh) * invB;
  const float nu_loss = g1 * (block_sum(s_init, sh) * invB) + block_sum(s_main, sh) * invB;
  float weighted_c = block_sum(s_wc, sh) * invB, chi_loss = 0.f, tau_loss = 0.f, Dkl = 0.f;
  if (a.eps != 0.f) {
    lmax = block_max(lmax, sh);
    float z = 0.f;
    for (int b = tid; b < B; b += blockDim.x) z += expf(a.ell_buf[b] / tau - lmax);
    const float Z = block_sum(z, sh), logZ = logf(Z), logB = logf((float)B);
    float s_kl = 0.f, s_c = 0.f, s_l = 0.f, s_sl = 0.f;
    for (int b = tid; b < B; b += blockDim.x) {
      const float ell = a.ell_buf[b], lg = ell / tau - lmax;
      const float sm = expf(lg) / Z, wt = sm * (float)B, lw = (lg - logZ) + logB;
      s_kl += wt * lw - wt + 1.f; s_c += wt * a.w_buf[b] * a.cost[b]; s_l += wt * ell; s_sl += sm * ell;
    }
    Dkl = block_sum(s_kl, sh) * invB;
    weighted_c = block_sum(s_c, sh) * invB;
    chi_loss = block_sum(s_l, sh) * invB;
    const float S_ell = block_sum(s_sl, sh);
    for (int b = tid; b < B; b += blockDim.x) {   // d chi_loss / d ell_b = s_b + (s_b / tau)(ell_b - sum_j s_j ell_j)
      const float ell = a.ell_buf[b], sm = expf(ell / tau - lmax) / Z;
      const float dl = sm + (sm / tau) * (ell - S_ell);
      const float w = a.w_buf[b], ini = a.init[b] / a.p0, nd = a.gamma * (1.f - a.done[b]);
      int cs, cn;
      cop_min(a.q_chi + (size_t)b * a.n_chi, a.n_chi, cs);
      cop_min(a.q_chi + (size_t)(B + b) * a.n_chi, a.n_chi, cn);
      a.dq_chi[(size_t)b * a.n_chi + cs] = dl * (g1 * ini - w);
      a.dq_chi[(size_t)(B + b) * a.n_chi + cn] = dl * w * nd;
    }
    tau_loss = tau * (a.eps - Dkl);
  }
  if (tid == 0) {
    const float lm_loss = lm * (a.thres - weighted_c);
    a.stats[0] = chi_loss; a.stats[1] = tau_loss; a.stats[2] = Dkl; a.stats[3] = Df; a.stats[4] = td;
    a.stats[5] = nu_loss; a.stats[6] = lm_loss; a.stats[8] = tau; a.stats[9] = lm;
    ds->cop_lm_old = lm;
    // d softplus(x)/dx = sigmoid(x); tau first, then lambda (coptidice.py:177-193); gradients predate both steps
    if (a.eps != 0.f) {
      const float g = (1.f / (1.f + expf(-ds->cop_tau))) * (a.eps - Dkl);
      ds->cop_tau = cop_scalar_adam(ds->cop_tau, g, ds->cop_m[0], ds->cop_v[0], ds->cop_t[0], a.scalar_lr);
    }
    const float gl = (1.f / (1.f + expf(-ds->cop_lmbda))) * (a.thres - weighted_c);
    ds->cop_lmbda = cop_scalar_adam(ds->cop_lmbda, gl, ds->cop_m[1], ds->cop_v[1], ds->cop_t[1], a.scalar_lr);
  }
}
// noisy inputs of the policy phase (coptidice.py:201-202): x + eps * std * 0.1
static __global__ void k_cop_noise(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ sd,
                                   int rows, int d, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * d) out[i] = x[i] + (eps[i] * sd[i % d]) * 0.1f;
}
// policy extraction (coptidice.py:203-212): w from the UPDATED nu network and the step's original lambda',
// loss = -mean(w log N(x; mu, std)) with the pre-tanh Normal; gradient wrt (mu | log_std-before-clamp)
static __global__ void __launch_bounds__(1024) k_cop_actor_loss(const float* __restrict__ q_nu, int n_nu, int B, int adim,
                                                                const float* __restrict__ rew, const float* __restrict__ cost,
                                                                const float* __restrict__ done, float gamma, float alpha,
                                                                int ftype, const float* __restrict__ mh,
                                                                const float* __restrict__ xact, float* __restrict__ dmh,
                                                                float* stat, const DevState* ds) {
  __shared__ float sh[33];
  const float lm = ds->cop_lm_old, invB = 1.f / (float)B;
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int dummy;
    const float nu_s = cop_min(q_nu + (size_t)b * n_nu, n_nu, dummy);
    const float nu_n = cop_min(q_nu + (size_t)(B + b) * n_nu, n_nu, dummy);
    const float e = (rew[b] - lm * cost[b]) + gamma * (1.f - done[b]) * nu_n - nu_s;
    const float w = fmaxf(cop_fprime_inv(e / alpha, ftype), 0.f);
    float logp = 0.f;
    for (int d = 0; d < adim; ++d) {
      const float mu = mh[(size_t)b * 2 * adim + d], raw = mh[(size_t)b * 2 * adim + adim + d];
      const float ls = fminf(fmaxf(raw, -20.f), 2.f), sd = expf(ls), z = (xact[(size_t)b * adim + d] - mu) / sd;
      logp += -0.5f * z * z - ls - 0.91893853320467274f;
      dmh[(size_t)b * 2 * adim + d] = -w * invB * (z / sd);
      dmh[(size_t)b * 2 * adim + adim + d] = (raw >= -20.f && raw <= 2.f) ? -w * invB * (z * z - 1.f) : 0.f;
    }
    s += w * logp;
  }
  const float tot = block_sum(s, sh);
  if (threadIdx.x == 0) *stat = -tot * invB;
}

}  // namespace osrl
