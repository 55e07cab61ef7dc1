// CDT kernels (sm_100a): LayerNorm fwd/bwd, 40-token causal+padding attention fwd/bwd, token
// embedding glue, head losses, gradient-norm clip.  Replaces TransformerBlock (net.py:391-441),
// CDT.forward (cdt.py:166-265) and the loss part of CDTTrainer.train_one_step (cdt.py:343-418).
// The linear projections (98 % of CDT's FLOPs) run on the shared multi-task GEMM.
#pragma once
#include "kernels.cuh"

namespace osrl {

// ------------------------------------------------------------------ token glue
// te[bt,:] = timestep_emb[time_steps[bt],:] (cdt.py:180); ctg_t = 50 - ctg (cost_transform, cdt.py:79,187)
static __global__ void k_cdt_prep(const long long* __restrict__ ts, const float* __restrict__ ctg, int BT, int E,
                                  const float* __restrict__ table, int table_rows, float* __restrict__ te,
                                  float* __restrict__ ctg_t) {
  const int n = BT * E;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int bt = e / E, c = e % E;
    long long r = ts[bt];
    r = r < 0 ? 0 : (r >= table_rows ? table_rows - 1 : r);
    te[e] = table[(size_t)r * E + c];
    if (c == 0) ctg_t[bt] = 50.f - ctg[bt];
  }
}
// d timestep_emb[ts[bt], :] += sum over the 4 tokens of step bt of d x0   (embedding backward)
static __global__ void k_cdt_te_scatter(const long long* __restrict__ ts, const float* __restrict__ dx0, int BT, int E,
                                        int table_rows, float* __restrict__ gtable) {
  const int n = BT * E;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int bt = e / E, c = e % E;
    long long r = ts[bt];
    r = r < 0 ? 0 : (r >= table_rows ? table_rows - 1 : r);
    const float* p = dx0 + (size_t)bt * 4 * E + c;
    atomicAdd(gtable + (size_t)r * E + c, p[0] + p[E] + p[2 * E] + p[3 * E]);
  }
}

// ------------------------------------------------------------------ trajectory buffer: sample + gather
// Replaces SequenceDataset.__iter__/__prepare_sample + collate + .to(device) (dataset.py:749-787).
// Packed transition row: [obs(o) | act(a) | return_to_go*reward_scale | cost_to_go*cost_scale | cost | pad].
// traj ~ Categorical(sample_prob) through an alias table, start ~ U[0, len-1] (Philox), T rows from `start`,
// zero-padded past the trajectory end with mask = 0; time_steps = start + t (not clipped, dataset.py:755).
enum { STREAM_SEQ = 2 };
__host__ __device__ __forceinline__ void draw_sequence(uint64_t seed, uint64_t step, uint32_t rank, uint32_t i,
                                                       const float* prob, const int* alias, const long long* off,
                                                       int n_traj, int& traj, int& start) {
  uint32_t c[4] = {i, (uint32_t)step, (uint32_t)STREAM_SEQ | ((uint32_t)(step >> 32) << 8), rank};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const int slot = (int)(((uint64_t)c[0] * (uint64_t)n_traj) >> 32);
  const float u = (float)(c[1] >> 8) * (1.0f / 16777216.0f);
  traj = (u < prob[slot]) ? slot : alias[slot];
  const long long len = off[traj + 1] - off[traj];
  start = (int)(((uint64_t)c[2] * (uint64_t)len) >> 32);
}
static __global__ void k_seq_gather(const float* __restrict__ rows_, const long long* __restrict__ off, int n_traj,
                                    int stride, int o, int a, int T, const float* __restrict__ prob,
                                    const int* __restrict__ alias, const int* __restrict__ traj_in,
                                    const int* __restrict__ start_in, uint64_t seed, const DevState* st, uint32_t rank,
                                    int B, float* states, float* actions, float* returns, float* ctg, long long* ts,
                                    float* mask, float* costs, int* traj_out, int* start_out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * T) return;
  const int b = w / T, t = w % T;
  int traj, start;
  if (traj_in) { traj = traj_in[b]; start = start_in[b]; }
  else draw_sequence(seed, st->step, rank, (uint32_t)b, prob, alias, off, n_traj, traj, start);
  if (t == 0 && lane == 0) {
    if (traj_out) traj_out[b] = traj;
    if (start_out) start_out[b] = start;
  }
  const long long len = off[traj + 1] - off[traj];
  const bool valid = (long long)start + t < len;
  const float* __restrict__ row = rows_ + (off[traj] + start + t) * (long long)stride;
  for (int c = lane; c < o; c += 32) states[(size_t)w * o + c] = valid ? row[c] : 0.f;
  for (int c = lane; c < a; c += 32) actions[(size_t)w * a + c] = valid ? row[o + c] : 0.f;
  if (lane == 0) {
    returns[w] = valid ? row[o + a] : 0.f;
    ctg[w] = valid ? row[o + a + 1] : 0.f;
    costs[w] = valid ? row[o + a + 2] : 0.f;
    mask[w] = valid ? 1.f : 0.f;
    ts[w] = (long long)start + t;
  }
}

// ------------------------------------------------------------------ LayerNorm (eps 1e-5, biased variance)
static __global__ void k_ln_fwd(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                int E) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= rows) return;
  const float* xr = x + (size_t)w * E;
  float s = 0.f;
  for (int c = lane; c < E; c += 32) s += xr[c];
  s = warp_sum(s);
  const float mu = s / (float)E;
  float v = 0.f;
  for (int c = lane; c < E; c += 32) { const float d = xr[c] - mu; v += d * d; }
  v = warp_sum(v);
  const float rs = rsqrtf(v / (float)E + 1e-5f);
  if (lane == 0) { mean[w] = mu; rstd[w] = rs; }
  float* yr = y + (size_t)w * E;
  for (int c = lane; c < E; c += 32) yr[c] = (xr[c] - mu) * rs * g[c] + b[c];
}
// dx (= or +=) and per-block partial sums of d gamma / d beta.  blockDim = 256 (8 warps), E = 32*PER <= 512.
#define OSRL_LN_MAXE 512
template <int PER>
static __global__ void k_ln_bwd(const float* __restrict__ dy, const float* __restrict__ x,
                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ g, float* __restrict__ dx, int accumulate,
                                float* __restrict__ part_dg, float* __restrict__ part_db, int rows, int E) {
  __shared__ float sg[OSRL_LN_MAXE], sb[OSRL_LN_MAXE];
  for (int c = threadIdx.x; c < E; c += blockDim.x) { sg[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float ag[PER], ab[PER];
  constexpr int per = PER;
#pragma unroll
  for (int q = 0; q < per; ++q) { ag[q] = 0.f; ab[q] = 0.f; }
  for (int r = blockIdx.x * wpb + wib; r < rows; r += gridDim.x * wpb) {
    const float mu = mean[r], rs = rstd[r];
    const float* xr = x + (size_t)r * E;
    const float* dr = dy + (size_t)r * E;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < per; ++q) {
      const int c = lane + 32 * q;
      const float xh = (xr[c] - mu) * rs, dg = dr[c] * g[c];
      s1 += dg; s2 += dg * xh;
      ag[q] += dr[c] * xh; ab[q] += dr[c];
    }
    s1 = warp_sum(s1) / (float)E; s2 = warp_sum(s2) / (float)E;
    float* dxr = dx + (size_t)r * E;
#pragma unroll
    for (int q = 0; q < per; ++q) {
      const int c = lane + 32 * q;
      const float xh = (xr[c] - mu) * rs;
      const float v = rs * (dr[c] * g[c] - s1 - xh * s2);
      dxr[c] = accumulate ? dxr[c] + v : v;
    }
  }
#pragma unroll
  for (int q = 0; q < per; ++q) { atomicAdd(&sg[lane + 32 * q], ag[q]); atomicAdd(&sb[lane + 32 * q], ab[q]); }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    part_dg[(size_t)blockIdx.x * E + c] = sg[c];
    part_db[(size_t)blockIdx.x * E + c] = sb[c];
  }
}
// d gamma / d beta = column sums of the per-block partials.  blockDim = (32 columns, 16 row groups): group g adds blocks
// g, g + 16, ... (four loads in flight), the 16 group sums are then added in group order -- deterministic, and 16 x
// shorter than one thread walking all `nblk` partials of its column (0.28 ms per CDT step at B = 2048).
static __global__ void k_ln_param_reduce(const float* __restrict__ part_dg, const float* __restrict__ part_db, int nblk,
                                         int E, float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float sa[16][33], sb_[16][33];
  const int c = blockIdx.x * 32 + threadIdx.x, g = threadIdx.y;
  float a = 0.f, b = 0.f;
  if (c < E) {
    int k = g;
    for (; k + 48 < nblk; k += 64) {
      const float a0 = part_dg[(size_t)k * E + c], a1 = part_dg[(size_t)(k + 16) * E + c];
      const float a2 = part_dg[(size_t)(k + 32) * E + c], a3 = part_dg[(size_t)(k + 48) * E + c];
      const float b0 = part_db[(size_t)k * E + c], b1 = part_db[(size_t)(k + 16) * E + c];
      const float b2 = part_db[(size_t)(k + 32) * E + c], b3 = part_db[(size_t)(k + 48) * E + c];
      a += a0; a += a1; a += a2; a += a3;
      b += b0; b += b1; b += b2; b += b3;
    }
    for (; k < nblk; k += 16) { a += part_dg[(size_t)k * E + c]; b += part_db[(size_t)k * E + c]; }
  }
  sa[g][threadIdx.x] = a; sb_[g][threadIdx.x] = b;
  __syncthreads();
  if (g == 0 && c < E) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) { ta += sa[q][threadIdx.x]; tb += sb_[q][threadIdx.x]; }
    dg[c] = ta; db[c] = tb;
  }
}

// ------------------------------------------------------------------ dropout: out = in * multiplier (in place allowed)
static __global__ void k_mul_mask(const float* in, const float* __restrict__ mult, long long n4, float* out) {
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[q], m = reinterpret_cast<const float4*>(mult)[q];
    reinterpret_cast<float4*>(out)[q] = make_float4(a.x * m.x, a.y * m.y, a.z * m.z, a.w * m.w);
  }
}

// ------------------------------------------------------------------ attention (nn.MultiheadAttention, net.py:406-441)
// qkv [B, L, 3E] (q | k | v, head h at columns h*D..), key j is attendable by query i iff j <= i and
// the step of token j is valid (key_padding_mask = ~mask repeated over the 4 tokens of a step, cdt.py:203-205).
// One CTA per batch element, one thread per (head, query/key row).
// [H][L][L] dropout multipliers of one batch element -> shared [H][L][L + 1] (the pad makes both access patterns of
// the kernels -- lanes over query rows, lanes over key columns -- conflict free); 16-byte coalesced global loads
__device__ __forceinline__ void attn_stage_drop(const float* __restrict__ src, int H, int L, float* __restrict__ Pd) {
  const int n = H * L * L;
  if ((L & 3) == 0) {
    for (int e4 = threadIdx.x; e4 < n / 4; e4 += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(src)[e4];
      const int e = e4 * 4, row = e / L, j = e - row * L;
      float* d = Pd + (size_t)row * (L + 1) + j;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int row = e / L, j = e - row * L;
      Pd[(size_t)row * (L + 1) + j] = src[e];
    }
  }
}
template <int D>
// pdrop (optional) [B, H, L, L]: dropout multipliers of the attention weights (nn.MultiheadAttention dropout):
// O = (softmax(S) * pdrop) V -- the normaliser is the un-dropped sum, so lse is unchanged.
static __global__ void k_attn_fwd(const float* __restrict__ qkv, const float* __restrict__ mask, int L, int H,
                                  int tok_per_step, float* __restrict__ out, float* __restrict__ lse,
                                  const float* __restrict__ pdrop) {
  extern __shared__ __align__(16) float sm[];
  // gridDim.y head groups per batch element: HL heads (EL columns) per CTA -- smaller tiles and fewer threads per CTA
  // let several CTAs share an SM (the kernels run at one 125-register, 137 KB CTA per SM otherwise)
  const int E = H * D, b = blockIdx.x, T = L / tok_per_step;
  const int HL = H / gridDim.y, hg = blockIdx.y, EL = HL * D, c0 = hg * EL;
  float* Ks = sm;              // [L][EL]
  float* Vs = sm + L * EL;     // [L][EL]
  float* valid = sm + 2 * L * EL;           // [L] key j attendable
  float* Pd = valid + ((L + 3) & ~3);       // [HL][L][L + 1] dropout multipliers of this batch element (pdrop only)
  const float* base = qkv + (size_t)b * L * 3 * E;
  for (int e4 = threadIdx.x; e4 < L * EL / 4; e4 += blockDim.x) {
    const int e = e4 * 4, j = e / EL, c = e % EL;
    *reinterpret_cast<float4*>(Ks + e) = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * E + E + c0 + c);
    *reinterpret_cast<float4*>(Vs + e) = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * E + 2 * E + c0 + c);
  }
  // the key-padding mask and the dropout multipliers are read in every iteration of the inner loop: staged once
  for (int j = threadIdx.x; j < L; j += blockDim.x) valid[j] = mask[(size_t)b * T + j / tok_per_step];
  if (pdrop) attn_stage_drop(pdrop + ((size_t)b * H + (size_t)hg * HL) * L * L, HL, L, Pd);
  __syncthreads();
  const int h = threadIdx.x / L, i = threadIdx.x % L;   // h: head inside the group
  if (h >= HL) return;
  const int hh = hg * HL + h;
  float q[D], acc[D];
#pragma unroll
  for (int c = 0; c < D; ++c) { q[c] = base[(size_t)i * 3 * E + hh * D + c]; acc[c] = 0.f; }
  const float scale = rsqrtf((float)D);
  const float* __restrict__ drow = pdrop ? Pd + (size_t)(h * L + i) * (L + 1) : nullptr;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j <= i; ++j) {
    if (valid[j] <= 0.f) continue;
    float kj[D], vj[D];
    {
      const float4* __restrict__ kp = reinterpret_cast<const float4*>(Ks + j * EL + h * D);
      const float4* __restrict__ vp = reinterpret_cast<const float4*>(Vs + j * EL + h * D);
#pragma unroll
      for (int q4 = 0; q4 < D / 4; ++q4) {
        const float4 a = kp[q4], v = vp[q4];
        kj[4 * q4] = a.x; kj[4 * q4 + 1] = a.y; kj[4 * q4 + 2] = a.z; kj[4 * q4 + 3] = a.w;
        vj[4 * q4] = v.x; vj[4 * q4 + 1] = v.y; vj[4 * q4 + 2] = v.z; vj[4 * q4 + 3] = v.w;
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) s = fmaf(q[c], kj[c], s);
    s *= scale;
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), p = expf(s - mn);
    l = l * corr + p;
    const float pv = drow ? p * drow[j] : p;
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = acc[c] * corr + pv * vj[c];
    m = mn;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < D; ++c) out[((size_t)b * L + i) * E + hh * D + c] = acc[c] * inv;
  lse[((size_t)b * H + hh) * L + i] = m + logf(l);
}

// row r, head h of a [L][E] shared tile as D/4 float4 (16-byte shared loads: a quarter of the LDS instructions of
// the scalar form, which bound this kernel -- every FMA operand of the inner loops comes from shared memory)
template <int D>
__device__ __forceinline__ void ld_row(const float* __restrict__ tile, int r, int E, int h, float (&v)[D]) {
  const float4* __restrict__ p = reinterpret_cast<const float4*>(tile + r * E + h * D);
#pragma unroll
  for (int q = 0; q < D / 4; ++q) {
    const float4 x = p[q];
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}
template <int D>
static __global__ void k_attn_bwd(const float* __restrict__ qkv, const float* __restrict__ mask, int L, int H,
                                  int tok_per_step, const float* __restrict__ out, const float* __restrict__ dout,
                                  const float* __restrict__ lse, float* __restrict__ dqkv,
                                  const float* __restrict__ pdrop) {
  extern __shared__ __align__(16) float sm[];
  const int E = H * D, b = blockIdx.x, T = L / tok_per_step;
  const int HL = H / gridDim.y, hg = blockIdx.y, EL = HL * D, c0 = hg * EL;   // head group, see k_attn_fwd
  float* Qs = sm;
  float* Ks = sm + L * EL;
  float* Vs = sm + 2 * L * EL;
  float* dOs = sm + 3 * L * EL;
  float* Ls = sm + 4 * L * EL;     // [HL*L] log-sum-exp
  float* Ds = Ls + HL * L;         // [HL*L] rowsum(dO * O)
  float* valid = Ds + HL * L;               // [L] key j attendable
  float* Pd = valid + ((L + 3) & ~3);       // [HL][L][L + 1] dropout multipliers (pdrop only), see k_attn_fwd
  const float* base = qkv + (size_t)b * L * 3 * E;
  for (int j = threadIdx.x; j < L; j += blockDim.x) valid[j] = mask[(size_t)b * T + j / tok_per_step];
  if (pdrop) attn_stage_drop(pdrop + ((size_t)b * H + (size_t)hg * HL) * L * L, HL, L, Pd);
  for (int e4 = threadIdx.x; e4 < L * EL / 4; e4 += blockDim.x) {   // 16-byte global loads / shared stores
    const int e = e4 * 4, j = e / EL, c = e % EL;
    *reinterpret_cast<float4*>(Qs + e) = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * E + c0 + c);
    *reinterpret_cast<float4*>(Ks + e) = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * E + E + c0 + c);
    *reinterpret_cast<float4*>(Vs + e) = *reinterpret_cast<const float4*>(base + (size_t)j * 3 * E + 2 * E + c0 + c);
    *reinterpret_cast<float4*>(dOs + e) = *reinterpret_cast<const float4*>(dout + ((size_t)b * L + j) * E + c0 + c);
  }
  const int h = threadIdx.x / L, i = threadIdx.x % L;   // h: head inside the group
  const bool active = h < HL;
  const int hh = hg * HL + h;
  const float scale = rsqrtf((float)D);
  if (active) {
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c)
      d += dout[((size_t)b * L + i) * E + hh * D + c] * out[((size_t)b * L + i) * E + hh * D + c];
    Ds[h * L + i] = d;
    Ls[h * L + i] = lse[((size_t)b * H + hh) * L + i];
  }
  __syncthreads();
  float* dbase = dqkv + (size_t)b * L * 3 * E;
  if (active) {
    float qi[D], doi[D], ki[D], vi[D];
    ld_row<D>(Qs, i, EL, h, qi);
    ld_row<D>(dOs, i, EL, h, doi);
    ld_row<D>(Ks, i, EL, h, ki);
    ld_row<D>(Vs, i, EL, h, vi);
    // ---- dq for query row i
    float dq[D];
#pragma unroll
    for (int c = 0; c < D; ++c) dq[c] = 0.f;
    const float li = Ls[h * L + i], di = Ds[h * L + i];
    const int LP = L + 1;
    const float* __restrict__ dmat = pdrop ? Pd + (size_t)h * L * LP : nullptr;   // [L][L + 1] of this head (shared)
    for (int j = 0; j <= i; ++j) {
      if (valid[j] <= 0.f) continue;
      float kj[D], vj[D];
      ld_row<D>(Ks, j, EL, h, kj);
      ld_row<D>(Vs, j, EL, h, vj);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        s = fmaf(qi[c], kj[c], s);
        dp = fmaf(doi[c], vj[c], dp);
      }
      const float p = expf(s * scale - li);
      if (dmat) dp *= dmat[i * LP + j];  // d softmax = (dO V^T) * dropout multiplier; rowsum(dP * P) is still dO.O
      const float ds = p * (dp - di) * scale;
#pragma unroll
      for (int c = 0; c < D; ++c) dq[c] = fmaf(ds, kj[c], dq[c]);
    }
#pragma unroll
    for (int c = 0; c < D; c += 4)
      *reinterpret_cast<float4*>(dbase + (size_t)i * 3 * E + hh * D + c) = make_float4(dq[c], dq[c + 1], dq[c + 2], dq[c + 3]);
    // ---- dk, dv for key row j = i
    const int j = i;
    float dk[D], dv[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    if (valid[j] > 0.f) {
      // every lane walks the SAME query row r in the same iteration (rows above the lane's key are skipped by a
      // predicate, not by a later loop start): the row loads are then warp broadcasts.  Starting each lane at its own
      // r = j made the lanes of a warp read 32 different rows, 128 floats apart -- all in the same banks (ncu: 101 M
      // bank conflicts on 16.9 M shared loads, the kernel's whole run time).
      for (int r = 0; r < L; ++r) {
        if (r < j) continue;
        float qr[D], dor[D];
        ld_row<D>(Qs, r, EL, h, qr);
        ld_row<D>(dOs, r, EL, h, dor);
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          s = fmaf(qr[c], ki[c], s);
          dp = fmaf(dor[c], vi[c], dp);
        }
        const float p = expf(s * scale - Ls[h * L + r]);
        const float mrj = dmat ? dmat[r * LP + j] : 1.f;
        const float ds = p * (dp * mrj - Ds[h * L + r]) * scale;
        const float pv = p * mrj;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          dk[c] = fmaf(ds, qr[c], dk[c]);
          dv[c] = fmaf(pv, dor[c], dv[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < D; c += 4) {
      *reinterpret_cast<float4*>(dbase + (size_t)j * 3 * E + E + hh * D + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
      *reinterpret_cast<float4*>(dbase + (size_t)j * 3 * E + 2 * E + hh * D + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
    }
  }
}

// ------------------------------------------------------------------ head losses (cdt.py:357-394, 402-418)
// mh [BT, 2a] = (mu | log_std) on the state tokens; ah [BT, 2+o] = (cost logits | next-state prediction) on
// the action tokens.  Single CTA.  Writes d mh, d ah, the 9 logged stats, and steps log_temperature (Adam).
__device__ __forceinline__ double block_sum_d(double v, double* sh /*[33]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double x = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (l == 0) sh[32] = x;
  }
  __syncthreads();
  return sh[32];
}
static __global__ void k_cdt_loss(const float* __restrict__ mh, const float* __restrict__ ah,
                                  const float* __restrict__ actions, const float* __restrict__ costs,
                                  const float* __restrict__ states, const float* __restrict__ mask, int B, int T, int a,
                                  int o, float w_cost, float w_state, float target_entropy, float base_lr, int warmup,
                                  int group, DevState* ds, float* __restrict__ dmh, float* __restrict__ dah,
                                  float* __restrict__ stat, int phase, double* __restrict__ sums, int world) {
  // phase 1 (any grid): every CTA writes the six partial sums of its rows to sums[6 * blockIdx.x ..]; k_cdt_fold adds
  // them in block order (and, data parallel, the caller all-reduces the result -- the masked means cdt.py:358-359,385
  // are means over the GLOBAL batch, so their denominators are global counts and the gradients need no 1/world);
  // phase 2 (any grid) reads the six totals from sums[0..5], writes the gradients of its rows, CTA 0 the stats and the
  // temperature step.  phase 0: single CTA, everything in one launch (kept for reference).
  __shared__ double sh[33];
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  const int BT = B * T, wa = 2 + o;
  const double HALF_LOG_2PI = 0.91893853320467274178;
  double s_valid = 0, s_lp = 0, s_ent = 0, s_nll = 0, s_corr = 0, s_sl = 0;
  if (phase != 2)
  for (int r = gtid; r < BT; r += gstride) {
    const float m = mask[r];
    if (m > 0.f) {
      s_valid += 1.0;
      for (int j = 0; j < a; ++j) {
        const float mu = mh[(size_t)r * 2 * a + j], ls = mh[(size_t)r * 2 * a + a + j];
        const float d = actions[(size_t)r * a + j] - mu;
        const float var = expf(2.f * ls);
        s_lp += (double)(-(d * d) / (2.f * var) - ls) - HALF_LOG_2PI;
        s_ent += 0.5 + HALF_LOG_2PI + (double)ls;
      }
    }
    const float z0 = ah[(size_t)r * wa], z1 = ah[(size_t)r * wa + 1];
    const float mx = fmaxf(z0, z1);
    const float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
    const int c = (int)costs[r];
    s_nll += (double)((lse - (c ? z1 : z0)) * m);
    const int pred = z1 > z0 ? 1 : 0;
    s_corr += (pred == c) ? (double)m : 0.0;
    if ((r % T) < T - 1) {
      double e = 0;
      for (int j = 0; j < o; ++j) {
        const float d = ah[(size_t)r * wa + 2 + j] - states[(size_t)(r + 1) * o + j];
        e += (double)d * d;
      }
      s_sl += e * (double)m;
    }
  }
  double t_valid, t_lp, t_ent, t_nll, t_corr, t_sl;
  if (phase == 2) {
    t_valid = sums[0]; t_lp = sums[1]; t_ent = sums[2]; t_nll = sums[3]; t_corr = sums[4]; t_sl = sums[5];
  } else {
    t_valid = block_sum_d(s_valid, sh); t_lp = block_sum_d(s_lp, sh); t_ent = block_sum_d(s_ent, sh);
    t_nll = block_sum_d(s_nll, sh); t_corr = block_sum_d(s_corr, sh); t_sl = block_sum_d(s_sl, sh);
    if (phase == 1) {
      if (threadIdx.x == 0) {
        double* o_ = sums + 6 * blockIdx.x;
        o_[0] = t_valid; o_[1] = t_lp; o_[2] = t_ent; o_[3] = t_nll; o_[4] = t_corr; o_[5] = t_sl;
      }
      return;
    }
  }
  const double gBT = (double)BT * world, gB = (double)B * world;
  const double n_valid = t_valid;
  const double ll = t_lp / (n_valid * a);
  const double ent = t_ent / (n_valid * a);
  const double cost_loss = t_nll / gBT;
  const double acc = t_corr / n_valid;
  const double state_loss = (T > 1) ? t_sl / (gB * (T - 1) * o) : 0.0;
  const double log_temp = phase == 2 ? sums[6 * gridDim.x] : ds->log_temperature;
  const double temp = exp(log_temp);
  const double act_loss = -(ll + temp * ent);
  const float ca = (float)(1.0 / (n_valid * a));
  const float tf = (float)temp;
  const float cc = (float)(w_cost / gBT);
  const float cs = (T > 1) ? (float)(w_state * 2.0 / (gB * (T - 1) * o)) : 0.f;
  for (int r = gtid; r < BT; r += gstride) {
    const float m = mask[r] > 0.f ? 1.f : 0.f;
    for (int j = 0; j < a; ++j) {
      const float mu = mh[(size_t)r * 2 * a + j], ls = mh[(size_t)r * 2 * a + a + j];
      const float d = actions[(size_t)r * a + j] - mu;
      const float iv = expf(-2.f * ls);
      dmh[(size_t)r * 2 * a + j] = -m * ca * d * iv;
      dmh[(size_t)r * 2 * a + a + j] = -m * ca * ((d * d * iv - 1.f) + tf);
    }
    const float z0 = ah[(size_t)r * wa], z1 = ah[(size_t)r * wa + 1];
    const float mx = fmaxf(z0, z1);
    const float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
    const float inv = 1.f / (e0 + e1);
    const int c = (int)costs[r];
    const float mm = mask[r];
    dah[(size_t)r * wa] = cc * mm * (e0 * inv - (c == 0 ? 1.f : 0.f));
    dah[(size_t)r * wa + 1] = cc * mm * (e1 * inv - (c == 1 ? 1.f : 0.f));
    for (int j = 0; j < o; ++j) {
      float g = 0.f;
      if ((r % T) < T - 1) g = cs * mm * (ah[(size_t)r * wa + 2 + j] - states[(size_t)(r + 1) * o + j]);
      dah[(size_t)r * wa + 2 + j] = g;
    }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    stat[0] = (float)(-ll);          // nll
    stat[1] = (float)ent;            // ent
    stat[2] = (float)temp;           // ent_reg
    stat[3] = (float)(act_loss + w_cost * cost_loss + w_state * state_loss);  // all_loss
    stat[4] = (float)act_loss;
    stat[5] = (float)cost_loss;
    stat[6] = (float)acc;
    stat[7] = (float)state_loss;
    const int t = ds->adam_t[group];                     // already incremented for this step
    const double f = (double)(t + 1) / (double)warmup;
    stat[8] = (float)((double)base_lr * (f < 1.0 ? f : 1.0));   // scheduler.get_last_lr() after scheduler.step()
    // temperature: Adam(lr 1e-4) on temp * (entropy - target).detach()   (cdt.py:402-407)
    const double g = temp * (ent - (double)target_entropy);
    const int tt = t;                                     // temperature optimiser steps in lock-step
    const double m1 = ds->temp_m + (1.0 - 0.9) * (g - ds->temp_m);
    const double v1 = 0.999 * ds->temp_v + (1.0 - 0.999) * g * g;
    const double denom = sqrt(v1) / sqrt(1.0 - pow(0.999, (double)tt)) + 1e-8;
    ds->log_temperature = log_temp - (1e-4 / (1.0 - pow(0.9, (double)tt))) * m1 / denom;
    ds->temp_m = m1;
    ds->temp_v = v1;
  }
}

// per-CTA partial sums of k_cdt_loss phase 1 -> the six totals, added in block order (deterministic)
// (slot 6 * nblk keeps the temperature of THIS step: CTA 0 of phase 2 steps ds->log_temperature while other CTAs of the
// same launch still need the old value)
static __global__ void k_cdt_fold(double* __restrict__ sums, int nblk, const DevState* ds) {
  const int k = threadIdx.x;
  if (k == 6) sums[6 * nblk] = ds->log_temperature;
  if (k >= 6) return;
  double t = 0;
  for (int b = 0; b < nblk; ++b) t += sums[6 * b + k];
  __syncwarp(0x3f);
  sums[k] = t;
}

// ------------------------------------------------------------------ clip_grad_norm_ (cdt.py:399)
static __global__ void k_sumsq_partial(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ float sh[33];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += g[i] * g[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
static __global__ void k_clip_coef(const float* __restrict__ part, int n, float max_norm, float* coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)part[i];
  const float total = (float)sqrt(s);
  const float c = max_norm / (total + 1e-6f);
  coef[0] = c < 1.f ? c : 1.f;
}

}  // namespace osrl
