// Step program of CDT (cdt.py:166-265 forward, 343-418 train_one_step; TransformerBlock net.py:391-441).
//
// Token layout: N = B*4T tokens, row (b*T + t)*4 + tau with tau = (rtg, ctg, state, action) (cdt.py:198-200).
// All projections are tasks of the shared tensor-core GEMM (token sub-sets are addressed with leading
// dimension 4E); LayerNorm, attention, the head losses and the gradient clip are the kernels of
// cdt_kernels.cuh.  Residual-stream gradients are accumulated in place in `dres`.
#include "cdt_kernels.cuh"
#include "engine.h"
#include <map>

namespace osrl {

static Lin row_lin(int64_t w, int64_t b, int in, int out) {
  Lin l;
  l.w = w; l.b = b; l.in = in; l.out = out;
  return l;
}

template <int D>
static void set_attn_attr(int smem_bwd) {
  OSRL_CUDA(cudaFuncSetAttribute(k_attn_bwd<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bwd));
  OSRL_CUDA(cudaFuncSetAttribute(k_attn_fwd<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bwd));
}

static void emit_ln_fwd(Engine& e, Program& p, const float* x, int64_t gw, int64_t gb, float* y, float* mean,
                        float* rstd, int rows, int E) {
  const float* g = e.P + gw;
  const float* b = e.P + gb;
  KOP(p, e, 8.0 * rows * E, (k_ln_fwd<<<(rows + 7) / 8, 256, 0, s>>>(x, g, b, y, mean, rstd, rows, E)));
}
static void emit_ln_bwd(Engine& e, Program& p, const float* dy, const float* x, const float* mean, const float* rstd,
                        int64_t gw, int64_t gb, float* dx, int accumulate, int rows, int E) {
  const int nblk = std::min(592, (rows + 7) / 8);
  float* pg = e.ws((size_t)nblk * E);
  float* pb = e.ws((size_t)nblk * E);
  const float* g = e.P + gw;
  float* dg = e.G + gw;
  float* db = e.G + gb;
  const int per = E / 32;
  Engine* ep = &e;
  p.add("k_ln_bwd", 16.0 * rows * E, 0.0, true, [=](cudaStream_t s) {
    switch (per) {
      case 1: k_ln_bwd<1><<<nblk, 256, 0, s>>>(dy, x, mean, rstd, g, dx, accumulate, pg, pb, rows, E); break;
      case 2: k_ln_bwd<2><<<nblk, 256, 0, s>>>(dy, x, mean, rstd, g, dx, accumulate, pg, pb, rows, E); break;
      case 4: k_ln_bwd<4><<<nblk, 256, 0, s>>>(dy, x, mean, rstd, g, dx, accumulate, pg, pb, rows, E); break;
      case 8: k_ln_bwd<8><<<nblk, 256, 0, s>>>(dy, x, mean, rstd, g, dx, accumulate, pg, pb, rows, E); break;
      default: k_ln_bwd<16><<<nblk, 256, 0, s>>>(dy, x, mean, rstd, g, dx, accumulate, pg, pb, rows, E); break;
    }
    ep->launches++;
  });
  KOP(p, e, 0.0, (k_ln_param_reduce<<<(E + 31) / 32, dim3(32, 16), 0, s>>>(pg, pb, nblk, E, dg, db)));
}

void build_cdt(Engine& e) {
  const osrl_config& c = e.plan.cfg;
  const CdtLay& L_ = e.plan.cdt;
  Program& p = e.body;
  const int B = e.B, T = c.seq_len, BT = B * T, E = c.embedding_dim, H = c.num_heads, D = E / H;
  const int Lq = 4 * T, N = B * Lq, o = c.obs_dim, a = c.act_dim, NL = c.num_layers;
  OSRL_REQUIRE(E == 32 || E == 64 || E == 128 || E == 256 || E == 512, "embedding_dim must be 32/64/128/256/512");
  // two head groups per batch element when that does not strand lanes (H/2 * Lq a multiple of 32: the default 8 heads
  // x 40 tokens -> 160 threads): CTAs of 5 warps and 69 KB instead of 10 warps and 137 KB, three per SM instead of one
  const int attn_groups = (H % 2 == 0 && ((H / 2) * Lq) % 32 == 0) ? 2 : 1;
  const int HLg = H / attn_groups, ELg = HLg * D;
  const int attn_threads = (HLg * Lq + 31) / 32 * 32;
  OSRL_REQUIRE(attn_threads <= 1024, "num_heads * 4 * seq_len must be <= 1024");
  // + key validity [Lq] and, with attention dropout, this CTA's multipliers [HLg][Lq][Lq + 1]
  const int attn_extra = ((Lq + 3) & ~3) + (c.attention_dropout > 0.f ? HLg * Lq * (Lq + 1) : 0);
  const int smem_fwd = (2 * Lq * ELg + attn_extra) * (int)sizeof(float);
  const int smem_bwd = (4 * Lq * ELg + 2 * HLg * Lq + attn_extra) * (int)sizeof(float);
  const dim3 attn_grid((unsigned)B, (unsigned)attn_groups);
  OSRL_REQUIRE(smem_bwd <= 227 * 1024, "sequence too long for the single-CTA attention kernel");
  if (D == 8) set_attn_attr<8>(smem_bwd);
  else if (D == 16) set_attn_attr<16>(smem_bwd);
  else set_attn_attr<32>(smem_bwd);

  // dropout multipliers (noise slots "drop_*", plan.cu): null where the probability is 0
  std::map<std::string, const float*> drop;
  for (size_t i = 0; i < e.plan.noise.size(); ++i) drop[e.plan.noise[i].first] = e.noise_buf[i];
  auto drop_of = [&](const std::string& name) -> const float* {
    auto it = drop.find(name);
    return it == drop.end() ? nullptr : it->second;
  };
  auto emit_mul = [&](const float* in, const float* mult, float* outp) {   // outp = in * mult over the N x E stream
    const long long n4 = (long long)N * E / 4;
    KOP(p, e, 12.0 * N * E, (k_mul_mask<<<1184, 256, 0, s>>>(in, mult, n4, outp)));
  };

  // ---------------- forward
  float* te = e.ws((size_t)BT * E);
  float* ctg_t = e.ws(BT);
  {
    const long long* ts = e.s_ts;
    const float* ctg = e.s_ctg;
    const float* table = e.P + L_.te;
    const int rows = L_.te_rows;
    KOP(p, e, 8.0 * BT * E, (k_cdt_prep<<<std::min((BT * E + 255) / 256, 1184), 256, 0, s>>>(ts, ctg, BT, E, table, rows, te, ctg_t)));
  }
  float* x0 = e.ws((size_t)N * E);
  {  // token embeddings + timestep embedding (cdt.py:178-195), written interleaved (ldc = 4E)
    std::vector<GemmTask> ts;
    auto emb = [&](const float* X, int ldx, const Lin& l, int tau) {
      GemmTask t = task_fwd(X, ldx, BT, e.P, l, x0 + (size_t)tau * E, 4 * E, ACT_NONE);
      t.resid = te; t.ldr = E;
      ts.push_back(t);
    };
    emb(e.s_returns, 1, L_.return_emb, 0);
    emb(ctg_t, 1, L_.cost_emb, 1);
    emb(e.s_states, o, L_.state_emb, 2);
    emb(e.s_actions, a, L_.action_emb, 3);
    emit_gemm(e, p, ts);
  }
  struct Saved { float *x_in, *h1, *m1, *r1, *qkv, *att, *lse, *x_mid, *h2, *m2, *r2, *z, *g; };
  std::vector<Saved> sv(NL);
  float* x = e.ws((size_t)N * E);
  float* mean0 = e.ws(N); float* rstd0 = e.ws(N);
  emit_ln_fwd(e, p, x0, L_.emb_norm_w, L_.emb_norm_b, x, mean0, rstd0, N, E);       // emb_norm (cdt.py:221)
  const float* d_emb = drop_of("drop_emb");
  if (d_emb) emit_mul(x, d_emb, x);                                                  // emb_drop (cdt.py:222)
  const float* mask = e.s_mask;
  std::vector<const float*> d_attn(NL), d_ra(NL), d_rb(NL);
  for (int i = 0; i < NL; ++i) {
    d_attn[i] = drop_of("drop_attn" + std::to_string(i));
    d_ra[i] = drop_of("drop_res" + std::to_string(i) + "a");
    d_rb[i] = drop_of("drop_res" + std::to_string(i) + "b");
  }
  for (int i = 0; i < NL; ++i) {
    const CdtLay::Blk& b = L_.blocks[i];
    Saved& s_ = sv[i];
    s_.x_in = x;
    s_.h1 = e.ws((size_t)N * E); s_.m1 = e.ws(N); s_.r1 = e.ws(N);
    s_.qkv = e.ws((size_t)N * 3 * E); s_.att = e.ws((size_t)N * E); s_.lse = e.ws((size_t)B * H * Lq);
    s_.x_mid = e.ws((size_t)N * E);
    s_.h2 = e.ws((size_t)N * E); s_.m2 = e.ws(N); s_.r2 = e.ws(N);
    s_.z = e.ws((size_t)N * 4 * E); s_.g = e.ws((size_t)N * 4 * E);
    float* x_out = e.ws((size_t)N * E);
    emit_ln_fwd(e, p, s_.x_in, b.n1w, b.n1b, s_.h1, s_.m1, s_.r1, N, E);
    emit_gemm(e, p, {task_fwd(s_.h1, E, N, e.P, b.in_proj, s_.qkv, 3 * E, ACT_NONE)});
    {
      const float* qkv = s_.qkv; float* att = s_.att; float* lse = s_.lse;
      const float* pd = d_attn[i];
      Engine* ep = &e;
      p.add("k_attn_fwd", 16.0 * N * E, 4.0 * B * H * Lq * Lq * D, true, [=](cudaStream_t s) {
        if (D == 8) k_attn_fwd<8><<<attn_grid, attn_threads, smem_fwd, s>>>(qkv, mask, Lq, H, 4, att, lse, pd);
        else if (D == 16) k_attn_fwd<16><<<attn_grid, attn_threads, smem_fwd, s>>>(qkv, mask, Lq, H, 4, att, lse, pd);
        else k_attn_fwd<32><<<attn_grid, attn_threads, smem_fwd, s>>>(qkv, mask, Lq, H, 4, att, lse, pd);
        ep->launches++;
      });
    }
    {  // x_mid = x_in + drop(out_proj(att))      (net.py:438-439)
      GemmTask t = task_fwd(s_.att, E, N, e.P, b.out_proj, s_.x_mid, E, ACT_NONE);
      t.resid = s_.x_in; t.ldr = E;
      t.mmask = d_ra[i]; t.ldmm = E;
      emit_gemm(e, p, {t});
    }
    emit_ln_fwd(e, p, s_.x_mid, b.n2w, b.n2b, s_.h2, s_.m2, s_.r2, N, E);
    {  // g = GELU(fc1(h2)), z keeps the pre-activation for the backward
      GemmTask t = task_fwd(s_.h2, E, N, e.P, b.fc1, s_.g, 4 * E, ACT_GELU);
      t.aux = s_.z; t.ldaux = 4 * E;
      emit_gemm(e, p, {t});
    }
    {  // x_out = x_mid + drop(fc2(g))            (net.py:440, mlp's trailing nn.Dropout :414)
      GemmTask t = task_fwd(s_.g, 4 * E, N, e.P, b.fc2, x_out, E, ACT_NONE);
      t.resid = s_.x_mid; t.ldr = E;
      t.mmask = d_rb[i]; t.ldmm = E;
      emit_gemm(e, p, {t});
    }
    x = x_out;
  }
  float* x_last = x;
  float* out = e.ws((size_t)N * E);
  float* meanL = e.ws(N); float* rstdL = e.ws(N);
  emit_ln_fwd(e, p, x_last, L_.out_norm_w, L_.out_norm_b, out, meanL, rstdL, N, E);   // out_norm (cdt.py:228)
  // heads: action distribution from the STATE token, cost / next-state predictions from the ACTION token
  float* mh = e.ws((size_t)BT * 2 * a);
  float* ah = e.ws((size_t)BT * (2 + o));
  emit_gemm(e, p, {task_fwd(out + 2 * E, 4 * E, BT, e.P, L_.act_head, mh, 2 * a, ACT_NONE),
                   task_fwd(out + 3 * E, 4 * E, BT, e.P, L_.aux_head, ah, 2 + o, ACT_NONE)});

  // ---------------- losses (cdt.py:357-394) + temperature step (:402-407)
  float* dmh = e.ws((size_t)BT * 2 * a);
  float* dah = e.ws((size_t)BT * (2 + o));
  {
    const float *actions = e.s_actions, *costs = e.s_costs, *states = e.s_states;
    const float wc = c.loss_cost_weight, wsw = c.loss_state_weight, tent = c.target_entropy, lr = c.learning_rate;
    const int warm = c.lr_warmup_steps > 0 ? c.lr_warmup_steps : 1, grp = e.plan.g_cdt;
    DevState* ds = e.ds;
    float* stat = e.stats;
    const int world = e.world;
    // three launches: partial sums on `nb` CTAs, a fold (+ all-reduce of the six totals under data parallelism: the masked
    // means are over the global batch), gradients on `nb` CTAs.  (One CTA for everything took 0.5 ms at B = 2048.)
    const int nb = std::max(1, std::min(148, (BT + 255) / 256));
    double* sums = (double*)e.ws((size_t)2 * 6 * nb + 16);
    KOP(p, e, 0.0, (k_cdt_loss<<<nb, 256, 0, s>>>(mh, ah, actions, costs, states, mask, B, T, a, o, wc, wsw, tent, lr,
                                                  warm, grp, ds, dmh, dah, stat, 1, sums, world)));
    KOP(p, e, 0.0, (k_cdt_fold<<<1, 32, 0, s>>>(sums, nb, ds)));
    if (world > 1) emit_allreduce(e, p, (float*)sums, 6, /*f64=*/true);
    KOP(p, e, 0.0, (k_cdt_loss<<<nb, 256, 0, s>>>(mh, ah, actions, costs, states, mask, B, T, a, o, wc, wsw, tent, lr,
                                                  warm, grp, ds, dmh, dah, stat, 2, sums, world)));
  }

  // ---------------- backward
  float* dout = e.ws((size_t)N * E);
  {
    const size_t bytes = (size_t)N * E * sizeof(float);
    p.add("memset", 0.0, 0.0, false, [=](cudaStream_t s) { cudaMemsetAsync(dout, 0, bytes, s); });
  }
  emit_gemm(e, p, {task_wgrad(dmh, 2 * a, out + 2 * E, 4 * E, BT, e.G, L_.act_head),
                   task_wgrad(dah, 2 + o, out + 3 * E, 4 * E, BT, e.G, L_.aux_head),
                   task_dgrad(dmh, 2 * a, BT, e.P, L_.act_head, dout + 2 * E, 4 * E, nullptr, 0, 0),
                   task_dgrad(dah, 2 + o, BT, e.P, L_.aux_head, dout + 3 * E, 4 * E, nullptr, 0, 0)});
  float* dres = e.ws((size_t)N * E);   // gradient wrt the residual stream at the current depth
  emit_ln_bwd(e, p, dout, x_last, meanL, rstdL, L_.out_norm_w, L_.out_norm_b, dres, 0, N, E);
  float* dg = e.ws((size_t)N * 4 * E);   // d z (pre-GELU)
  float* dh = e.ws((size_t)N * E);       // d LayerNorm output
  float* datt = e.ws((size_t)N * E);
  float* dqkv = e.ws((size_t)N * 3 * E);
  float* ddrop = (c.residual_dropout > 0.f) ? e.ws((size_t)N * E) : nullptr;   // dres * dropout multiplier
  for (int i = NL - 1; i >= 0; --i) {
    const CdtLay::Blk& b = L_.blocks[i];
    const Saved& s_ = sv[i];
    // MLP branch: x_out = x_mid + drop(fc2(GELU(fc1(LN2(x_mid)))))
    const float* dy2 = dres;
    if (d_rb[i]) { emit_mul(dres, d_rb[i], ddrop); dy2 = ddrop; }
    emit_gemm(e, p, {task_wgrad(dy2, E, s_.g, 4 * E, N, e.G, b.fc2),
                     task_dgrad(dy2, E, N, e.P, b.fc2, dg, 4 * E, s_.z, 4 * E, ACT_GELU)});
    emit_gemm(e, p, {task_wgrad(dg, 4 * E, s_.h2, E, N, e.G, b.fc1),
                     task_dgrad(dg, 4 * E, N, e.P, b.fc1, dh, E, nullptr, 0, 0)});
    emit_ln_bwd(e, p, dh, s_.x_mid, s_.m2, s_.r2, b.n2w, b.n2b, dres, 1, N, E);
    // attention branch: x_mid = x_in + drop(out_proj(attn(LN1(x_in))))
    const float* dy1 = dres;
    if (d_ra[i]) { emit_mul(dres, d_ra[i], ddrop); dy1 = ddrop; }
    emit_gemm(e, p, {task_wgrad(dy1, E, s_.att, E, N, e.G, b.out_proj),
                     task_dgrad(dy1, E, N, e.P, b.out_proj, datt, E, nullptr, 0, 0)});
    {
      const float *qkv = s_.qkv, *att = s_.att, *lse = s_.lse;
      const float* pd = d_attn[i];
      Engine* ep = &e;
      p.add("k_attn_bwd", 36.0 * N * E, 10.0 * B * H * Lq * Lq * D, true, [=](cudaStream_t s) {
        if (D == 8) k_attn_bwd<8><<<attn_grid, attn_threads, smem_bwd, s>>>(qkv, mask, Lq, H, 4, att, datt, lse, dqkv, pd);
        else if (D == 16) k_attn_bwd<16><<<attn_grid, attn_threads, smem_bwd, s>>>(qkv, mask, Lq, H, 4, att, datt, lse, dqkv, pd);
        else k_attn_bwd<32><<<attn_grid, attn_threads, smem_bwd, s>>>(qkv, mask, Lq, H, 4, att, datt, lse, dqkv, pd);
        ep->launches++;
      });
    }
    emit_gemm(e, p, {task_wgrad(dqkv, 3 * E, s_.h1, E, N, e.G, b.in_proj),
                     task_dgrad(dqkv, 3 * E, N, e.P, b.in_proj, dh, E, nullptr, 0, 0)});
    emit_ln_bwd(e, p, dh, s_.x_in, s_.m1, s_.r1, b.n1w, b.n1b, dres, 1, N, E);
  }
  float* dx0 = e.ws((size_t)N * E);
  if (d_emb) emit_mul(dres, d_emb, dres);   // back through emb_drop
  emit_ln_bwd(e, p, dres, x0, mean0, rstd0, L_.emb_norm_w, L_.emb_norm_b, dx0, 0, N, E);
  // embedding gradients: token tau of every step is row (bt*4 + tau) -> leading dimension 4E
  emit_gemm(e, p, {task_wgrad(dx0 + 0 * E, 4 * E, e.s_returns, 1, BT, e.G, L_.return_emb),
                   task_wgrad(dx0 + 1 * E, 4 * E, ctg_t, 1, BT, e.G, L_.cost_emb),
                   task_wgrad(dx0 + 2 * E, 4 * E, e.s_states, o, BT, e.G, L_.state_emb),
                   task_wgrad(dx0 + 3 * E, 4 * E, e.s_actions, a, BT, e.G, L_.action_emb)});
  {
    float* gte = e.G + L_.te;
    const size_t bytes = (size_t)L_.te_rows * E * sizeof(float);
    const long long* ts = e.s_ts;
    const int rows = L_.te_rows;
    p.add("memset", 0.0, 0.0, false, [=](cudaStream_t s) { cudaMemsetAsync(gte, 0, bytes, s); });
    KOP(p, e, 20.0 * BT * E, (k_cdt_te_scatter<<<std::min((BT * E + 255) / 256, 1184), 256, 0, s>>>(ts, dx0, BT, E, rows, gte)));
  }
  // ---------------- clip_grad_norm_ (cdt.py:399) + AdamW with warm-up (cdt.py:321-330)
  const Group& g = e.plan.groups[e.plan.g_cdt];
  emit_allreduce(e, p, e.G + g.begin, g.end - g.begin);   // data parallel: global gradient before the global-norm clip
  float* coef = nullptr;
  if (c.clip_grad > 0.f) {
    const int nb = 296;
    float* part = e.ws(nb);
    coef = e.ws(4);
    const float* G0 = e.G + g.begin;
    const long long n = g.end - g.begin;
    const float mx = c.clip_grad;
    KOP(p, e, 4.0 * n, (k_sumsq_partial<<<nb, 256, 0, s>>>(G0, n, part)));
    KOP(p, e, 0.0, (k_clip_coef<<<1, 32, 0, s>>>(part, nb, mx, coef)));
  }
  emit_adam(e, p, e.plan.g_cdt, g.begin, g.end, false, coef);
}

}  // namespace osrl
