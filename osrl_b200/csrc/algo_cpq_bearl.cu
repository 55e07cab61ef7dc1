// Step programs of CPQ (cpq.py:125-230, 294-313) and BEAR-Lag (bearl.py:144-335, 389-412).
#include "engine.h"

namespace osrl {



static void emit_squash(Engine& e, Program& p, const std::vector<SquashTask>& tasks) {
  SquashTask* d = e.upload(tasks);
  int mx = 1;
  for (auto& t : tasks) mx = std::max(mx, t.rows * t.a);
  const int bx = std::min((mx + 255) / 256, 148 * 4), ny = (int)tasks.size();
  KOP(p, e, 0.0, (k_squash_tasks<<<dim3(bx, ny), 256, 0, s>>>(d)));
}
static SquashTask squash(const float* mh, int row_div, int row_mod, const float* eps, float* dst, int ldd, int rows,
                         int a, int do_tanh, float scale, float* u_out = nullptr) {
  SquashTask t;
  t.mh = mh; t.row_div = row_div; t.row_mod = row_mod; t.eps = eps; t.dst = dst; t.ldd = ldd; t.u_out = u_out;
  t.rows = rows; t.a = a; t.do_tanh = do_tanh; t.scale = scale;
  return t;
}
// SquashedGaussianMLPActor as one MLP: ReLU trunk layers + the stacked [mu | log_std] head layer
static MlpLay sq_as_mlp(const SqActorLay& s) {
  MlpLay m = s.trunk;
  m.L.push_back(s.heads);
  return m;
}

// ====================================================================== CPQ
// phase: see build_bcql (0 = whole step, 1 = VAE update alone on the NEXT minibatch, 2 = the rest, VAE weights
// read from the snapshot)
void build_cpq(Engine& e, int phase) {
  const osrl_config& c = e.plan.cfg;
  const Plan& pl = e.plan;
  Program& p = phase == 1 ? e.pa : (phase == 2 ? e.pm : e.body);
  const bool do_vae = phase != 2, do_rest = phase != 1;
  const float* bobs = phase == 1 ? e.nb_obs : e.b_obs;
  const float* bact = phase == 1 ? e.nb_act : e.b_act;
  const float* Wvae = phase == 2 ? e.Psnap : e.P;
  const int B = e.B, S = c.sample_action_num, SB = S * B, o = c.obs_dim, a = c.act_dim, L = 2 * a, V = c.vae_hidden;
  const int in = o + a, din = o + L;
  const float lim = c.max_action, iw = e.inv_world();
  const EnsLay& cr = pl.critic;
  const EnsLay& cc = pl.cost_critic;
  const MlpLay actor = sq_as_mlp(pl.sq_actor);
  const int nh = (int)cr.h.size();
  const float *n_vae = e.noise_buf[0], *n_pc = e.noise_buf[1], *n_pcc = e.noise_buf[2], *n_ood = e.noise_buf[3],
              *n_pa = e.noise_buf[4];

  float* sa = e.ws((size_t)B * in);
  float* v_dec_in = e.ws((size_t)B * din);
  float* obs2 = e.ws((size_t)2 * B * o);        // [next_obs ; obs] rows for one actor pass
  float* qin1 = e.ws((size_t)B * in);           // [next_obs | a'(eps critic)]
  float* qin2 = e.ws((size_t)B * in);           // [next_obs | a'(eps cost)]
  float* qin_ood = e.ws((size_t)SB * in);       // [obs tiled S-major | raw sampled action]
  float* p_qin = e.ws((size_t)B * in);          // actor step
  {
    std::vector<CopyTask> ct;
    ct.push_back(copy_cols(sa, in, 0, bobs, o, 0, B, o));
    ct.push_back(copy_cols(sa, in, o, bact, a, 0, B, a));
    if (do_vae) ct.push_back(copy_cols(v_dec_in, din, 0, bobs, o, 0, B, o));
    if (do_rest) {
      ct.push_back(copy_cols(obs2, o, 0, e.b_nobs, o, 0, B, o));
      ct.push_back(copy_cols(obs2 + (size_t)B * o, o, 0, e.b_obs, o, 0, B, o));
      ct.push_back(copy_cols(qin1, in, 0, e.b_nobs, o, 0, B, o));
      ct.push_back(copy_cols(qin2, in, 0, e.b_nobs, o, 0, B, o));
      ct.push_back(copy_cols(qin_ood, in, 0, e.b_obs, o, 0, SB, o, 1, B));  // torch.tile (cpq.py:169-173)
      ct.push_back(copy_cols(p_qin, in, 0, e.b_obs, o, 0, B, o));
    }
    emit_copy(e, p, ct);
  }
  // ---- 1. VAE
  if (do_vae) emit_vae_update(e, p, sa, v_dec_in, n_vae, bact, 0);
  if (!do_rest) return;

  // ---- 2+3. critic and cost critic
  float* mh2 = e.ws((size_t)2 * B * 2 * a);  // actor (mu|log_std) on [next_obs ; obs]
  {
    std::vector<float*> h;
    emit_gemm(e, p, {mlp_fwd_hidden(e, p, e.P, actor, obs2, o, 2 * B, ACT_RELU, h, mh2, 2 * a)});
    emit_squash(e, p, {squash(mh2, 1, 1 << 30, n_pc, qin1 + o, in, B, a, 1, lim),      // cpq.py:141
                       squash(mh2, 1, 1 << 30, n_pcc, qin2 + o, in, B, a, 1, lim),     // cpq.py:159
                       // Normal(mu,std).sample([S]) on obs: pre-tanh, unscaled, S-major (cpq.py:164-168)
                       squash(mh2 + (size_t)B * 2 * a, 1, B, n_ood, qin_ood + o, in, SB, a, 0, 1.f)});
  }
  EnsBuf tq = ens_alloc(e, cr, B), tqc1 = ens_alloc(e, cc, B), tqc2 = ens_alloc(e, cc, B), tood = ens_alloc(e, cc, SB);
  EnsBuf oq = ens_alloc(e, cr, B), oqc = ens_alloc(e, cc, B);
  float* ood_h1 = e.ws((size_t)SB * V); float* ood_h2 = e.ws((size_t)SB * V); float* ood_ml = e.ws((size_t)SB * 2 * L);
  {
    std::vector<Stage> st(std::max(nh + 1, 3));
    ens_fwd(st, cr, e.T, qin1, in, B, tq, true);
    ens_fwd(st, cc, e.T, qin1, in, B, tqc1, true);
    ens_fwd(st, cc, e.T, qin2, in, B, tqc2, true);
    ens_fwd(st, cc, e.T, qin_ood, in, SB, tood, true);
    ens_fwd(st, cr, e.P, sa, in, B, oq);
    ens_fwd(st, cc, e.P, sa, in, B, oqc);
    // VAE encoder on the OOD samples (cpq.py:178; decoder output is discarded there)
    st[0].tasks.push_back(task_fwd(qin_ood, in, SB, Wvae, pl.vae.e1, ood_h1, V, ACT_RELU));
    st[1].tasks.push_back(task_fwd(ood_h1, V, SB, Wvae, pl.vae.e2, ood_h2, V, ACT_RELU));
    st[2].tasks.push_back(task_fwd(ood_h2, V, SB, Wvae, pl.vae.heads, ood_ml, 2 * L, ACT_NONE));
    emit_stages(e, p, st);
  }
  float* yq = e.ws(B); float* yc = e.ws(B);
  float* kl = e.ws(SB); float* qcmin = e.ws(SB);
  float* dq = e.ws((size_t)B * cr.n); float* dqc = e.ws((size_t)B * cc.n);
  float* ood_term = e.ws(4);
  {
    const float gm = c.gamma, qth = (float)pl.q_thres, qcth = (float)pl.qc_thres, alr = c.alpha_lr;
    const float *rew = e.b_rew, *cost = e.b_cost, *done = e.b_done;
    const float *tqv = tq.q, *t1 = tqc1.q, *t2 = tqc2.q, *tov = tood.q, *oqv = oq.q, *oqcv = oqc.q;
    const int nq = cr.n, nqc = cc.n;
    DevState* ds = e.ds;
    float *st1 = e.stats + 1, *st2 = e.stats + 2, *st3 = e.stats + 3;
    KOP(p, e, 0.0, (k_cpq_backup<<<(B + 127) / 128, 128, 0, s>>>(tqv, nq, t1, t2, nqc, B, gm, qth, rew, cost, done, yq, yc)));
    KOP(p, e, 0.0, (k_cpq_kl_rows<<<(SB + 127) / 128, 128, 0, s>>>(ood_ml, L, tov, nqc, SB, kl, qcmin)));
    KOP(p, e, 0.0, (k_cpq_ood<<<1, 1024, 0, s>>>(kl, qcmin, S, B, 0.75f, qcth, alr, ds, ood_term, st3)));
    KOP(p, e, 0.0, (k_critic_loss<<<1, 1024, 0, s>>>(oqv, yq, B, nq, dq, st1, iw, 0.f, nullptr)));
    // loss_cost_critic = MSE - exp(log_alpha) * (qc_ood.mean() - qc_thres)  (cpq.py:186-187)
    KOP(p, e, 0.0, (k_critic_loss<<<1, 1024, 0, s>>>(oqcv, yc, B, nqc, dqc, st2, iw, -1.f, ood_term)));
  }
  {
    EnsBuf gq = ens_alloc(e, cr, B), gqc = ens_alloc(e, cc, B);
    std::vector<Stage> st(nh + 1);
    ens_bwd(st, cr, e.P, e.G, sa, in, B, oq, gq, dq, nullptr, 0, 0, 0);
    ens_bwd(st, cc, e.P, e.G, sa, in, B, oqc, gqc, dqc, nullptr, 0, 0, 0);
    emit_stages(e, p, st);
    const Group& g1 = pl.groups[pl.g_critic];
    const Group& g2 = pl.groups[pl.g_cost];
    OSRL_REQUIRE(g1.end == g2.begin, "critic groups must be adjacent");
    emit_allreduce(e, p, e.G + g1.begin, g2.end - g1.begin, false, DP_GRAD);
    emit_adam(e, p, pl.g_critic, g1.begin, g2.end, true);
  }

  // ---- 4. actor (cpq.py:203-222)
  float* mh = e.ws((size_t)B * 2 * a); float* pu = e.ws((size_t)B * a);
  std::vector<float*> ah;
  emit_gemm(e, p, {mlp_fwd_hidden(e, p, e.P, actor, e.b_obs, o, B, ACT_RELU, ah, mh, 2 * a)});
  emit_squash(e, p, {squash(mh, 1, 1 << 30, n_pa, p_qin + o, in, B, a, 1, lim, pu)});
  EnsBuf pq = ens_alloc(e, cr, B), pqc = ens_alloc(e, cc, B);
  {
    std::vector<Stage> st(nh + 1);
    ens_fwd(st, cr, e.P, p_qin, in, B, pq);
    ens_fwd(st, cc, e.P, p_qin, in, B, pqc);
    emit_stages(e, p, st);
  }
  float* dpq = e.ws((size_t)B * cr.n);
  {
    const float qth = (float)pl.q_thres;
    const float *pqv = pq.q, *pqcv = pqc.q;
    const int nq = cr.n, nqc = cc.n;
    float* st4 = e.stats + 4;
    KOP(p, e, 0.0, (k_cpq_actor_loss<<<1, 1024, 0, s>>>(pqv, nq, pqcv, nqc, B, qth, dpq, st4, iw)));
  }
  float* da = e.ws((size_t)B * a);
  {
    EnsBuf gq = ens_alloc(e, cr, B);
    std::vector<Stage> st(nh + 1);
    ens_bwd(st, cr, e.P, nullptr, p_qin, in, B, pq, gq, dpq, da, a, o, a);
    emit_stages(e, p, st);
  }
  float* dmh = e.ws((size_t)B * 2 * a);
  KOP(p, e, 0.0, (k_squash_bwd<<<(B * a + 255) / 256, 256, 0, s>>>(da, a, pu, mh, n_pa, B, a, lim, dmh)));
  mlp_bwd(e, p, e.P, e.G, actor, e.b_obs, o, B, ACT_RELU, ah, dmh);
  {
    const Group& g = pl.groups[pl.g_actor];
    emit_allreduce(e, p, e.G + g.begin, g.end - g.begin, false, DP_GRAD);
    emit_adam(e, p, pl.g_actor, g.begin, g.end, true);
  }
}

// ====================================================================== BEAR-Lag
void build_bearl(Engine& e, int phase) {
  const osrl_config& c = e.plan.cfg;
  const Plan& pl = e.plan;
  Program& p = phase == 1 ? e.pa : (phase == 2 ? e.pm : e.body);
  const bool do_vae = phase != 2, do_rest = phase != 1;
  const float* bobs = phase == 1 ? e.nb_obs : e.b_obs;
  const float* bact = phase == 1 ? e.nb_act : e.b_act;
  const float* Wvae = phase == 2 ? e.Psnap : e.P;
  const int B = e.B, S = c.sample_action_num, R = B * S, N = c.num_samples_mmd_match, BN = B * N;
  const int o = c.obs_dim, a = c.act_dim, L = 2 * a, V = c.vae_hidden;
  const int in = o + a, din = o + L;
  OSRL_REQUIRE(a <= OSRL_MMD_MAXA, "act_dim too large for the MMD kernel");
  const float iw = e.inv_world();
  const EnsLay& cr = pl.critic;
  const EnsLay& cc = pl.cost_critic;
  const MlpLay actor = sq_as_mlp(pl.sq_actor);
  const int nh = (int)cr.h.size();
  const float *n_vae = e.noise_buf[0], *n_pc = e.noise_buf[1], *n_pcc = e.noise_buf[2], *n_z = e.noise_buf[3],
              *n_pa = e.noise_buf[4];

  float* sa = e.ws((size_t)B * in);
  float* v_dec_in = e.ws((size_t)B * din);
  float* t_qin = e.ws((size_t)2 * R * in);   // rows 0..R critic target, R..2R cost target
  float* m_dec_in = e.ws((size_t)BN * din);  // decode_multiple input
  float* p_qin = e.ws((size_t)B * in);
  {
    std::vector<CopyTask> ct;
    ct.push_back(copy_cols(sa, in, 0, bobs, o, 0, B, o));
    ct.push_back(copy_cols(sa, in, o, bact, a, 0, B, a));
    if (do_vae) ct.push_back(copy_cols(v_dec_in, din, 0, bobs, o, 0, B, o));
    if (do_rest) {
      ct.push_back(copy_cols(t_qin, in, 0, e.b_nobs, o, 0, 2 * R, o, S, B));   // repeat_interleave (bearl.py:160)
      ct.push_back(copy_cols(m_dec_in, din, 0, e.b_obs, o, 0, BN, o, N, B));   // net.py:348-351
      CopyTask z = copy_cols(m_dec_in, din, o, n_z, L, 0, BN, L);
      z.clamp = 1; z.lo = -0.5f; z.hi = 0.5f;
      ct.push_back(z);
      ct.push_back(copy_cols(p_qin, in, 0, e.b_obs, o, 0, B, o));
    }
    emit_copy(e, p, ct);
  }
  // ---- 1. VAE
  if (do_vae) emit_vae_update(e, p, sa, v_dec_in, n_vae, bact, 0);
  if (!do_rest) return;

  // ---- 2+3. critics: actor_old's (mu, std) depend on next_obs only -> one B-row pass, S samples each
  float* mht = e.ws((size_t)B * 2 * a);
  {
    std::vector<float*> h;
    emit_gemm(e, p, {mlp_fwd_hidden(e, p, e.T, actor, e.b_nobs, o, B, ACT_RELU, h, mht, 2 * a)});
    // bearl.py:163,188: tanh(u), not scaled by max_action
    emit_squash(e, p, {squash(mht, S, B, n_pc, t_qin + o, in, R, a, 1, 1.f),
                       squash(mht, S, B, n_pcc, t_qin + (size_t)R * in + o, in, R, a, 1, 1.f)});
  }
  EnsBuf tq = ens_alloc(e, cr, R), tqc = ens_alloc(e, cc, R);
  EnsBuf oq = ens_alloc(e, cr, B), oqc = ens_alloc(e, cc, B);
  {
    std::vector<Stage> st(nh + 1);
    ens_fwd(st, cr, e.T, t_qin, in, R, tq, true);
    ens_fwd(st, cc, e.T, t_qin + (size_t)R * in, in, R, tqc, true);
    ens_fwd(st, cr, e.P, sa, in, B, oq);
    ens_fwd(st, cc, e.P, sa, in, B, oqc);
    emit_stages(e, p, st);
  }
  float* y_q = e.ws(B); float* y_qc = e.ws(B);
  float* dq = e.ws((size_t)B * cr.n); float* dqc = e.ws((size_t)B * cc.n);
  {
    const float lm = c.lmbda, gm = c.gamma;
    const float *rew = e.b_rew, *cost = e.b_cost, *done = e.b_done;
    const float *tqv = tq.q, *tqcv = tqc.q, *oqv = oq.q, *oqcv = oqc.q;
    const int nq = cr.n, nqc = cc.n;
    float* st1 = e.stats + 1; float* st2 = e.stats + 2;
    const BackupLossArgs a0{tqv, nq, rew, 1, y_q, oqv, dq, st1}, a1{tqcv, nqc, cost, 0, y_qc, oqcv, dqc, st2};
    KOP(p, e, 0.0, (k_backup_critic_loss2<<<2, 1024, 0, s>>>(a0, a1, done, B, S, lm, gm, iw)));
  }
  {
    EnsBuf gq = ens_alloc(e, cr, B), gqc = ens_alloc(e, cc, B);
    std::vector<Stage> st(nh + 1);
    ens_bwd(st, cr, e.P, e.G, sa, in, B, oq, gq, dq, nullptr, 0, 0, 0);
    ens_bwd(st, cc, e.P, e.G, sa, in, B, oqc, gqc, dqc, nullptr, 0, 0, 0);
    emit_stages(e, p, st);
    const Group& g1 = pl.groups[pl.g_critic];
    const Group& g2 = pl.groups[pl.g_cost];
    OSRL_REQUIRE(g1.end == g2.begin, "critic groups must be adjacent");
    emit_allreduce(e, p, e.G + g1.begin, g2.end - g1.begin, false, DP_GRAD);
    emit_adam(e, p, pl.g_critic, g1.begin, g2.end, true);
  }

  // ---- 4. actor (bearl.py:208-280)
  float* raw_vae = e.ws((size_t)BN * a);  // decode_multiple pre-tanh output (net.py:353)
  {
    float* h1 = e.ws((size_t)BN * V); float* h2 = e.ws((size_t)BN * V);
    emit_vae_decode(e, p, Wvae, m_dec_in, BN, h1, h2, raw_vae, a, 1);
  }
  float* mh = e.ws((size_t)B * 2 * a);
  float* samp = e.ws((size_t)BN * a);   // tanh(u)
  float* raw = e.ws((size_t)BN * a);    // u
  std::vector<float*> ah;
  emit_gemm(e, p, {mlp_fwd_hidden(e, p, e.P, actor, e.b_obs, o, B, ACT_RELU, ah, mh, 2 * a)});
  emit_squash(e, p, {squash(mh, N, B, n_pa, samp, a, BN, a, 1, 1.f, raw)});
  // critics see actor_samples[:, 0, :] (bearl.py:240-242)
  { std::vector<CopyTask> ct{copy_cols(p_qin, in, o, samp, N * a, 0, B, a)}; emit_copy(e, p, ct); }
  EnsBuf pq = ens_alloc(e, cr, B), pqc = ens_alloc(e, cc, B);
  {
    std::vector<Stage> st(nh + 1);
    ens_fwd(st, cr, e.P, p_qin, in, B, pq);
    ens_fwd(st, cc, e.P, p_qin, in, B, pqc);
    emit_stages(e, p, st);
  }
  float* mmdv = e.ws(B); float* mmd_coef = e.ws(4);
  float* dpq = e.ws((size_t)B * cr.n); float* dpqc = e.ws((size_t)B * cc.n);
  const float sigma = c.mmd_sigma;
  const int lap = c.mmd_kernel;
  {
    const float thres = (float)pl.qc_thres, kp = c.pid_kp, ki = c.pid_ki, kd = c.pid_kd, mth = c.target_mmd_thresh,
                alr = c.alpha_lr;
    const int start = c.start_update_policy_step;
    const float *pqv = pq.q, *pqcv = pqc.q;
    const int nq = cr.n, nqc = cc.n;
    DevState* ds = e.ds;
    float* st3 = e.stats + 3;
    KOP(p, e, 0.0, (k_mmd_fwd<<<(B + 63) / 64, 64, 0, s>>>(raw_vae, raw, B, N, a, sigma, lap, mmdv)));
    const float* gmean = nullptr;
    if (e.world > 1) {  // PID error (net.py:380) and mean MMD (bearl.py:261) are means over the GLOBAL batch
      float* part = e.ws(4);
      KOP(p, e, 0.0, (k_rowmin_mean<<<1, 1024, 0, s>>>(pqcv, nqc, B, thres, iw, part)));
      KOP(p, e, 0.0, (k_mean_sub<<<1, 1024, 0, s>>>(mmdv, B, mth, iw, part + 1)));
      emit_allreduce(e, p, part, 2, false, DP_SCALAR);
      gmean = part;
    }
    KOP(p, e, 0.0, (k_bear_actor_loss<<<1, 1024, 0, s>>>(pqv, nq, pqcv, nqc, mmdv, B, thres, kp, ki, kd, mth, alr, start, ds,
                                                   dpq, dpqc, st3, mmd_coef, iw, gmean)));
  }
  float* da_q = e.ws((size_t)B * a); float* da_qc = e.ws((size_t)B * a);
  {
    EnsBuf gq = ens_alloc(e, cr, B), gqc = ens_alloc(e, cc, B);
    std::vector<Stage> st(nh + 1);
    ens_bwd(st, cr, e.P, nullptr, p_qin, in, B, pq, gq, dpq, da_q, a, o, a);
    ens_bwd(st, cc, e.P, nullptr, p_qin, in, B, pqc, gqc, dpqc, da_qc, a, o, a);
    emit_stages(e, p, st);
  }
  float* dmh = e.ws((size_t)B * 2 * a);
  KOP(p, e, 0.0, (k_bear_actor_bwd<<<(B + 63) / 64, 64, 0, s>>>(raw_vae, raw, mmdv, mmd_coef, da_q, da_qc, a, mh, n_pa, B, N,
                                                        a, sigma, lap, dmh)));
  mlp_bwd(e, p, e.P, e.G, actor, e.b_obs, o, B, ACT_RELU, ah, dmh);
  {
    const Group& g = pl.groups[pl.g_actor];
    emit_allreduce(e, p, e.G + g.begin, g.end - g.begin, false, DP_GRAD);
    emit_adam(e, p, pl.g_actor, g.begin, g.end, true);
  }
}

}  // namespace osrl
