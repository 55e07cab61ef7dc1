// Step programs of BC (bc.py:45-52,103-109) and BCQ-Lag (bcql.py:122-234, 283-306).
#include "engine.h"

namespace osrl {



// ====================================================================== BC
void build_bc(Engine& e) {
  const osrl_config& c = e.plan.cfg;
  Program& p = e.body;
  const int B = e.B, o = c.obs_dim, a = c.act_dim;
  const MlpLay& m = e.plan.mlp_actor;
  float* u = e.ws((size_t)B * a);
  float* dpre = e.ws((size_t)B * a);
  std::vector<float*> h;
  GemmTask last = mlp_fwd_hidden(e, p, e.P, m, e.b_obs, o, B, ACT_RELU, h, u, a);
  last.act = ACT_TANH;            // net.py:83-85: act_limit * tanh(.)
  last.scale = c.max_action;
  emit_gemm(e, p, {last});
  float* stat = e.stats;
  const float lim = c.max_action, iw = e.inv_world();
  const float* act = e.b_act;
  KOP(p, e, 0.0, (k_bc_loss<<<1, 1024, 0, s>>>(u, act, B * a, lim, dpre, stat, iw)));
  mlp_bwd(e, p, e.P, e.G, m, e.b_obs, o, B, ACT_RELU, h, dpre);
  const Group& g = e.plan.groups[e.plan.g_actor];
  emit_allreduce(e, p, e.G + g.begin, g.end - g.begin, false, DP_GRAD);
  emit_adam(e, p, e.plan.g_actor, g.begin, g.end, false);
}

// ====================================================================== BCQ-Lag
// phase 0: the whole step into e.body.  Pipelined graphs (engine.cu) build the step a second time in two halves
// that can run concurrently: phase 1 (into e.pa, reading the NEXT minibatch e.nb_*) = the VAE update and, right
// after it, the two VAE decodes of that step (target pass on 2R rows, actor pass on B rows): they depend only on
// the VAE, the minibatch and noise, so their outputs are handed over (Engine::handoff) to phase 2 = everything else
// (into e.pm, reading the current minibatch), which then starts at the actor_old perturbation.
void build_bcql(Engine& e, int phase) {
  const osrl_config& c = e.plan.cfg;
  const Plan& pl = e.plan;
  Program& p = phase == 1 ? e.pa : (phase == 2 ? e.pm : e.body);
  const bool do_vae = phase != 2, do_rest = phase != 1;
  const float* bobs = phase == 1 ? e.nb_obs : e.b_obs;
  const float* bact = phase == 1 ? e.nb_act : e.b_act;
  const float* Wvae = phase == 2 ? e.Psnap : e.P;
  const int B = e.B, S = c.sample_action_num, R = B * S, o = c.obs_dim, a = c.act_dim, L = 2 * a, V = c.vae_hidden;
  const int in = o + a, din = o + L;
  const float lim = c.max_action, philim = (float)((double)c.phi * (double)c.max_action);
  const float iw = e.inv_world();
  const EnsLay& cr = pl.critic;
  const EnsLay& cc = pl.cost_critic;
  const MlpLay& act = pl.mlp_actor;
  const int nh = (int)cr.h.size();
  const float *n_vae = e.noise_buf[0], *n_zc = e.noise_buf[1], *n_zcc = e.noise_buf[2], *n_za = e.noise_buf[3];

  // ---------------- derived inputs (concat / repeat_interleave / clamp), one launch
  float* sa = e.ws((size_t)B * in);          // [obs | act]
  float* v_dec_in = e.ws((size_t)B * din);   // [obs | z]           (VAE update)
  float* t_dec_in = e.ws((size_t)2 * R * din);  // rows 0..R: critic target, R..2R: cost target
  float* t_ain = e.ws((size_t)2 * R * in);   // [next_obs rep | a_vae]
  float* t_qin = e.ws((size_t)2 * R * in);   // [next_obs rep | a_target]
  float* p_dec_in = e.ws((size_t)B * din);   // actor step
  float* p_ain = e.ws((size_t)B * in);
  float* p_qin = e.ws((size_t)B * in);
  // pipelined halves: decode outputs travel through (next, cur) buffer pairs copied before the graph forks
  float *t_av = nullptr, *p_av = nullptr;   // [2R, a] / [B, a] VAE actions: written in phase 1, read in phase 2
  const bool dside = e.decode_side;         // (off: the decodes stay in phase 2 and read the VAE weight snapshot)
  if (!dside) {
  } else if (phase == 1) {
    t_av = e.ws((size_t)2 * R * a); p_av = e.ws((size_t)B * a);
    e.handoff.push_back({t_av, e.ws((size_t)2 * R * a), (size_t)2 * R * a * sizeof(float)});
    e.handoff.push_back({p_av, e.ws((size_t)B * a), (size_t)B * a * sizeof(float)});
  } else if (phase == 2) {
    OSRL_REQUIRE(e.handoff.size() == 2, "phase 1 must be built first");
    t_av = e.handoff[0].cur; p_av = e.handoff[1].cur;
  }
  const bool do_dec_inputs = dside ? phase != 2 : phase != 1;   // [obs | z] decoder inputs: wherever the decodes run
  const float* dec_nobs = phase == 1 ? e.nb_nobs : e.b_nobs;
  {
    std::vector<CopyTask> ct;
    ct.push_back(copy_cols(sa, in, 0, bobs, o, 0, B, o));     // (both halves use [obs | act]: VAE loss, online critics)
    ct.push_back(copy_cols(sa, in, o, bact, a, 0, B, a));
    if (do_vae) ct.push_back(copy_cols(v_dec_in, din, 0, bobs, o, 0, B, o));
    if (do_dec_inputs) {
      ct.push_back(copy_cols(t_dec_in, din, 0, dec_nobs, o, 0, 2 * R, o, S, B));   // repeat_interleave (bcql.py:138)
      CopyTask z1 = copy_cols(t_dec_in, din, o, n_zc, L, 0, R, L);                 // z.clamp(-0.5, 0.5) (net.py:334)
      z1.clamp = 1; z1.lo = -0.5f; z1.hi = 0.5f;
      CopyTask z2 = copy_cols(t_dec_in + (size_t)R * din, din, o, n_zcc, L, 0, R, L);
      z2.clamp = 1; z2.lo = -0.5f; z2.hi = 0.5f;
      CopyTask z3 = copy_cols(p_dec_in, din, o, n_za, L, 0, B, L);
      z3.clamp = 1; z3.lo = -0.5f; z3.hi = 0.5f;
      ct.push_back(z1); ct.push_back(z2); ct.push_back(z3);
      ct.push_back(copy_cols(p_dec_in, din, 0, bobs, o, 0, B, o));
    }
    if (do_rest) {
      ct.push_back(copy_cols(t_ain, in, 0, e.b_nobs, o, 0, 2 * R, o, S, B));
      ct.push_back(copy_cols(t_qin, in, 0, e.b_nobs, o, 0, 2 * R, o, S, B));
      ct.push_back(copy_cols(p_ain, in, 0, e.b_obs, o, 0, B, o));
      ct.push_back(copy_cols(p_qin, in, 0, e.b_obs, o, 0, B, o));
      if (phase == 2 && dside) {   // VAE actions decoded by the other half one replay earlier
        ct.push_back(copy_cols(t_ain, in, o, t_av, a, 0, 2 * R, a));
        ct.push_back(copy_cols(p_ain, in, o, p_av, a, 0, B, a));
      }
    }
    emit_copy(e, p, ct);
  }

  // ---------------- 1. VAE update (bcql.py:122-132)
  if (do_vae) emit_vae_update(e, p, sa, v_dec_in, n_vae, bact, 0);
  if (phase == 1 && !dside) return;
  if (phase == 1) {   // the step's two decodes with the freshly updated VAE (bcql.py:141,164,189), handed to phase 2
    float* th1 = e.ws((size_t)2 * R * V); float* th2 = e.ws((size_t)2 * R * V);
    emit_vae_decode(e, p, e.P, t_dec_in, 2 * R, th1, th2, t_av, a, 0, true);
    float* ph1 = e.ws((size_t)B * V); float* ph2 = e.ws((size_t)B * V);
    emit_vae_decode(e, p, e.P, p_dec_in, B, ph1, ph2, p_av, a, 0);
    return;
  }

  // ---------------- 2+3. critic and cost-critic updates (bcql.py:134-179), merged launch-by-launch
  // target actions on 2R rows: current VAE decode -> actor_old perturbation
  {
    if (phase == 0 || !dside) {
      float* th1 = e.ws((size_t)2 * R * V); float* th2 = e.ws((size_t)2 * R * V);
      emit_vae_decode(e, p, Wvae, t_dec_in, 2 * R, th1, th2, t_ain + o, in, 0, true);
    }
    std::vector<float*> ah;
    GemmTask last = mlp_fwd_hidden(e, p, e.T, act, t_ain, in, 2 * R, ACT_TANH, ah, t_qin + o, in, true);
    last.act = ACT_TANH; last.scale = philim;           // net.py:61: phi*act_limit*pi(.)
    last.resid = t_ain + o; last.ldr = in;              // + act
    last.clamp = 1; last.lo = -lim; last.hi = lim;      // net.py:62
    emit_gemm(e, p, {last});
  }
  EnsBuf tq = ens_alloc(e, cr, R), tqc = ens_alloc(e, cc, R);        // targets (critic_old / cost_critic_old)
  EnsBuf oq = ens_alloc(e, cr, B), oqc = ens_alloc(e, cc, B);        // online nets on (s, a)
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
  }
  {
    const Group& g1 = pl.groups[pl.g_critic];
    const Group& g2 = pl.groups[pl.g_cost];
    OSRL_REQUIRE(g1.end == g2.begin, "critic groups must be adjacent");
    emit_allreduce(e, p, e.G + g1.begin, g2.end - g1.begin, false, DP_GRAD);
    // same lr and step count for both groups -> one fused Adam+Polyak launch over the union
    emit_adam(e, p, pl.g_critic, g1.begin, g2.end, true);
  }

  // ---------------- 4. actor update (bcql.py:181-216)
  float* pt = e.ws((size_t)B * a);  // tanh(l3) of the perturbation net
  std::vector<float*> pah;
  {
    if (phase == 0 || !dside) {
      float* ph1 = e.ws((size_t)B * V); float* ph2 = e.ws((size_t)B * V);
      emit_vae_decode(e, p, Wvae, p_dec_in, B, ph1, ph2, p_ain + o, in, 0);
    }
    GemmTask last = mlp_fwd_hidden(e, p, e.P, act, p_ain, in, B, ACT_TANH, pah, p_qin + o, in);
    last.act = ACT_TANH; last.scale = philim;
    last.aux = pt; last.ldaux = a;
    last.resid = p_ain + o; last.ldr = in;
    last.clamp = 1; last.lo = -lim; last.hi = lim;
    emit_gemm(e, p, {last});
  }
  EnsBuf pq = ens_alloc(e, cr, B), pqc = ens_alloc(e, cc, B);
  {
    std::vector<Stage> st(nh + 1);
    ens_fwd(st, cr, e.P, p_qin, in, B, pq);
    ens_fwd(st, cc, e.P, p_qin, in, B, pqc);
    emit_stages(e, p, st);
  }
  float* dpq = e.ws((size_t)B * cr.n); float* dpqc = e.ws((size_t)B * cc.n);
  {
    const float thres = (float)pl.qc_thres, kp = c.pid_kp, ki = c.pid_ki, kd = c.pid_kd;
    const float *pqv = pq.q, *pqcv = pqc.q;
    const int nq = cr.n, nqc = cc.n;
    DevState* ds = e.ds;
    float* st3 = e.stats + 3;
    const float* gmean = nullptr;
    if (e.world > 1) {  // PID error is a mean over the GLOBAL batch (net.py:380): 4-byte all-reduce before the backward
      float* part = e.ws(4);
      KOP(p, e, 0.0, (k_rowmin_mean<<<1, 1024, 0, s>>>(pqcv, nqc, B, thres, iw, part)));
      emit_allreduce(e, p, part, 1, false, DP_SCALAR);
      gmean = part;
    }
    KOP(p, e, 0.0, (k_bcql_actor_loss<<<1, 1024, 0, s>>>(pqv, nq, pqcv, nqc, B, thres, kp, ki, kd, ds, dpq, dpqc, st3, iw,
                                                   gmean)));
  }
  float* da_q = e.ws((size_t)B * a); float* da_qc = e.ws((size_t)B * a);
  {
    EnsBuf gq = ens_alloc(e, cr, B), gqc = ens_alloc(e, cc, B);
    std::vector<Stage> st(nh + 1);
    ens_bwd(st, cr, e.P, nullptr, p_qin, in, B, pq, gq, dpq, da_q, a, o, a);
    ens_bwd(st, cc, e.P, nullptr, p_qin, in, B, pqc, gqc, dpqc, da_qc, a, o, a);
    emit_stages(e, p, st);
  }
  float* dpre = e.ws((size_t)B * a);
  {
    const float* av = p_ain + o;
    KOP(p, e, 0.0, (k_perturb_bwd<<<(B * a + 255) / 256, 256, 0, s>>>(da_q, da_qc, a, pt, av, in, B, a, philim, lim, dpre)));
  }
  mlp_bwd(e, p, e.P, e.G, act, p_ain, in, B, ACT_TANH, pah, dpre);
  {
    const Group& g = pl.groups[pl.g_actor];
    emit_allreduce(e, p, e.G + g.begin, g.end - g.begin, false, DP_GRAD);
    emit_adam(e, p, pl.g_actor, g.begin, g.end, true);
  }
}

}  // namespace osrl
