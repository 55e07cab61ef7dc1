// Step program of COptiDICE (osrl/algorithms/coptidice.py:125-227, 285-287).
//
// One step = (1) nu / chi / tau / lambda: both critic ensembles are evaluated on observations and next_observations
// in ONE pass over the 2B stacked rows (so their weight gradients are one reduction), k_cop_main turns the 2 x 2B x n
// outputs into every loss, the output gradients (routed to each row's arg-min member: `predict` is a min over the
// ensemble, net.py:235-238), the batch-softmax chi weights and the two scalar Adam steps; then the backward passes
// and one Adam launch over nu_network + chi_network (adjacent groups, same lr).  (2) policy extraction: nu again with
// the UPDATED weights, the actor on noisy observations, the weighted Gaussian log-likelihood and its backward.
#include "engine.h"

namespace osrl {

void build_coptidice(Engine& e) {
  const osrl_config& c = e.plan.cfg;
  const Plan& pl = e.plan;
  Program& p = e.body;
  const int B = e.B, o = c.obs_dim, a = c.act_dim;
  const EnsLay& nu = pl.critic;
  const EnsLay& chi = pl.cost_critic;
  MlpLay actor = pl.sq_actor.trunk;
  actor.L.push_back(pl.sq_actor.heads);
  const int nh = (int)nu.h.size();
  const float *n_obs = e.noise_buf[0], *n_act = e.noise_buf[1];
  const bool use_chi = c.cost_ub_epsilon != 0.f;

  float* xcat = e.ws((size_t)2 * B * o);   // [observations ; next_observations]
  emit_copy(e, p, {copy_cols(xcat, o, 0, e.b_obs, o, 0, B, o), copy_cols(xcat + (size_t)B * o, o, 0, e.b_nobs, o, 0, B, o)});
  EnsBuf fnu = ens_alloc(e, nu, 2 * B), fchi = ens_alloc(e, chi, 2 * B);
  {
    std::vector<Stage> st(nh + 1);
    ens_fwd(st, nu, e.P, xcat, o, 2 * B, fnu);
    if (use_chi) ens_fwd(st, chi, e.P, xcat, o, 2 * B, fchi);
    emit_stages(e, p, st);
  }
  float* dq_nu = e.ws((size_t)2 * B * nu.n);
  float* dq_chi = e.ws((size_t)2 * B * chi.n);
  {
    CopArgs ca;
    ca.q_nu = fnu.q; ca.q_chi = fchi.q; ca.n_nu = nu.n; ca.n_chi = chi.n; ca.B = B;
    ca.rew = e.b_rew; ca.cost = e.b_cost; ca.done = e.b_done; ca.init = e.b_init;
    ca.gamma = c.gamma; ca.alpha = c.alpha; ca.eps = c.cost_ub_epsilon; ca.p0 = c.init_state_propotion;
    ca.thres = (float)pl.qc_thres; ca.scalar_lr = c.scalar_lr; ca.ftype = c.f_type;
    ca.dq_nu = dq_nu; ca.dq_chi = dq_chi;
    ca.e_buf = e.ws(B); ca.w_buf = e.ws(B); ca.ell_buf = e.ws(B);
    ca.stats = e.stats;
    DevState* ds = e.ds;
    KOP(p, e, 0.0, (k_cop_main<<<1, 1024, 0, s>>>(ca, ds)));
  }
  {
    EnsBuf gnu = ens_alloc(e, nu, 2 * B), gchi = ens_alloc(e, chi, 2 * B);
    std::vector<Stage> st(nh + 1);
    ens_bwd(st, nu, e.P, e.G, xcat, o, 2 * B, fnu, gnu, dq_nu, nullptr, 0, 0, 0);
    if (use_chi) ens_bwd(st, chi, e.P, e.G, xcat, o, 2 * B, fchi, gchi, dq_chi, nullptr, 0, 0, 0);
    emit_stages(e, p, st);
  }
  {
    const Group& g1 = pl.groups[pl.g_critic];
    const Group& g2 = pl.groups[pl.g_cost];
    OSRL_REQUIRE(g1.end == g2.begin, "nu / chi groups must be adjacent");
    if (use_chi) emit_adam(e, p, pl.g_critic, g1.begin, g2.end, false);   // same lr, same step count: one launch
    else emit_adam(e, p, pl.g_critic, g1.begin, g1.end, false);
  }
  // ---------------- policy extraction (coptidice.py:200-212)
  EnsBuf fnu2 = ens_alloc(e, nu, 2 * B);
  float* xin = e.ws((size_t)B * o);
  float* xact = e.ws((size_t)B * a);
  {
    const float *obs = e.b_obs, *act = e.b_act, *osd = e.cop_obs_std, *asd = e.cop_act_std;
    KOP(p, e, 0.0, (k_cop_noise<<<(B * o + 255) / 256, 256, 0, s>>>(obs, n_obs, osd, B, o, xin)));
    KOP(p, e, 0.0, (k_cop_noise<<<(B * a + 255) / 256, 256, 0, s>>>(act, n_act, asd, B, a, xact)));
  }
  float* mh = e.ws((size_t)B * 2 * a);
  std::vector<float*> h;
  {
    std::vector<Stage> st(nh + 1);
    ens_fwd(st, nu, e.P, xcat, o, 2 * B, fnu2, /*nograd=*/true);
    emit_stages(e, p, st);
    emit_gemm(e, p, {mlp_fwd_hidden(e, p, e.P, actor, xin, o, B, ACT_RELU, h, mh, 2 * a)});
  }
  float* dmh = e.ws((size_t)B * 2 * a);
  {
    const float *q2 = fnu2.q, *rew = e.b_rew, *cost = e.b_cost, *done = e.b_done;
    const int nn = nu.n, ft = c.f_type;
    const float gm = c.gamma, al = c.alpha;
    float* st7 = e.stats + 7;
    const DevState* ds = e.ds;
    KOP(p, e, 0.0, (k_cop_actor_loss<<<1, 1024, 0, s>>>(q2, nn, B, a, rew, cost, done, gm, al, ft, mh, xact, dmh, st7, ds)));
  }
  mlp_bwd(e, p, e.P, e.G, actor, xin, o, B, ACT_RELU, h, dmh);
  const Group& ga = pl.groups[pl.g_actor];
  emit_adam(e, p, pl.g_actor, ga.begin, ga.end, false);
}

}  // namespace osrl
