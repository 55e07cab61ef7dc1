// Reusable program blocks shared by BCQ-Lag / CPQ / BEAR-Lag: VAE update, VAE decode,
// plain-MLP forward/backward.
#include "engine.h"

namespace osrl {

static void flush(Engine& e, Program& p, std::vector<GemmTask> tasks) { emit_gemm(e, p, tasks); }

// VAE.decode trunk (net.py:337-339): dec_in [rows, o+L] -> relu d1 -> relu d2 -> d3.
// mode 0: out = act_lim * tanh(d3)    mode 1: out = raw d3 (BEAR decode_multiple, net.py:353)
void emit_vae_decode(Engine& e, Program& p, const float* W, const float* dec_in, int rows, float* h1, float* h2,
                     float* out, int ldout, int mode, bool nograd) {
  const VaeLay& v = e.plan.vae;
  const int V = v.d1.out, ldin = v.d1.in;
  if (fz_mlp_ok(v.d1, v.d2, v.d3) && !fz_unfuse_first(rows, nograd)) {   // the whole decoder in one fused launch
    FzTask t = fz_fwd3(dec_in, ldin, rows, W, v.d1, v.d2, v.d3, ACT_RELU, nograd ? nullptr : h1, V, nograd ? nullptr : h2,
                       V, out, ldout);
    if (mode == 0) { t.ract = ACT_TANH; t.rscale = e.plan.cfg.max_action; }
    emit_fz(e, p, {t});
    return;
  }
  GemmTask first = task_fwd(dec_in, ldin, rows, W, v.d1, h1, V, ACT_RELU);
  if (nograd) { first.pk_gcols = V; first.c_dead = 1; }
  flush(e, p, {first});
  flush(e, p, {task_fwd(h1, V, rows, W, v.d2, h2, V, ACT_RELU)});
  if (mode == 0) flush(e, p, {task_fwd(h2, V, rows, W, v.d3, out, ldout, ACT_TANH, e.plan.cfg.max_action)});
  else flush(e, p, {task_fwd(h2, V, rows, W, v.d3, out, ldout, ACT_NONE)});
}

// vae_loss + backward + Adam (bcql.py:122-132 == cpq.py:125-135 == bearl.py:144-154).
// sa = [obs|act] [B, o+a]; dec_in [B, o+L] has its obs columns already filled.
void emit_vae_update(Engine& e, Program& p, const float* sa, float* dec_in, const float* eps, const float* act,
                     int stat_index) {
  const osrl_config& c = e.plan.cfg;
  const VaeLay& v = e.plan.vae;
  const int B = e.B, o = c.obs_dim, a = c.act_dim, L = 2 * a, V = c.vae_hidden;
  const float iw = e.inv_world();
  float* h1 = e.ws((size_t)B * V); float* h2 = e.ws((size_t)B * V);
  float* ml = e.ws((size_t)B * 2 * L); float* sd = e.ws((size_t)B * L);
  float* g1 = e.ws((size_t)B * V); float* g2 = e.ws((size_t)B * V);
  float* u = e.ws((size_t)B * a); float* dpre3 = e.ws((size_t)B * a);
  float* dg2 = e.ws((size_t)B * V); float* dg1 = e.ws((size_t)B * V);
  float* dz = e.ws((size_t)B * L); float* dml = e.ws((size_t)B * 2 * L);
  float* dh2 = e.ws((size_t)B * V); float* dh1 = e.ws((size_t)B * V);
  float* stat = e.stats + stat_index;
  if (fz_mlp_ok(v.e1, v.e2, v.heads) && fz_mlp_ok(v.d1, v.d2, v.d3) && L <= 16) {
    // fused: encoder, decoder, and the two backward chains are one launch each; weight gradients in two
    const float lim = c.max_action, beta = c.beta;
    emit_fz(e, p, {fz_fwd3(sa, o + a, B, e.P, v.e1, v.e2, v.heads, ACT_RELU, h1, V, h2, V, ml, 2 * L)});
    KOP(p, e, 20.0 * B * L, (k_vae_reparam<<<(B * L + 255) / 256, 256, 0, s>>>(ml, eps, B, L, sd, dec_in, o + L, o)));
    {
      FzTask t = fz_fwd3(dec_in, o + L, B, e.P, v.d1, v.d2, v.d3, ACT_RELU, g1, V, g2, V, u, a);
      t.ract = ACT_TANH; t.rscale = lim;
      emit_fz(e, p, {t});
    }
    KOP(p, e, 12.0 * B * a + 12.0 * B * L, (k_vae_loss<<<1, 1024, 0, s>>>(u, act, B, a, lim, ml, sd, L, beta, dpre3, stat, iw)));
    {
      FzTask t = fz_bwd_mid(dpre3, a, B, e.P, v.d2, v.d3, ACT_RELU, g1, V, g2, V, dg2, V, dg1, V);
      fz_add_dx(t, fz_new_group(), 0, 1, e.P, v.d1, o, L, dz, L);   // d loss / d z: the latent columns of d1's input
      emit_fz(e, p, {t});
    }
    KOP(p, e, 28.0 * B * L, (k_vae_reparam_bwd<<<(B * L + 255) / 256, 256, 0, s>>>(dz, L, 0, ml, sd, eps, B, L, beta, dml, iw)));
    emit_fz(e, p, {fz_bwd_mid(dml, 2 * L, B, e.P, v.e2, v.heads, ACT_RELU, h1, V, h2, V, dh2, V, dh1, V)});
    p.begin_par();
    emit_fz(e, p, {fz_wgrad(dg2, V, g1, V, B, e.G, v.d2), fz_wgrad(dh2, V, h1, V, B, e.G, v.e2)});
    flush(e, p, {task_wgrad(dpre3, a, g2, V, B, e.G, v.d3), task_wgrad(dg1, V, dec_in, o + L, B, e.G, v.d1),
                 task_wgrad(dml, 2 * L, h2, V, B, e.G, v.heads), task_wgrad(dh1, V, sa, o + a, B, e.G, v.e1)});
    p.end_par();
    const Group& g = e.plan.groups[e.plan.g_vae];
    emit_allreduce(e, p, e.G + g.begin, g.end - g.begin, false, DP_GRAD);
    emit_adam(e, p, e.plan.g_vae, g.begin, g.end, false);
    return;
  }
  // encoder
  flush(e, p, {task_fwd(sa, o + a, B, e.P, v.e1, h1, V, ACT_RELU)});
  flush(e, p, {task_fwd(h1, V, B, e.P, v.e2, h2, V, ACT_RELU)});
  flush(e, p, {task_fwd(h2, V, B, e.P, v.heads, ml, 2 * L, ACT_NONE)});
  KOP(p, e, 20.0 * B * L, (k_vae_reparam<<<(B * L + 255) / 256, 256, 0, s>>>(ml, eps, B, L, sd, dec_in, o + L, o)));
  // decoder
  flush(e, p, {task_fwd(dec_in, o + L, B, e.P, v.d1, g1, V, ACT_RELU)});
  flush(e, p, {task_fwd(g1, V, B, e.P, v.d2, g2, V, ACT_RELU)});
  flush(e, p, {task_fwd(g2, V, B, e.P, v.d3, u, a, ACT_TANH, c.max_action)});
  const float lim = c.max_action, beta = c.beta;
  KOP(p, e, 12.0 * B * a + 12.0 * B * L, (k_vae_loss<<<1, 1024, 0, s>>>(u, act, B, a, lim, ml, sd, L, beta, dpre3, stat, iw)));
  // backward
  flush(e, p, {task_wgrad(dpre3, a, g2, V, B, e.G, v.d3), task_dgrad(dpre3, a, B, e.P, v.d3, dg2, V, g2, V, ACT_RELU)});
  flush(e, p, {task_wgrad(dg2, V, g1, V, B, e.G, v.d2), task_dgrad(dg2, V, B, e.P, v.d2, dg1, V, g1, V, ACT_RELU)});
  flush(e, p, {task_wgrad(dg1, V, dec_in, o + L, B, e.G, v.d1),
               task_dgrad(dg1, V, B, e.P, v.d1, dz, L, nullptr, 0, 0, o, L)});
  KOP(p, e, 28.0 * B * L, (k_vae_reparam_bwd<<<(B * L + 255) / 256, 256, 0, s>>>(dz, L, 0, ml, sd, eps, B, L, beta, dml, iw)));
  flush(e, p, {task_wgrad(dml, 2 * L, h2, V, B, e.G, v.heads),
               task_dgrad(dml, 2 * L, B, e.P, v.heads, dh2, V, h2, V, ACT_RELU)});
  flush(e, p, {task_wgrad(dh2, V, h1, V, B, e.G, v.e2), task_dgrad(dh2, V, B, e.P, v.e2, dh1, V, h1, V, ACT_RELU)});
  flush(e, p, {task_wgrad(dh1, V, sa, o + a, B, e.G, v.e1)});
  const Group& g = e.plan.groups[e.plan.g_vae];
  emit_allreduce(e, p, e.G + g.begin, g.end - g.begin, false, DP_GRAD);
  emit_adam(e, p, e.plan.g_vae, g.begin, g.end, false);
}

// Plain MLP forward: hidden activation `hact` on all but the last layer; the last layer's task is
// returned un-emitted so the caller can attach its epilogue (tanh/scale/resid/clamp/aux) and merge it
// into a launch.  h[j] = output of layer j (j < n-1).
GemmTask mlp_fwd_hidden(Engine& e, Program& p, const float* W, const MlpLay& m, const float* X, int ldx, int rows,
                        int hact, std::vector<float*>& h, float* out, int ldout, bool nograd) {
  const int n = (int)m.L.size();
  const float* cur = X;
  int ld = ldx;
  h.clear();
  if (n == 3 && fz_mlp_ok(m.L[0], m.L[1], m.L[2]) && !fz_unfuse_first(rows, nograd)) {
    // fused network: nothing is launched here -- the returned task is a marker that carries the caller's last-layer
    // epilogue to emit_gemm, which completes the fused task (Engine::fz_pending) and launches it
    float *y0 = nullptr, *y1 = nullptr;
    if (!nograd) {
      y0 = e.ws((size_t)rows * m.L[0].out); y1 = e.ws((size_t)rows * m.L[1].out);
      h.push_back(y0); h.push_back(y1);
    }
    e.fz_pending.push_back(fz_fwd3(X, ldx, rows, W, m.L[0], m.L[1], m.L[2], hact, y0, m.L[0].out, y1, m.L[1].out, out, ldout));
    GemmTask mk = task_fwd(nullptr, m.L[1].out, rows, W, m.L[2], out, ldout, ACT_NONE);
    mk.fz_pending = (int)e.fz_pending.size();
    return mk;
  }
  for (int j = 0; j + 1 < n; ++j) {
    float* y = e.ws((size_t)rows * m.L[j].out);
    GemmTask lay = task_fwd(cur, ld, rows, W, m.L[j], y, m.L[j].out, hact);
    if (nograd && j == 0 && n > 2) { lay.pk_gcols = m.L[j].out; lay.c_dead = 1; }
    emit_gemm(e, p, {lay});
    h.push_back(y);
    cur = y;
    ld = m.L[j].out;
  }
  return task_fwd(cur, ld, rows, W, m.L[n - 1], out, ldout, ACT_NONE);
}

// Plain MLP backward from dpre (gradient wrt the last layer's pre-activation) [rows, out].
void mlp_bwd(Engine& e, Program& p, const float* W, float* Gsec, const MlpLay& m, const float* X, int ldx, int rows,
             int hact, const std::vector<float*>& h, const float* dpre) {
  const int n = (int)m.L.size();
  const float* dy = dpre;
  int lddy = m.L[n - 1].out;
  if (n == 3 && fz_mlp_ok(m.L[0], m.L[1], m.L[2]) && h.size() == 2) {
    const int H0 = m.L[0].out, H1 = m.L[1].out;
    float* d1 = e.ws((size_t)rows * H1);
    float* d0 = e.ws((size_t)rows * H0);
    emit_fz(e, p, {fz_bwd_mid(dpre, lddy, rows, W, m.L[1], m.L[2], hact, h[0], H0, h[1], H1, d1, H1, d0, H0)});
    p.begin_par();
    emit_fz(e, p, {fz_wgrad(d1, H1, h[0], H0, rows, Gsec, m.L[1])});
    emit_gemm(e, p, {task_wgrad(dpre, lddy, h[1], H1, rows, Gsec, m.L[2]), task_wgrad(d0, H0, X, ldx, rows, Gsec, m.L[0])});
    p.end_par();
    return;
  }
  for (int j = n - 1; j >= 0; --j) {
    const float* xin = j == 0 ? X : h[j - 1];
    const int ldin = j == 0 ? ldx : m.L[j - 1].out;
    std::vector<GemmTask> ts{task_wgrad(dy, lddy, xin, ldin, rows, Gsec, m.L[j])};
    float* dx = nullptr;
    if (j > 0) {
      dx = e.ws((size_t)rows * m.L[j].in);
      ts.push_back(task_dgrad(dy, lddy, rows, W, m.L[j], dx, m.L[j].in, h[j - 1], m.L[j - 1].out, hact));
    }
    emit_gemm(e, p, ts);
    dy = dx;
    lddy = m.L[j].in;
  }
}

}  // namespace osrl

namespace osrl {
void emit_stages(Engine& e, Program& p, std::vector<Stage>& st) {
  for (auto& s : st) {   // the launches of one stage are independent of each other
    p.begin_par();
    emit_fz(e, p, s.fz);
    emit_gemm(e, p, s.tasks);
    p.end_par();
  }
}
}  // namespace osrl
