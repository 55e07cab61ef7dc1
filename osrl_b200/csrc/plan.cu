// Parameter plan: arena layout + state_dict table of each algorithm (pure CPU code).
//
// Arena sections P (trained), T (targets), G (grads), M, V (Adam moments) all have the same
// float layout, so a parameter's offset is valid in all five.  Optimiser groups are contiguous
// ranges; inside a group, tensors are arranged layer-major across ensemble members so an
// ensemble layer is ONE GEMM (see EnsLay).  The table keeps the reference's state_dict names,
// shapes and order (e.g. bcql.py:85-105 registration order).
#include <cmath>

#include "engine.h"

namespace osrl {
namespace {

struct Alloc {
  int64_t top = 0;
  int64_t take(int64_t n) {
    int64_t o = top;
    top += (n + 3) / 4 * 4;
    return o;
  }
};

Lin alloc_lin(Alloc& a, int in, int out) {
  Lin l;
  l.in = in;
  l.out = out;
  l.w = a.take((int64_t)in * out);
  l.b = a.take(out);
  return l;
}
MlpLay alloc_mlp(Alloc& a, const std::vector<int>& sizes) {
  MlpLay m;
  for (size_t j = 0; j + 1 < sizes.size(); ++j) m.L.push_back(alloc_lin(a, sizes[j], sizes[j + 1]));
  return m;
}
EnsLay alloc_ens(Alloc& a, int n, int in, const std::vector<int>& h) {
  EnsLay e;
  e.n = n;
  e.in = in;
  e.h = h;
  e.first = alloc_lin(a, in, n * h[0]);
  for (size_t l = 1; l < h.size(); ++l) {
    std::vector<Lin> row;
    // weights of all nets contiguous, then biases contiguous (so a layer's biases form one [n*h] vector)
    int64_t w0 = a.take((int64_t)n * h[l] * h[l - 1]);
    int64_t b0 = a.take((int64_t)n * h[l]);
    for (int i = 0; i < n; ++i) {
      Lin x;
      x.in = h[l - 1];
      x.out = h[l];
      x.w = w0 + (int64_t)i * h[l] * h[l - 1];
      x.b = b0 + (int64_t)i * h[l];
      row.push_back(x);
    }
    e.mid.push_back(row);
  }
  e.w_last = a.take((int64_t)n * h.back());
  e.b_last = a.take(n);
  return e;
}
VaeLay alloc_vae(Alloc& a, int o, int act, int V, int L) {
  VaeLay v;
  v.e1 = alloc_lin(a, o + act, V);
  v.e2 = alloc_lin(a, V, V);
  v.heads = alloc_lin(a, V, 2 * L);
  v.d1 = alloc_lin(a, o + L, V);
  v.d2 = alloc_lin(a, V, V);
  v.d3 = alloc_lin(a, V, act);
  return v;
}
SqActorLay alloc_sq(Alloc& a, int o, int act, const std::vector<int>& h) {
  SqActorLay s;
  std::vector<int> sizes{o};
  for (int x : h) sizes.push_back(x);
  s.trunk = alloc_mlp(a, sizes);
  s.heads = alloc_lin(a, h.back(), 2 * act);
  return s;
}

void emit_lin(std::vector<ParamEntry>& t, const std::string& name, int64_t w, int64_t b, int out, int in, int sec,
              int grp) {
  t.push_back({name + ".weight", out, in, w, sec, grp});
  t.push_back({name + ".bias", out, 0, b, sec, grp});
}
void emit_mlp(std::vector<ParamEntry>& t, const std::string& prefix, const MlpLay& m, int sec, int grp) {
  for (size_t j = 0; j < m.L.size(); ++j)
    emit_lin(t, prefix + "." + std::to_string(2 * j), m.L[j].w, m.L[j].b, m.L[j].out, m.L[j].in, sec, grp);
}
// lists: {"q1_nets","q2_nets"} (double) or {"q_nets"} (single); n = lists.size() * num
void emit_ens(std::vector<ParamEntry>& t, const std::string& prefix, const std::vector<std::string>& lists, int num,
              const EnsLay& e, int sec, int grp) {
  const int nh = (int)e.h.size();
  for (size_t li = 0; li < lists.size(); ++li)
    for (int i = 0; i < num; ++i) {
      const int net = (int)li * num + i;
      const std::string p = prefix + "." + lists[li] + "." + std::to_string(i) + ".";
      emit_lin(t, p + "0", e.first.w + (int64_t)net * e.h[0] * e.in, e.first.b + (int64_t)net * e.h[0], e.h[0], e.in,
               sec, grp);
      for (int l = 1; l < nh; ++l)
        emit_lin(t, p + std::to_string(2 * l), e.mid[l - 1][net].w, e.mid[l - 1][net].b, e.h[l], e.h[l - 1], sec, grp);
      emit_lin(t, p + std::to_string(2 * nh), e.w_last + (int64_t)net * e.h.back(), e.b_last + net, 1, e.h.back(),
               sec, grp);
    }
}
void emit_vae(std::vector<ParamEntry>& t, const std::string& p, const VaeLay& v, int L, int sec, int grp) {
  emit_lin(t, p + ".e1", v.e1.w, v.e1.b, v.e1.out, v.e1.in, sec, grp);
  emit_lin(t, p + ".e2", v.e2.w, v.e2.b, v.e2.out, v.e2.in, sec, grp);
  emit_lin(t, p + ".mean", v.heads.w, v.heads.b, L, v.heads.in, sec, grp);
  emit_lin(t, p + ".log_std", v.heads.w + (int64_t)L * v.heads.in, v.heads.b + L, L, v.heads.in, sec, grp);
  emit_lin(t, p + ".d1", v.d1.w, v.d1.b, v.d1.out, v.d1.in, sec, grp);
  emit_lin(t, p + ".d2", v.d2.w, v.d2.b, v.d2.out, v.d2.in, sec, grp);
  emit_lin(t, p + ".d3", v.d3.w, v.d3.b, v.d3.out, v.d3.in, sec, grp);
}
void emit_sq(std::vector<ParamEntry>& t, const std::string& p, const SqActorLay& s, int act, int sec, int grp) {
  emit_mlp(t, p + ".net", s.trunk, sec, grp);
  emit_lin(t, p + ".mu_layer", s.heads.w, s.heads.b, act, s.heads.in, sec, grp);
  emit_lin(t, p + ".log_std_layer", s.heads.w + (int64_t)act * s.heads.in, s.heads.b + act, act, s.heads.in, sec, grp);
}

std::vector<int> hidden(const int32_t* h, int n) {
  OSRL_REQUIRE(n >= 1 && n <= OSRL_MAX_HIDDEN, "hidden layer count must be 1..4");
  std::vector<int> v(h, h + n);
  for (int x : v) OSRL_REQUIRE(x > 0, "hidden size must be positive");
  return v;
}

int add_group(Plan& p, const std::string& name, int64_t begin, int64_t end, float lr, bool tgt) {
  Group g;
  g.name = name;
  g.begin = begin;
  g.end = end;
  g.lr = lr;
  g.has_target = tgt;
  p.groups.push_back(g);
  return (int)p.groups.size() - 1;
}

}  // namespace

Plan make_plan(const osrl_config& cfg) {
  Plan p;
  p.cfg = cfg;
  OSRL_REQUIRE(cfg.obs_dim > 0 && cfg.act_dim > 0, "obs_dim/act_dim must be positive");
  OSRL_REQUIRE(cfg.world_size >= 1 && cfg.rank >= 0 && cfg.rank < cfg.world_size, "bad world_size/rank");
  const int o = cfg.obs_dim, a = cfg.act_dim, L = 2 * a;
  Alloc al;
  // bcql.py:109-110 / cpq.py:103-105 (python double arithmetic)
  if (cfg.algo != OSRL_ALGO_BC && cfg.algo != OSRL_ALGO_CDT) {   // (COptiDICE: coptidice.py:101-102)
    const double g = (double)cfg.gamma;
    p.q_thres = p.qc_thres =
        (double)cfg.cost_limit * (1.0 - std::pow(g, (double)cfg.episode_len)) / (1.0 - g) / (double)cfg.episode_len;
  }
  switch (cfg.algo) {
    case OSRL_ALGO_BC: {
      std::vector<int> sizes{o};
      for (int x : hidden(cfg.a_hidden, cfg.n_a_hidden)) sizes.push_back(x);
      sizes.push_back(a);
      int64_t b0 = al.top;
      p.mlp_actor = alloc_mlp(al, sizes);
      p.g_actor = add_group(p, "actor", b0, al.top, cfg.actor_lr, false);
      emit_mlp(p.table, "actor.pi", p.mlp_actor, 0, p.g_actor);
      p.stat_names = {"loss/actor_loss"};
      break;
    }
    case OSRL_ALGO_BCQL: {
      OSRL_REQUIRE(cfg.num_q >= 1 && cfg.num_qc >= 1 && cfg.sample_action_num >= 1 && cfg.vae_hidden > 0, "bad BCQL config");
      auto ah = hidden(cfg.a_hidden, cfg.n_a_hidden);
      auto ch = hidden(cfg.c_hidden, cfg.n_c_hidden);
      int64_t b0 = al.top;
      p.vae = alloc_vae(al, o, a, cfg.vae_hidden, L);
      p.g_vae = add_group(p, "vae", b0, al.top, cfg.vae_lr, false);
      b0 = al.top;
      p.critic = alloc_ens(al, 2 * cfg.num_q, o + a, ch);
      p.g_critic = add_group(p, "critic", b0, al.top, cfg.critic_lr, true);
      b0 = al.top;
      p.cost_critic = alloc_ens(al, 2 * cfg.num_qc, o + a, ch);
      p.g_cost = add_group(p, "cost_critic", b0, al.top, cfg.critic_lr, true);
      b0 = al.top;
      std::vector<int> sizes{o + a};
      for (int x : ah) sizes.push_back(x);
      sizes.push_back(a);
      p.mlp_actor = alloc_mlp(al, sizes);
      p.g_actor = add_group(p, "actor", b0, al.top, cfg.actor_lr, true);
      const std::vector<std::string> dq{"q1_nets", "q2_nets"};
      emit_mlp(p.table, "actor.pi", p.mlp_actor, 0, p.g_actor);
      emit_ens(p.table, "critic", dq, cfg.num_q, p.critic, 0, p.g_critic);
      emit_ens(p.table, "cost_critic", dq, cfg.num_qc, p.cost_critic, 0, p.g_cost);
      emit_vae(p.table, "vae", p.vae, L, 0, p.g_vae);
      emit_mlp(p.table, "actor_old.pi", p.mlp_actor, 1, -1);
      emit_ens(p.table, "critic_old", dq, cfg.num_q, p.critic, 1, -1);
      emit_ens(p.table, "cost_critic_old", dq, cfg.num_qc, p.cost_critic, 1, -1);
      p.stat_names = {"loss/loss_vae",   "loss/critic_loss", "loss/cost_critic_loss",
                      "loss/actor_loss", "loss/qc_penalty",  "loss/lagrangian"};
      const int64_t B = cfg.batch_size, S = cfg.sample_action_num;
      p.noise = {{"vae_eps", B * L}, {"z_critic", B * S * L}, {"z_cost", B * S * L}, {"z_actor", B * L}};
      break;
    }
    case OSRL_ALGO_CPQ:
    case OSRL_ALGO_BEARL: {
      const bool bear = cfg.algo == OSRL_ALGO_BEARL;
      OSRL_REQUIRE(cfg.num_q >= 1 && cfg.num_qc >= 1 && cfg.sample_action_num >= 1 && cfg.vae_hidden > 0, "bad config");
      auto ah = hidden(cfg.a_hidden, cfg.n_a_hidden);
      auto ch = hidden(cfg.c_hidden, cfg.n_c_hidden);
      const int mult = bear ? 2 : 1;
      int64_t b0 = al.top;
      p.vae = alloc_vae(al, o, a, cfg.vae_hidden, L);
      p.g_vae = add_group(p, "vae", b0, al.top, cfg.vae_lr, false);
      b0 = al.top;
      p.critic = alloc_ens(al, mult * cfg.num_q, o + a, ch);
      p.g_critic = add_group(p, "critic", b0, al.top, cfg.critic_lr, true);
      b0 = al.top;
      p.cost_critic = alloc_ens(al, mult * cfg.num_qc, o + a, ch);
      p.g_cost = add_group(p, "cost_critic", b0, al.top, cfg.critic_lr, true);
      b0 = al.top;
      p.sq_actor = alloc_sq(al, o, a, ah);
      p.g_actor = add_group(p, "actor", b0, al.top, cfg.actor_lr, true);
      const std::vector<std::string> lists = bear ? std::vector<std::string>{"q1_nets", "q2_nets"}
                                                  : std::vector<std::string>{"q_nets"};
      emit_sq(p.table, "actor", p.sq_actor, a, 0, p.g_actor);
      emit_ens(p.table, "critic", lists, cfg.num_q, p.critic, 0, p.g_critic);
      if (bear) {  // bearl.py:96-113: actor, critic, cost_critic, vae
        emit_ens(p.table, "cost_critic", lists, cfg.num_qc, p.cost_critic, 0, p.g_cost);
        emit_vae(p.table, "vae", p.vae, L, 0, p.g_vae);
      } else {     // cpq.py:77-92: actor, critic, vae, cost_critic
        emit_vae(p.table, "vae", p.vae, L, 0, p.g_vae);
        emit_ens(p.table, "cost_critic", lists, cfg.num_qc, p.cost_critic, 0, p.g_cost);
      }
      emit_sq(p.table, "actor_old", p.sq_actor, a, 1, -1);
      emit_ens(p.table, "critic_old", lists, cfg.num_q, p.critic, 1, -1);
      emit_ens(p.table, "cost_critic_old", lists, cfg.num_qc, p.cost_critic, 1, -1);
      const int64_t B = cfg.batch_size, S = cfg.sample_action_num;
      if (bear) {
        const int64_t N = cfg.num_samples_mmd_match;
        OSRL_REQUIRE(N >= 1 && N <= 32, "num_samples_mmd_match must be 1..32");
        p.stat_names = {"loss/loss_vae",  "loss/critic_loss", "loss/cost_critic_loss", "loss/actor_loss",
                        "loss/mmd_loss",  "loss/qc_penalty",  "loss/lagrangian",       "loss/alpha_value"};
        p.noise = {{"vae_eps", B * L}, {"pi_critic", B * S * a}, {"pi_cost", B * S * a}, {"z_mmd", B * N * L},
                   {"pi_actor", B * N * a}};
      } else {
        p.qc_thres = (double)cfg.qc_scalar * p.q_thres;
        p.stat_names = {"loss/loss_vae", "loss/critic_loss", "loss/cost_critic_loss", "loss/alpha_value",
                        "loss/actor_loss"};
        p.noise = {{"vae_eps", B * L},          {"pi_critic", B * a}, {"pi_cost", B * a},
                   {"ood_sample", S * B * a},   {"pi_actor", B * a}};
      }
      break;
    }
    case OSRL_ALGO_COPTIDICE: {
      // coptidice.py:106-119: actor (squashed Gaussian), nu_network / chi_network = EnsembleQCritic(state_dim, act_dim=0).
      // tau and lmbda are plain tensors outside state_dict: they live in DevState (osrl_scalars_get / set).
      OSRL_REQUIRE(cfg.num_nu >= 1 && cfg.num_chi >= 1 && cfg.alpha > 0.f && cfg.init_state_propotion > 0.f &&
                       cfg.f_type >= 0 && cfg.f_type <= 2, "bad COptiDICE config");
      OSRL_REQUIRE(cfg.world_size == 1, "COptiDICE is single-GPU: its chi weights are a softmax over the whole batch");
      auto ah = hidden(cfg.a_hidden, cfg.n_a_hidden);
      auto ch = hidden(cfg.c_hidden, cfg.n_c_hidden);
      int64_t b0 = al.top;
      p.sq_actor = alloc_sq(al, o, a, ah);
      p.g_actor = add_group(p, "actor", b0, al.top, cfg.actor_lr, false);
      b0 = al.top;
      p.critic = alloc_ens(al, cfg.num_nu, o, ch);          // nu_network
      p.g_critic = add_group(p, "nu_network", b0, al.top, cfg.critic_lr, false);
      b0 = al.top;
      p.cost_critic = alloc_ens(al, cfg.num_chi, o, ch);    // chi_network
      p.g_cost = add_group(p, "chi_network", b0, al.top, cfg.critic_lr, false);
      emit_sq(p.table, "actor", p.sq_actor, a, 0, p.g_actor);
      emit_ens(p.table, "nu_network", {"q_nets"}, cfg.num_nu, p.critic, 0, p.g_critic);
      emit_ens(p.table, "chi_network", {"q_nets"}, cfg.num_chi, p.cost_critic, 0, p.g_cost);
      p.stat_names = {"loss/chi_loss", "loss/tau_loss",   "loss/D_kl",       "loss/Df",  "loss/td_error",
                      "loss/nu_loss",  "loss/lmbda_loss", "loss/actor_loss", "loss/tau", "loss/lmbda"};
      const int64_t B = cfg.batch_size;
      p.noise = {{"obs_eps", B * o}, {"act_eps", B * a}};
      break;
    }
    case OSRL_ALGO_CDT: {
      const int E = cfg.embedding_dim, T = cfg.seq_len, NL = cfg.num_layers, H = cfg.num_heads;
      OSRL_REQUIRE(E > 0 && E % 32 == 0 && E <= 512, "embedding_dim must be a multiple of 32, <= 512");
      OSRL_REQUIRE(H > 0 && E % H == 0 && (E / H == 8 || E / H == 16 || E / H == 32), "head dim must be 8, 16 or 32");
      OSRL_REQUIRE(T > 0 && NL > 0 && cfg.episode_len > 0, "bad CDT config");
      OSRL_REQUIRE(cfg.use_rew && cfg.use_cost && cfg.cost_transform && cfg.stochastic,
                   "this build covers the reference's configured CDT mode: use_rew, use_cost, cost_transform, stochastic");
      for (float dp : {cfg.attention_dropout, cfg.residual_dropout, cfg.embedding_dropout})
        OSRL_REQUIRE(dp >= 0.f && dp < 1.f, "dropout probabilities must be in [0, 1)");
      {   // dropout multipliers travel as noise slots (replayable, see osrl_noise): one per dropout site
        const int64_t NE = (int64_t)cfg.batch_size * 4 * T * E, NA = (int64_t)cfg.batch_size * H * 4 * T * 4 * T;
        if (cfg.embedding_dropout > 0.f) p.noise.push_back({"drop_emb", NE});
        for (int i = 0; i < NL; ++i) {
          if (cfg.attention_dropout > 0.f) p.noise.push_back({"drop_attn" + std::to_string(i), NA});
          if (cfg.residual_dropout > 0.f) {
            p.noise.push_back({"drop_res" + std::to_string(i) + "a", NE});
            p.noise.push_back({"drop_res" + std::to_string(i) + "b", NE});
          }
        }
      }
      CdtLay& L_ = p.cdt;
      L_.E = E;
      L_.te_rows = cfg.episode_len + T;
      const int64_t b0 = al.top;
      L_.emb_norm_w = al.take(E); L_.emb_norm_b = al.take(E);
      L_.out_norm_w = al.take(E); L_.out_norm_b = al.take(E);
      L_.te = al.take((int64_t)L_.te_rows * E);
      L_.state_emb = alloc_lin(al, o, E);
      L_.action_emb = alloc_lin(al, a, E);
      L_.cost_emb = alloc_lin(al, 1, E);
      L_.return_emb = alloc_lin(al, 1, E);
      for (int i = 0; i < NL; ++i) {
        CdtLay::Blk b;
        b.n1w = al.take(E); b.n1b = al.take(E); b.n2w = al.take(E); b.n2b = al.take(E);
        b.in_proj = alloc_lin(al, E, 3 * E);
        b.out_proj = alloc_lin(al, E, E);
        b.fc1 = alloc_lin(al, E, 4 * E);
        b.fc2 = alloc_lin(al, 4 * E, E);
        L_.blocks.push_back(b);
      }
      L_.act_head = alloc_lin(al, E, 2 * a);
      L_.aux_head = alloc_lin(al, E, 2 + o);
      p.g_cdt = add_group(p, "cdt", b0, al.top, cfg.learning_rate, false);
      Group& g = p.groups[p.g_cdt];
      g.beta1 = std::round((double)cfg.adam_beta1 * 1e6) / 1e6;   // floats from the ABI -> the python doubles
      g.beta2 = std::round((double)cfg.adam_beta2 * 1e6) / 1e6;
      g.wd = cfg.weight_decay;
      g.warmup = cfg.lr_warmup_steps;
      auto& t = p.table;
      const int G_ = p.g_cdt;
      t.push_back({"emb_norm.weight", E, 0, L_.emb_norm_w, 0, G_});
      t.push_back({"emb_norm.bias", E, 0, L_.emb_norm_b, 0, G_});
      t.push_back({"out_norm.weight", E, 0, L_.out_norm_w, 0, G_});
      t.push_back({"out_norm.bias", E, 0, L_.out_norm_b, 0, G_});
      t.push_back({"timestep_emb.weight", L_.te_rows, E, L_.te, 0, G_});
      emit_lin(t, "state_emb", L_.state_emb.w, L_.state_emb.b, E, o, 0, G_);
      emit_lin(t, "action_emb", L_.action_emb.w, L_.action_emb.b, E, a, 0, G_);
      emit_lin(t, "cost_emb", L_.cost_emb.w, L_.cost_emb.b, E, 1, 0, G_);
      emit_lin(t, "return_emb", L_.return_emb.w, L_.return_emb.b, E, 1, 0, G_);
      for (int i = 0; i < NL; ++i) {
        const CdtLay::Blk& b = L_.blocks[i];
        const std::string pre = "blocks." + std::to_string(i) + ".";
        t.push_back({pre + "norm1.weight", E, 0, b.n1w, 0, G_});
        t.push_back({pre + "norm1.bias", E, 0, b.n1b, 0, G_});
        t.push_back({pre + "norm2.weight", E, 0, b.n2w, 0, G_});
        t.push_back({pre + "norm2.bias", E, 0, b.n2b, 0, G_});
        t.push_back({pre + "attention.in_proj_weight", 3 * E, E, b.in_proj.w, 0, G_});
        t.push_back({pre + "attention.in_proj_bias", 3 * E, 0, b.in_proj.b, 0, G_});
        emit_lin(t, pre + "attention.out_proj", b.out_proj.w, b.out_proj.b, E, E, 0, G_);
        emit_lin(t, pre + "mlp.0", b.fc1.w, b.fc1.b, 4 * E, E, 0, G_);
        emit_lin(t, pre + "mlp.2", b.fc2.w, b.fc2.b, E, 4 * E, 0, G_);
      }
      emit_lin(t, "action_head.mu", L_.act_head.w, L_.act_head.b, a, E, 0, G_);
      emit_lin(t, "action_head.log_std", L_.act_head.w + (int64_t)a * E, L_.act_head.b + a, a, E, 0, G_);
      emit_lin(t, "state_pred_head", L_.aux_head.w + 2 * (int64_t)E, L_.aux_head.b + 2, o, E, 0, G_);
      emit_lin(t, "cost_pred_head", L_.aux_head.w, L_.aux_head.b, 2, E, 0, G_);
      p.stat_names = {"nll", "ent", "ent_reg", "all_loss", "act_loss", "cost_loss", "cost_acc", "state_loss", "train_lr"};
      break;
    }
    default:
      throw Err(OSRL_ERR_UNSUPPORTED, "algorithm id not supported by this build");
  }
  p.nP = al.top;
  OSRL_REQUIRE((int)p.groups.size() <= OSRL_MAX_GROUPS, "too many optimiser groups");
  OSRL_REQUIRE((int)p.noise.size() <= OSRL_MAX_NOISE, "too many noise slots");
  return p;
}

}  // namespace osrl
