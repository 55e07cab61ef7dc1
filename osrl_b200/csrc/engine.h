// Engine internals shared by the per-algorithm program builders (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/osrl_b200.h"
#include "gemm.cuh"
#include "gemm_fz.cuh"
#include "kernels.cuh"

namespace osrl {

struct Err : std::runtime_error {
  int code;
  Err(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define OSRL_CUDA(x)                                                                                   \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess)                                                                             \
      throw ::osrl::Err(OSRL_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                                           std::to_string(__LINE__));                                  \
  } while (0)
#define OSRL_REQUIRE(c, msg)                                   \
  do {                                                         \
    if (!(c)) throw ::osrl::Err(OSRL_ERR_ARG, std::string(msg)); \
  } while (0)

// ------------------------------------------------------------------ parameter layout (pure CPU)
struct Lin {  // one nn.Linear: weight [out, in] row-major at w, bias [out] at b (float offsets in a section)
  int64_t w = 0, b = 0;
  int in = 0, out = 0;
};
struct MlpLay { std::vector<Lin> L; };
// An ensemble of n identical ReLU MLPs in -> h[0] -> ... -> 1 (net.py:208-287), stored layer-major:
// first layer of all nets stacked [n*h0, in]; middle layers block-diagonal; last layers stacked [n, h_last].
struct EnsLay {
  int n = 0, in = 0;
  std::vector<int> h;
  Lin first;                          // out = n*h[0]
  std::vector<std::vector<Lin>> mid;  // mid[l-1][net] : h[l-1] -> h[l]
  int64_t w_last = 0, b_last = 0;     // [n, h.back()], [n]
};
struct VaeLay { Lin e1, e2, heads /*[2L, V]: mean rows then log_std rows*/, d1, d2, d3; };
struct SqActorLay { MlpLay trunk; Lin heads; /*[2a, H]: mu rows then log_std rows*/ };

struct CdtLay {  // CDT parameter offsets (cdt.py:45-148)
  int E = 0, te_rows = 0;
  int64_t emb_norm_w = 0, emb_norm_b = 0, out_norm_w = 0, out_norm_b = 0, te = 0;
  Lin state_emb, action_emb, cost_emb, return_emb;
  struct Blk { int64_t n1w, n1b, n2w, n2b; Lin in_proj, out_proj, fc1, fc2; };
  std::vector<Blk> blocks;
  Lin act_head;  // [2a, E]: mu rows then log_std rows (DiagGaussianActor, net.py:509-533)
  Lin aux_head;  // [2+o, E]: cost_pred_head rows then state_pred_head rows
};

struct ParamEntry {
  std::string name;
  int64_t rows, cols, offset;
  int section, group;
};
struct Group {
  std::string name;
  int64_t begin = 0, end = 0;
  float lr = 0.f;
  double beta1 = 0.9, beta2 = 0.999;  // kept in double: torch derives 1-beta and beta^t from python doubles
  float eps = 1e-8f, wd = 0.f;
  int warmup = 0;
  bool has_target = false;
};

struct Plan {
  osrl_config cfg{};
  std::vector<ParamEntry> table;  // in the reference's state_dict order
  std::vector<Group> groups;
  int64_t nP = 0;
  // layouts (which ones are used depends on cfg.algo)
  MlpLay mlp_actor;       // BC actor / BCQL perturbation actor
  SqActorLay sq_actor;    // CPQ / BEAR-Lag
  EnsLay critic, cost_critic;
  VaeLay vae;
  CdtLay cdt;
  int g_cdt = -1;
  int g_actor = -1, g_critic = -1, g_cost = -1, g_vae = -1;
  double qc_thres = 0.0, q_thres = 0.0;
  std::vector<std::string> stat_names;
  std::vector<std::pair<std::string, int64_t>> noise;  // slot name, floats per step
};
Plan make_plan(const osrl_config& cfg);

// ------------------------------------------------------------------ runtime
using Op = std::function<void(cudaStream_t)>;

struct Engine;
struct OpMeta {
  std::string name;     // kernel (or collective) name
  double bytes = 0.0;   // algorithmic bytes of this launch (operands read once + results written once)
  double flops = 0.0;   // algorithmic flops of this launch
  bool kernel = true;   // one of OUR kernels (false: NCCL collective)
  int par = 0;          // > 0: ops that share this id are independent of each other (run on forked graph branches)
};
struct Program {
  std::vector<Op> ops;
  std::vector<OpMeta> meta;
  int kernels = 0;
  int par_open = 0, par_next = 0;
  void add(const std::string& name, double bytes, double flops, bool kernel, Op fn) {
    ops.push_back(std::move(fn));
    meta.push_back({name, bytes, flops, kernel, par_open});
    if (kernel) ++kernels;
  }
  // launches emitted between begin_par() and end_par() do not depend on each other: when the step is captured they
  // become parallel branches of the graph (engine.cu: run_ops) instead of one serial chain
  void begin_par() { par_open = ++par_next; }
  void end_par() {
    for (auto& m : meta)   // an accumulate-into-zeroed-buffer pair (split-K) must stay ordered: give the group up
      if (m.par == par_open && m.name == "memset")
        for (auto& m2 : meta)
          if (m2.par == par_open) m2.par = 0;
    par_open = 0;
  }
};
// elementwise / reduction kernel launch as one program op: KOP(p, e, bytes, (kernel<<<...>>>(args)))
#define KOP(p, e, bytes_, ...)                                              \
  do {                                                                      \
    ::osrl::Engine* ep_ = &(e);                                             \
    std::string nm_ = #__VA_ARGS__;                                         \
    nm_ = nm_.substr(nm_.find_first_not_of("( "));                          \
    nm_ = nm_.substr(0, nm_.find_first_of("<( "));                          \
    (p).add(nm_, (double)(bytes_), 0.0, true, [=](cudaStream_t s) {         \
      __VA_ARGS__;                                                          \
      ep_->launches++;                                                      \
    });                                                                     \
  } while (0)

struct Engine {
  Plan plan;
  int device = 0;
  float *P = nullptr, *T = nullptr, *G = nullptr, *M = nullptr, *V = nullptr;
  DevState* ds = nullptr;
  AdamGroupCfg* d_groups = nullptr;
  // staging minibatch (device)
  int B = 0;
  float *b_obs = nullptr, *b_nobs = nullptr, *b_act = nullptr, *b_rew = nullptr, *b_cost = nullptr, *b_done = nullptr;
  int64_t* b_idx = nullptr;
  float *b_init = nullptr, *cop_obs_std = nullptr, *cop_act_std = nullptr;   // COptiDICE: is_init column, dataset std vectors
  // CDT sequence minibatch staging (device): [B*T, .] row-major
  float *s_states = nullptr, *s_actions = nullptr, *s_returns = nullptr, *s_ctg = nullptr, *s_mask = nullptr,
        *s_costs = nullptr;
  long long* s_ts = nullptr;
  // resident trajectory buffer (CDT)
  float* sq_rows = nullptr;
  long long* sq_off = nullptr;
  float* sq_prob = nullptr;
  int* sq_alias = nullptr;
  int sq_ntraj = 0, sq_stride = 0;
  float *sq_first_ret = nullptr, *sq_first_cret = nullptr;   // osrl_seq_preprocess: unscaled return / cost return of each episode
  int *s_traj = nullptr, *s_start = nullptr;
  // noise slots
  std::vector<float*> noise_buf;
  NoiseSlot *d_slots_all = nullptr, *d_slots_dyn = nullptr;
  // resident dataset
  float* ds_rows = nullptr;
  int64_t ds_n = 0;
  int ds_stride = 0;
  // program + graphs
  Program body;
  cudaGraphExec_t g_body = nullptr, g_sampled = nullptr;
  // pipelined osrl_steps(k >= 2): the VAE update of step s+1 (program pa, minibatch nb_*, counter vae_step) runs
  // concurrently with the critic / actor updates of step s (program pm, VAE weights read from the snapshot Psnap)
  Program pa, pm;
  bool pipelined = false;
  float *nb_obs = nullptr, *nb_nobs = nullptr, *nb_act = nullptr, *nb_rew = nullptr, *nb_cost = nullptr, *nb_done = nullptr;
  int64_t* nb_idx = nullptr;
  float* Psnap = nullptr;
  struct Handoff { float* next; float* cur; size_t bytes; };   // side-branch results of step s+1 -> main branch, copied
  std::vector<Handoff> handoff;                                // next -> cur before each fork
  bool decode_side = false;   // BCQ-Lag: the step's VAE decodes run on the VAE branch too (OSRL_PIPELINE_DECODE)
  NoiseSlot *d_slots_vae = nullptr, *d_slots_rest = nullptr;
  cudaGraphExec_t g_pro = nullptr, g_mid = nullptr, g_last = nullptr;
  // osrl_steps_host: the same graphs with the device draw replaced by a read of the host-batch queue
  cudaGraphExec_t g_xbody = nullptr, g_xpro = nullptr, g_xmid = nullptr, g_xlast = nullptr;
  HostQueue* xq = nullptr;                       // device copy of the queue descriptor
  float *x_ring = nullptr, *x_st_side = nullptr, *x_st_main = nullptr;   // pinned, mapped
  int x_cap = 0;                                 // batches the ring holds
  bool x_pending = false;                        // an asynchronous osrl_steps_host call may still read the ring
  int x_last_k = 0, x_last_pipelined = 0;
  uint32_t side_stat_mask = 0;                   // stats written by the pipelined VAE branch
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t cap_stream = nullptr;
  std::vector<std::pair<cudaStream_t, std::vector<cudaStream_t>>> par_streams;   // helper streams of forked launches
  std::vector<cudaEvent_t> par_events;
  size_t par_events_used = 0;
  float* stats = nullptr;
  float* stats_pinned[2] = {nullptr, nullptr};   // osrl_stats_lagged double buffer
  cudaEvent_t stats_ev[2] = {nullptr, nullptr};
  int stats_slot = 0, stats_calls = 0;
  std::vector<void*> allocs;
  // packed tf32 hi/lo activation images announced by producer GEMM tasks (GemmTask::pk_*), looked up by the
  // consumer that reads the same activations: see emit_gemm
  struct PackReg { const float* C; int ldc, M, N, gcols, gstride, ks, dead; float* hi; float* lo; };
  std::vector<PackReg> pack_regs;
  int64_t launches = 0;
  bool profile_in_graph = false;
  // fused tcgen05 path (gemm_fz.cuh): networks whose last layer is still being described by the caller
  // (mlp_fwd_hidden returns a marker GemmTask, emit_gemm completes and launches the fused task)
  std::vector<FzTask> fz_pending;
  long long* fz_dbg = nullptr;   // OSRL_FZ_DBG timeline of the last fused launch (debug)
  int fz_dbg_ctas = 0;
  struct FzDbg { long long* buf; int ctas; std::string name; };
  std::vector<FzDbg> fz_dbg_all;
  struct FzAlloc { float* rpart; unsigned* rcnt; };
  std::vector<std::pair<int, FzAlloc>> fz_groups;   // reduce-group id -> partial-sum buffers (emit_fz)
  // data parallel
  void* comm = nullptr;
  void* comm2 = nullptr;   // pipelined VAE branch
  // data parallel over NVLink peer memory (dp_peer.cuh): set up by osrl_comm_init when every rank could map every
  // peer (OSRL_DP=nccl keeps the NCCL collectives)
  bool peer_on = false;
  DpPeers peers{};
  DpPeers* d_peers = nullptr;
  DpFlags* dp_flags = nullptr;
  std::vector<void*> ipc_open;
  std::vector<int> dp_slot_group;   // fused reduce+Adam sites: slot -> optimiser group
  int dp_scal_slots = 0;
  int world = 1, rank = 0;

  float* ws(size_t n);  // zero-initialised device workspace
  template <class Tt>
  Tt* upload(const std::vector<Tt>& v) {  // host vector -> device array owned by the engine
    void* p = nullptr;
    OSRL_CUDA(cudaMalloc(&p, (v.empty() ? 1 : v.size()) * sizeof(Tt)));
    if (!v.empty()) OSRL_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(Tt), cudaMemcpyHostToDevice));
    allocs.push_back(p);
    return (Tt*)p;
  }
  float inv_world() const { return 1.f / (float)world; }
};

// program-building helpers (engine.cu)
struct Stage {
  std::vector<GemmTask> tasks;
  std::vector<FzTask> fz;   // fused tcgen05 tasks of this stage (launched before `tasks`)
};
// ---- fused path (engine.cu).  fz_on(): OSRL_GEMM=fz (the default).  A network qualifies when it has two hidden
// layers whose widths are multiples of 4 (16-byte rows), a first layer of <= 16 inputs and a last one of <= 16 outputs.
bool fz_on();
bool fz_mlp_ok(const Lin& l0, const Lin& l1, const Lin& l2);
FzTask fz_blank();
// forward of a whole 3-layer network: X [rows, l0.in] -> act(l0) -> act(l1) -> l2 -> out [rows, l2.out].
// h1 / h2 (optional): fp32 copies of the hidden activations for the backward pass.  The final epilogue
// (FzTask::ract / rscale / rresid / rclamp / raux) is the caller's to set.
FzTask fz_fwd3(const float* X, int ldx, int rows, const float* W, const Lin& l0, const Lin& l1, const Lin& l2, int hact,
               float* h1, int ldh1, float* h2, int ldh2, float* out, int ldo);
// the same without the generated first layer: H1 [rows, l1.in] (already activated) -> act(l1) -> l2 -> out.  Used for
// the large no-grad passes, where regenerating the first layer in every column tile costs more than one thin launch.
FzTask fz_fwd2(const float* H1, int ldh1, int rows, const float* W, const Lin& l1, const Lin& l2, int hact, float* h2,
               int ldh2, float* out, int ldo);
bool fz_unfuse_first(int rows, bool nograd);
// backward through the last and middle layers: dq [rows, l2.out] -> d1 = (dq W2) * act'(h2) (stored if d1 given)
// -> d0 = (d1 W1) * act'(h1) (stored if d0 given).  fz_add_dx() then folds the first layer's input gradient in.
FzTask fz_bwd_mid(const float* dq, int lddq, int rows, const float* W, const Lin& l1, const Lin& l2, int hact,
                  const float* h1, int ldh1, const float* h2, int ldh2, float* d1, int ldd1, float* d0, int ldd0);
// dX [rows, ncols] = sum over the `members` tasks of a group of  d0 W0[:, col0 : col0+ncols]
int fz_new_group();
void fz_add_dx(FzTask& t, int group, int member, int members, const float* W, const Lin& l0, int col0, int ncols,
               float* dX, int lddx);
FzTask fz_wgrad(const float* dY, int lddy, const float* X, int ldx, int rows, float* Gsec, const Lin& l);
void emit_fz(Engine& e, Program& p, std::vector<FzTask> tasks);
GemmTask task_fwd(const float* X, int ldx, int rows, const float* W, const Lin& l, float* Y, int ldy, int act,
                  float scale = 1.f);
GemmTask task_dgrad(const float* dY, int lddy, int rows, const float* W, const Lin& l, float* dX, int lddx,
                    const float* Hprev, int ldh, int dact, int col0 = 0, int ncols = -1);
GemmTask task_wgrad(const float* dY, int lddy, const float* X, int ldx, int rows, float* Gsec, const Lin& l);
void prepare_kernels();
void emit_gemm(Engine& e, Program& p, const std::vector<GemmTask>& tasks);
void emit_copy(Engine& e, Program& p, const std::vector<CopyTask>& tasks);
void emit_adam(Engine& e, Program& p, int group, int64_t begin, int64_t end, bool polyak,
               const float* clip_coef = nullptr);
// kind: DP_PLAIN = a collective whose result is used as is (NCCL); DP_GRAD = the gradients of the optimiser group the
// next emit_adam steps (peer mode: skipped, emit_adam reduces and steps in one kernel); DP_SCALAR = <= 8 fp32
// cross-batch scalars (peer mode: k_dp_scalar)
enum { DP_PLAIN = 0, DP_GRAD = 1, DP_SCALAR = 2 };
void emit_allreduce(Engine& e, Program& p, float* buf, int64_t count, bool f64 = false, int kind = DP_PLAIN);
CopyTask copy_cols(float* dst, int ldd, int dcol0, const float* src, int lds, int scol0, int rows, int cols,
                   int row_div = 1, int row_mod = 1 << 30);

struct EnsBuf {
  int rows = 0;
  std::vector<float*> h;  // h[l] [rows, n*H_l]
  float* q = nullptr;     // [rows, n]
};
EnsBuf ens_alloc(Engine& e, const EnsLay& l, int rows);
// forward: appends to stages[0..nh] (nh+1 stages)
// nograd: nothing but the next layer reads the first layer's activations (target / sampling passes): on large-row
// passes they are then produced as packed tf32 images for the tcgen05 kernel instead of fp32 (GemmTask::pk_*)
void ens_fwd(std::vector<Stage>& st, const EnsLay& l, const float* W, const float* X, int ldx, int rows, EnsBuf& buf,
             bool nograd = false);
// backward: stages[0] = last layer ... stages[nh] = first layer.  Gsec == nullptr: input-gradient only.
// dX (optional) receives d loss / d X[:, xcol0 : xcol0+xcols].
void ens_bwd(std::vector<Stage>& st, const EnsLay& l, const float* W, float* Gsec, const float* X, int ldx, int rows,
             const EnsBuf& act, EnsBuf& grad, const float* dq, float* dX, int lddx, int xcol0, int xcols);

struct MlpBuf {
  int rows = 0;
  std::vector<float*> h;  // outputs of every layer
};
// blocks.cu
void emit_vae_decode(Engine& e, Program& p, const float* W, const float* dec_in, int rows, float* h1, float* h2,
                     float* out, int ldout, int mode, bool nograd = false);
void emit_vae_update(Engine& e, Program& p, const float* sa, float* dec_in, const float* eps, const float* act,
                     int stat_index);
GemmTask mlp_fwd_hidden(Engine& e, Program& p, const float* W, const MlpLay& m, const float* X, int ldx, int rows,
                        int hact, std::vector<float*>& h, float* out, int ldout, bool nograd = false);
void mlp_bwd(Engine& e, Program& p, const float* W, float* Gsec, const MlpLay& m, const float* X, int ldx, int rows,
             int hact, const std::vector<float*>& h, const float* dpre);
void emit_stages(Engine& e, Program& p, std::vector<Stage>& st);

void build_bc(Engine& e);
void build_bcql(Engine& e, int phase = 0);
void build_cpq(Engine& e, int phase = 0);
void build_bearl(Engine& e, int phase = 0);
void build_cdt(Engine& e);
void build_coptidice(Engine& e);

}  // namespace osrl
