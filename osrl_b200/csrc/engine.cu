// Engine runtime + C ABI (see include/osrl_b200.h).
#include "engine.h"
#include "gemm_mma.cuh"
#include "gemm_tc5.cuh"
#include "gemm_thin.cuh"
#include "gemm_fz.cuh"
#include "cdt_kernels.cuh"
#include "preproc_kernels.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>

namespace osrl {

thread_local std::string g_err;

// ------------------------------------------------------------------ memory helpers
float* Engine::ws(size_t n) {
  if (n == 0) n = 1;
  void* p = nullptr;
  OSRL_CUDA(cudaMalloc(&p, (n + 4) * sizeof(float)));
  OSRL_CUDA(cudaMemset(p, 0, (n + 4) * sizeof(float)));
  allocs.push_back(p);
  return (float*)p;
}
// ------------------------------------------------------------------ GEMM task constructors
static GemmTask blank_task() {
  GemmTask t;
  memset(&t, 0, sizeof(t));
  t.scale = 1.f;
  return t;
}
GemmTask task_fwd(const float* X, int ldx, int rows, const float* W, const Lin& l, float* Y, int ldy, int act,
                  float scale) {
  GemmTask t = blank_task();
  t.A = X; t.lda = ldx; t.a_kc = 1;
  t.B = W + l.w; t.ldb = l.in; t.b_kc = 1;
  t.C = Y; t.ldc = ldy;
  t.M = rows; t.N = l.out; t.K = l.in;
  t.bias = W + l.b;
  t.act = act; t.scale = scale;
  return t;
}
GemmTask task_dgrad(const float* dY, int lddy, int rows, const float* W, const Lin& l, float* dX, int lddx,
                    const float* Hprev, int ldh, int dact, int col0, int ncols) {
  GemmTask t = blank_task();
  if (ncols < 0) ncols = l.in - col0;
  t.A = dY; t.lda = lddy; t.a_kc = 1;
  t.B = W + l.w + col0; t.ldb = l.in; t.b_kc = 0;
  t.C = dX; t.ldc = lddx;
  t.M = rows; t.N = ncols; t.K = l.out;
  t.dact = dact; t.dact_src = Hprev; t.ld_dact = ldh;
  return t;
}
GemmTask task_wgrad(const float* dY, int lddy, const float* X, int ldx, int rows, float* Gsec, const Lin& l) {
  GemmTask t = blank_task();
  t.A = dY; t.lda = lddy; t.a_kc = 0;
  t.B = X; t.ldb = ldx; t.b_kc = 0;
  t.C = Gsec + l.w; t.ldc = l.in;
  t.M = l.out; t.N = l.in; t.K = rows;
  t.colsum = Gsec + l.b;
  return t;
}

// ------------------------------------------------------------------ launch emitters
// tile shapes: {BM, BN, BK, TM, TN, NSTAGE}
#define OSRL_GEMM_CFG0 128, 64, 16, 8, 4, 4
#define OSRL_GEMM_CFG1 64, 64, 32, 4, 4, 4
#define OSRL_GEMM_CFG2 32, 32, 32, 2, 2, 6
// tensor-core (3xTF32 mma.sync) tile shapes: {BM, BN, BK, WARPS_M, WARPS_N, NSTAGE}
#define OSRL_MMA_CFG0 128, 64, 16, 4, 2, 4, 1
#define OSRL_MMA_CFG1 64, 64, 32, 2, 4, 4, 1
#define OSRL_MMA_CFG2 32, 32, 32, 2, 2, 6, 1
// k-group variants (gemm_mma.cuh) for the short-K, less-than-a-wave launches, where the per-tile critical path
// is the kernel time; long K (split-K weight gradients) and multi-wave launches stay on the plain tiles
#define OSRL_MMA_CFG1K 64, 64, 32, 2, 4, 4, 2
#define OSRL_MMA_CFG2K 32, 32, 32, 2, 2, 6, 4
// Every kernel exists in two variants: the basic one (the MLP algorithms: bias/ReLU/Tanh/residual/clamp
// epilogues) and FULL (adds exact GELU + split-K accumulation, used by the launches of the CDT program that
// need them).  Keeping the extras out of the basic variant is worth ~20 % on the BCQ-Lag step.
template <int BM, int BN, int BK, int TM, int TN, int NS, bool FULL>
static void launch_gemm(const TaskPack& d, int ntasks, int tiles, cudaStream_t s) {
  using Cfg = GemmCfg<BM, BN, BK, TM, TN, NS>;
  k_gemm_tasks<BM, BN, BK, TM, TN, NS, FULL><<<tiles, Cfg::NT, Cfg::SMEM_BYTES, s>>>(d, ntasks);
}
template <int BM, int BN, int BK, int TM, int TN, int NS>
static void prepare_gemm() {
  using Cfg = GemmCfg<BM, BN, BK, TM, TN, NS>;
  OSRL_CUDA(cudaFuncSetAttribute(k_gemm_tasks<BM, BN, BK, TM, TN, NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(k_gemm_tasks<BM, BN, BK, TM, TN, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES));
}
template <int BM, int BN, int BK, int WM, int WN, int NS, int KG, bool FULL>
static void launch_mma(const TaskPack& d, int ntasks, int tiles, cudaStream_t s) {
  using Cfg = MmaCfg<BM, BN, BK, WM, WN, NS, KG>;
  k_gemm_mma<BM, BN, BK, WM, WN, NS, KG, FULL><<<tiles, Cfg::NT, Cfg::SMEM_BYTES, s>>>(d, ntasks);
}
template <int BM, int BN, int BK, int WM, int WN, int NS, int KG>
static void prepare_mma() {
  using Cfg = MmaCfg<BM, BN, BK, WM, WN, NS, KG>;
  OSRL_CUDA(cudaFuncSetAttribute(k_gemm_mma<BM, BN, BK, WM, WN, NS, KG, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(k_gemm_mma<BM, BN, BK, WM, WN, NS, KG, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
}
// OSRL_GEMM (read when an engine's program is built): "ffma" = CUDA-core kernel (gemm.cuh); "mma" = 3xTF32
// mma.sync kernel for everything; "tc5" = mma.sync + the tcgen05/TMEM kernel (gemm_tc5.cuh) for the large forward
// layers (round 1); default "fz" = the fused tcgen05 kernel (gemm_fz.cuh): whole 3-layer network passes in one launch,
// every other tiled GEMM (all operand layouts) on tcgen05 as well.
static std::string gemm_mode() {
  const char* e = getenv("OSRL_GEMM");
  return e ? std::string(e) : std::string("fz");
}
bool fz_on() { return gemm_mode() == "fz"; }
static bool use_mma() { return gemm_mode() != "ffma"; }
static TaskPack make_pack(const std::vector<GemmTask>& tasks) {
  OSRL_REQUIRE(tasks.size() <= (size_t)PACK_MAX, "task pack overflow");
  TaskPack p;
  memset(&p, 0, sizeof(p));
  for (size_t i = 0; i < tasks.size(); ++i) p.t[i] = tasks[i];
  return p;
}
// task lists longer than one parameter pack run as several launches
template <class F>
static bool split_packs(const std::vector<GemmTask>& tasks, F&& emit_one) {
  if (tasks.size() <= (size_t)PACK_MAX) return false;
  for (size_t i = 0; i < tasks.size(); i += PACK_MAX)
    emit_one(std::vector<GemmTask>(tasks.begin() + i, tasks.begin() + std::min(tasks.size(), i + PACK_MAX)));
  return true;
}
// problems with an extent <= 16 (first / last MLP layers and their gradients) -> gemm_thin.cuh.  OSRL_THIN=0 keeps
// them on the tiled kernels (A/B timing, tests).
static int thin_kind(const GemmTask& t) {
  static const bool on = [] { const char* v = getenv("OSRL_THIN"); return !(v && v[0] == '0'); }();
  if (!on || gemm_mode() == "ffma") return THIN_NONE;
  if (t.act == ACT_GELU || t.dact == ACT_GELU || t.mmask) return THIN_NONE;
  if (t.K <= 16) return t.colsum ? THIN_NONE : THIN_K;
  if (!t.a_kc && !t.b_kc) {   // batch reduction: 16 k-groups per CTA walk K serially, long K stays on split-K tiles
    if (t.K > 4096) return THIN_NONE;
    if (t.N <= 16) return THIN_R_WIDE_M;
    if (t.M <= 16) return THIN_R_WIDE_N;
    return THIN_NONE;
  }
  if (t.N <= 16 && t.a_kc && !t.colsum) return THIN_N;
  return THIN_NONE;
}
// does `t` qualify as a packed-image producer (GemmTask::pk_*)?  Large-row thin-K forward layer whose column groups
// are whole consumers.  OSRL_PACK=0 disables the path (A/B timing, tests).
static bool pack_producer_ok(const GemmTask& t) {
  const char* v = getenv("OSRL_PACK");   // read at program build, like OSRL_GEMM
  const bool on = !(v && v[0] == '0');
  return on && (gemm_mode() == "tc5" || gemm_mode() == "fz") && t.thin == THIN_K && t.pk_gcols >= 64 && t.pk_gcols % 4 == 0 && t.M >= 512 &&
         t.N % t.pk_gcols == 0 && t.C != nullptr;
}
static void emit_thin(Engine& e, Program& p, std::vector<GemmTask> tasks) {
  if (split_packs(tasks, [&](std::vector<GemmTask> part) { emit_thin(e, p, part); })) return;
  int tot = 0;
  double bytes = 0.0, flops = 0.0;
  for (auto& t : tasks) {
    if (t.pk_gcols > 0 && pack_producer_ok(t)) {
      const int groups = t.N / t.pk_gcols, ks = (t.pk_gcols + 31) / 32, rbs = 2 * ((t.M + 127) / 128);
      const size_t per_group = (size_t)rbs * ks * 2048;
      OSRL_REQUIRE(per_group < ((size_t)1 << 31), "packed image too large");
      t.pk_ks = ks;
      t.pk_gstride = (int)per_group;
      t.pk_hi = e.ws(per_group * groups);   // zero-initialised: the row / k padding is never written
      t.pk_lo = e.ws(per_group * groups);
      e.pack_regs.push_back({t.C, t.ldc, t.M, t.N, t.pk_gcols, t.pk_gstride, ks, t.c_dead, t.pk_hi, t.pk_lo});
    } else {
      t.pk_gcols = 0; t.pk_hi = t.pk_lo = nullptr; t.c_dead = 0;
    }
    int tn = 1;
    t.klen = 16;   // THIN_N: rows per CTA (two per warp)
    t.a_vec = ((uintptr_t)t.A % 16 == 0) && (t.lda % 4 == 0) && ((t.a_kc ? t.K : t.M) % 4 == 0);
    t.b_vec = ((uintptr_t)t.B % 16 == 0) && (t.ldb % 4 == 0) && ((t.b_kc ? t.K : t.N) % 4 == 0);
    const int n = thin_tiles(t, t.thin, &tn);
    t.tile0 = tot; t.tiles_n = tn; t.tiles_mn = n; t.ksplit = 1;
    tot += n;
    bytes += 4.0 * ((double)t.M * t.K + (double)t.K * t.N + (double)t.M * t.N);
    flops += 2.0 * (double)t.M * t.N * t.K;
  }
  const TaskPack d = make_pack(tasks);
  const int nt = (int)tasks.size(), tiles = tot;
  Engine* ep = &e;
  p.add("k_gemm_thin", bytes, flops, true, [=](cudaStream_t s) {
    k_gemm_thin<<<tiles, THIN_THREADS, 0, s>>>(d, nt);
    ep->launches++;
  });
}
static bool tc5_eligible(const GemmTask& t) {
  return t.a_kc && t.b_kc && t.a_vec && t.b_vec && t.ksplit <= 1 && t.M >= 512 && t.K >= 64 && t.N >= 64;
}
static void emit_tc5(Engine& e, Program& p, std::vector<GemmTask> tasks, bool apack) {
  if (split_packs(tasks, [&](std::vector<GemmTask> part) { emit_tc5(e, p, part, apack); })) return;
  int tot = 0;
  double bytes = 0.0, flops = 0.0;
  bool full = false;
  const char* env = getenv("OSRL_TC5_BN");
  const int BNsel = (env && std::string(env) == "128" && !apack) ? 128 : 64;
  for (auto& t : tasks) {
    const int tm = (t.M + tc5::BM - 1) / tc5::BM, tn = (t.N + BNsel - 1) / BNsel;
    t.tile0 = tot; t.tiles_n = tn; t.tiles_mn = tm * tn;
    tot += tm * tn;
    bytes += 4.0 * ((double)t.M * t.K + (double)t.K * t.N + (double)t.M * t.N);
    flops += 2.0 * (double)t.M * t.N * t.K;
    full = full || t.act == ACT_GELU || t.dact == ACT_GELU || t.mmask != nullptr;
  }
  const TaskPack d = make_pack(tasks);
  const int nt = (int)tasks.size(), tiles = tot;
  Engine* ep = &e;
  p.add(apack ? "k_gemm_tc5<128,64,32,apack>" : (BNsel == 128 ? "k_gemm_tc5<128,128,32>" : "k_gemm_tc5<128,64,32>"), bytes,
        flops, true, [=](cudaStream_t s) {
    using S128 = tc5::Shape<128, 3>;
    using S64 = tc5::Shape<64, 2>;
    if (apack) {
      if (full) tc5::k_gemm_tc5<64, 2, 2, true, true><<<tiles, tc5::THREADS, S64::SMEM_BYTES, s>>>(d, nt);
      else tc5::k_gemm_tc5<64, 2, 2, false, true><<<tiles, tc5::THREADS, S64::SMEM_BYTES, s>>>(d, nt);
    } else if (BNsel == 128) {
      if (full) tc5::k_gemm_tc5<128, 3, 1, true><<<tiles, tc5::THREADS, S128::SMEM_BYTES, s>>>(d, nt);
      else tc5::k_gemm_tc5<128, 3, 1, false><<<tiles, tc5::THREADS, S128::SMEM_BYTES, s>>>(d, nt);
    } else {
      if (full) tc5::k_gemm_tc5<64, 2, 2, true><<<tiles, tc5::THREADS, S64::SMEM_BYTES, s>>>(d, nt);
      else tc5::k_gemm_tc5<64, 2, 2, false><<<tiles, tc5::THREADS, S64::SMEM_BYTES, s>>>(d, nt);
    }
    ep->launches++;
  });
}
using FzKern = void (*)(const FzPack, int);
static FzKern fz_kernel(int asrc, int bkc, int red) {
  switch (asrc * 4 + bkc * 2 + red) {
#define FZ_CASE(a, b, r) case a * 4 + b * 2 + r: return fz::k_fz<a, b, r>;
    FZ_CASE(0, 0, 0) FZ_CASE(0, 0, 1) FZ_CASE(0, 1, 0) FZ_CASE(0, 1, 1)
    FZ_CASE(1, 0, 0) FZ_CASE(1, 0, 1) FZ_CASE(1, 1, 0) FZ_CASE(1, 1, 1)
    FZ_CASE(2, 0, 0) FZ_CASE(2, 0, 1) FZ_CASE(2, 1, 0) FZ_CASE(2, 1, 1)
    FZ_CASE(3, 0, 0) FZ_CASE(3, 0, 1) FZ_CASE(3, 1, 0) FZ_CASE(3, 1, 1)
#undef FZ_CASE
  }
  return nullptr;
}
static int fz_variant(const FzTask& t) {
  const int asrc = t.a_gen == GEN_FIRST ? fz::A_FIRST : (t.a_gen == GEN_LASTD ? fz::A_LASTD : (t.a_kc ? fz::A_KC : fz::A_MC));
  return asrc * 4 + (t.b_kc ? 2 : 0) + (t.red ? 1 : 0);
}
void prepare_kernels() {
  // 227 KB per CTA minus the kernel's static shared memory (barriers + reduce-weight table, ~4.3 KB)
  for (int v = 0; v < 16; ++v)
    OSRL_CUDA(cudaFuncSetAttribute((const void*)fz_kernel(v >> 2, (v >> 1) & 1, v & 1),
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  prepare_gemm<OSRL_GEMM_CFG0>();
  prepare_gemm<OSRL_GEMM_CFG1>();
  prepare_gemm<OSRL_GEMM_CFG2>();
  prepare_mma<OSRL_MMA_CFG0>();
  prepare_mma<OSRL_MMA_CFG1>();
  prepare_mma<OSRL_MMA_CFG2>();
  prepare_mma<OSRL_MMA_CFG1K>();
  prepare_mma<OSRL_MMA_CFG2K>();
  OSRL_CUDA(cudaFuncSetAttribute(tc5::k_gemm_tc5<128, 3, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc5::Shape<128, 3>::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(tc5::k_gemm_tc5<128, 3, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc5::Shape<128, 3>::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(tc5::k_gemm_tc5<64, 2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc5::Shape<64, 2>::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(tc5::k_gemm_tc5<64, 2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc5::Shape<64, 2>::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(tc5::k_gemm_tc5<64, 2, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc5::Shape<64, 2>::SMEM_BYTES));
  OSRL_CUDA(cudaFuncSetAttribute(tc5::k_gemm_tc5<64, 2, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 tc5::Shape<64, 2>::SMEM_BYTES));
}
static int count_tiles(std::vector<GemmTask>& ts, int BM, int BN, bool assign) {
  int tot = 0;
  for (auto& t : ts) {
    const int tm = (t.M + BM - 1) / BM, tn = (t.N + BN - 1) / BN;
    if (assign) { t.tile0 = tot; t.tiles_n = tn; t.tiles_mn = tm * tn; }
    tot += tm * tn * (t.ksplit > 1 ? t.ksplit : 1);
  }
  return tot;
}

// ------------------------------------------------------------------ fused tcgen05 path (gemm_fz.cuh)
FzTask fz_blank() {
  FzTask t;
  memset(&t, 0, sizeof(t));
  t.scale = 1.f; t.rscale = 1.f; t.a_kc = 1; t.b_kc = 1; t.c_store = 1;
  return t;
}
// middle + last layer fusable (reduce epilogue / generated last-layer dgrad); fz_mlp_ok: the first layer too
static bool fz_mid_ok(const Lin& l0, const Lin& l1, const Lin& l2) {
  return fz_on() && l2.out <= fz::RED_MAX && l0.out == l1.in && l1.out == l2.in && l1.in % 4 == 0 && l1.out % 4 == 0;
}
bool fz_mlp_ok(const Lin& l0, const Lin& l1, const Lin& l2) { return fz_mid_ok(l0, l1, l2) && l0.in <= fz::GK_MAX; }
int fz_new_group() {
  static int next = 0;
  return ++next;
}
FzTask fz_from_gemm(const GemmTask& g) {
  FzTask t = fz_blank();
  t.A = g.A; t.B = g.B; t.C = g.C; t.M = g.M; t.N = g.N; t.K = g.K; t.lda = g.lda; t.ldb = g.ldb; t.ldc = g.ldc;
  t.a_kc = g.a_kc; t.b_kc = g.b_kc;
  t.bias = g.bias; t.resid = g.resid; t.dact_src = g.dact_src; t.aux = g.aux;
  t.ldr = g.ldr; t.ld_dact = g.ld_dact; t.ldaux = g.ldaux; t.act = g.act; t.clamp = g.clamp; t.dact = g.dact;
  t.scale = g.scale; t.lo = g.lo; t.hi = g.hi; t.colsum = g.colsum;
  t.ksplit = g.ksplit > 1 ? g.ksplit : 1; t.klen = g.klen;
  t.mmask = g.mmask; t.ldmm = g.ldmm;
  return t;
}
FzTask fz_fwd3(const float* X, int ldx, int rows, const float* W, const Lin& l0, const Lin& l1, const Lin& l2, int hact,
               float* h1, int ldh1, float* h2, int ldh2, float* out, int ldo) {
  FzTask t = fz_blank();
  t.a_gen = GEN_FIRST; t.gx = X; t.ldgx = ldx; t.gk = l0.in; t.gw = W + l0.w; t.gw_ld = l0.in; t.gb = W + l0.b;
  t.gact = hact; t.gstore = h1; t.ldgs = ldh1;
  t.B = W + l1.w; t.ldb = l1.in; t.b_kc = 1;
  t.M = rows; t.N = l1.out; t.K = l1.in;
  t.bias = W + l1.b; t.act = hact;
  t.C = h2; t.ldc = ldh2; t.c_store = h2 != nullptr;
  t.red = 1; t.red_n = l2.out; t.rw = W + l2.w; t.rs_j = l2.in; t.rs_n = 1; t.rbias = W + l2.b;
  t.r_group = fz_new_group(); t.r_slot0 = 0; t.r_slots = (l1.out + fz::BN - 1) / fz::BN;
  t.rout = out; t.ldro = ldo;
  return t;
}
FzTask fz_fwd2(const float* H1, int ldh1, int rows, const float* W, const Lin& l1, const Lin& l2, int hact, float* h2,
               int ldh2, float* out, int ldo) {
  FzTask t = fz_blank();
  t.A = H1; t.lda = ldh1; t.a_kc = 1;
  t.B = W + l1.w; t.ldb = l1.in; t.b_kc = 1;
  t.M = rows; t.N = l1.out; t.K = l1.in;
  t.bias = W + l1.b; t.act = hact;
  t.C = h2; t.ldc = ldh2; t.c_store = h2 != nullptr;
  t.red = 1; t.red_n = l2.out; t.rw = W + l2.w; t.rs_j = l2.in; t.rs_n = 1; t.rbias = W + l2.b;
  t.r_group = fz_new_group(); t.r_slot0 = 0; t.r_slots = (l1.out + fz::BN - 1) / fz::BN;
  t.rout = out; t.ldro = ldo;
  return t;
}
// large no-grad passes (targets on B*S rows): every column tile of a fused task regenerates the first layer (bound by
// broadcast shared-memory reads of its weights) and the kernel runs one CTA per SM, so on multi-wave launches the
// per-CTA prologue / epilogue is not hidden.  Those passes keep round 1's split -- thin first layer writing packed tf32
// images, gemm_tc5 (two CTAs per SM, bulk-copied A) for the middle layer, thin last layer.  Measured on B200 (BCQ-Lag
// B=256): actor_old on 5120 rows 57 us fused vs 39 us split; CPQ / BEAR-Lag at B=512 (5120-row passes twice per step)
// 0.46 / 0.56 ms fused vs 0.33 / 0.49 ms.  OSRL_FZ_UNFUSE=0 fuses them too (A/B timing).
bool fz_unfuse_first(int rows, bool nograd) {
  static const bool on = [] { const char* v = getenv("OSRL_FZ_UNFUSE"); return !(v && v[0] == '0'); }();
  return on && nograd && rows >= 1024;
}
FzTask fz_bwd_mid(const float* dq, int lddq, int rows, const float* W, const Lin& l1, const Lin& l2, int hact,
                  const float* h1, int ldh1, const float* h2, int ldh2, float* d1, int ldd1, float* d0, int ldd0) {
  FzTask t = fz_blank();
  t.a_gen = GEN_LASTD; t.gx = dq; t.ldgx = lddq; t.gk = l2.out; t.gw = W + l2.w; t.gw_ld = l2.in; t.gact = hact;
  t.gmask = h2; t.ldgm = ldh2; t.gstore = d1; t.ldgs = ldd1;
  t.B = W + l1.w; t.ldb = l1.in; t.b_kc = 0;          // d0[r, j] = sum_k d1[r, k] W1[k, j]
  t.M = rows; t.N = l1.in; t.K = l1.out;
  t.dact = hact; t.dact_src = h1; t.ld_dact = ldh1;
  t.C = d0; t.ldc = ldd0; t.c_store = d0 != nullptr;
  return t;
}
void fz_add_dx(FzTask& t, int group, int member, int members, const float* W, const Lin& l0, int col0, int ncols,
               float* dX, int lddx) {
  OSRL_REQUIRE(ncols <= fz::RED_MAX && t.N == l0.out, "fz_add_dx: bad first layer");
  const int tn = (t.N + fz::BN - 1) / fz::BN;
  t.red = 1; t.red_n = ncols; t.rw = W + l0.w + col0; t.rs_j = 1; t.rs_n = l0.in;
  t.r_group = group; t.r_slot0 = member * tn; t.r_slots = members * tn;
  t.rout = dX; t.ldro = lddx;
}
FzTask fz_wgrad(const float* dY, int lddy, const float* X, int ldx, int rows, float* Gsec, const Lin& l) {
  FzTask t = fz_blank();
  t.A = dY; t.lda = lddy; t.a_kc = 0;
  t.B = X; t.ldb = ldx; t.b_kc = 0;
  t.C = Gsec + l.w; t.ldc = l.in;
  t.M = l.out; t.N = l.in; t.K = rows;
  t.colsum = Gsec + l.b;
  return t;
}
void emit_fz(Engine& e, Program& p, std::vector<FzTask> tasks) {
  if (tasks.empty()) return;
  {   // one kernel variant per launch (gemm_fz.cuh): split the list by variant, keeping the order inside each
    const int v0 = fz_variant(tasks[0]);
    std::vector<FzTask> same, other;
    for (auto& t : tasks) (fz_variant(t) == v0 ? same : other).push_back(t);
    if (!other.empty()) {
      emit_fz(e, p, same);
      emit_fz(e, p, other);
      return;
    }
  }
  if (tasks.size() > (size_t)FZ_PACK) {
    for (size_t i = 0; i < tasks.size(); i += FZ_PACK)
      emit_fz(e, p, std::vector<FzTask>(tasks.begin() + i, tasks.begin() + std::min(tasks.size(), i + FZ_PACK)));
    return;
  }
  int tot = 0, genf = 0;
  double bytes = 0.0, flops = 0.0;
  bool fused = false;
  for (auto& t : tasks) {
    OSRL_REQUIRE(t.M > 0 && t.N > 0 && t.K > 0, "empty fused gemm task");
    const int tm = (t.M + fz::BM - 1) / fz::BM, tn = (t.N + fz::BN - 1) / fz::BN;
    t.tile0 = tot; t.tiles_n = tn; t.tiles_mn = tm * tn;
    if (t.ksplit < 1) t.ksplit = 1;
    if (t.ksplit > 1)
      OSRL_REQUIRE(t.klen % fz::BK == 0 && !t.red && t.a_gen == GEN_NONE && !t.bias && !t.dact && t.act == ACT_NONE,
                   "split-K needs a plain epilogue");
    tot += tm * tn * t.ksplit;
    if (t.a_gen == GEN_NONE) {
      OSRL_REQUIRE(t.A != nullptr, "fused gemm task without an A operand");
      t.a_vec = ((uintptr_t)t.A % 16 == 0) && (t.lda % 4 == 0) && ((t.a_kc ? t.K : t.M) % 4 == 0);
      OSRL_REQUIRE(t.a_vec, "fused gemm: A rows must be 16-byte aligned (bulk copies)");
    } else {
      OSRL_REQUIRE(t.gk >= 1 && t.gk <= fz::GK_MAX && t.K % 4 == 0, "A generation needs gk <= 16 and K % 4 == 0");
      if (t.gstore) OSRL_REQUIRE((uintptr_t)t.gstore % 16 == 0 && t.ldgs % 4 == 0, "gstore must be 16-byte aligned");
      if (t.a_gen == GEN_LASTD)
        OSRL_REQUIRE(t.gmask && (uintptr_t)t.gmask % 16 == 0 && t.ldgm % 4 == 0, "gmask must be 16-byte aligned");
      fused = true;
    }
    t.b_vec = ((uintptr_t)t.B % 16 == 0) && (t.ldb % 4 == 0) && ((t.b_kc ? t.K : t.N) % 4 == 0);
    OSRL_REQUIRE(t.b_vec && t.K % 4 == 0, "fused gemm: B rows must be 16-byte aligned (bulk copies)");
    t.epi_vec = (!t.c_store || ((uintptr_t)t.C % 16 == 0 && t.ldc % 4 == 0)) &&
                (!t.dact || ((uintptr_t)t.dact_src % 16 == 0 && t.ld_dact % 4 == 0));
    if (t.colsum) OSRL_REQUIRE(!t.a_kc && t.a_gen == GEN_NONE, "colsum needs an mn-contiguous A operand");
    if (t.red) {
      OSRL_REQUIRE(t.red_n >= 1 && t.red_n <= fz::RED_MAX && t.rout && t.r_group > 0 && t.r_slots >= tn, "bad reduce epilogue");
      Engine::FzAlloc* al = nullptr;
      for (auto& g : e.fz_groups)
        if (g.first == t.r_group) al = &g.second;
      if (!al) {
        Engine::FzAlloc a;
        a.rpart = e.ws((size_t)t.r_slots * t.M * t.red_n);
        a.rcnt = (unsigned*)e.ws((size_t)tm);
        e.fz_groups.push_back({t.r_group, a});
        al = &e.fz_groups.back().second;
      }
      t.rpart = al->rpart; t.rcnt = al->rcnt;
      fused = true;
    }
    genf = std::max(genf, fz::gen_floats(t));
    const double kin = t.a_gen == GEN_NONE ? (double)t.K : (double)t.gk;   // operands actually read from memory
    bytes += 4.0 * ((double)t.M * kin + (double)t.K * t.N + (t.c_store ? (double)t.M * t.N : 0.0) +
                    (t.gstore ? (double)t.M * t.K : 0.0) + (t.red ? (double)t.M * t.red_n : 0.0));
    flops += 2.0 * (double)t.M * t.N * t.K + 2.0 * (double)t.M * t.K * (t.a_gen ? t.gk : 0) +
             2.0 * (double)t.M * t.N * (t.red ? t.red_n : 0);
  }
  FzPack d;
  memset(&d, 0, sizeof(d));
  for (size_t i = 0; i < tasks.size(); ++i) { d.t[i] = tasks[i]; d.tile0[i] = tasks[i].tile0; }
  if (getenv("OSRL_FZ_DBG")) {   // kernel timeline (clock64 stamps per CTA), printed by osrl_debug_gemm
    d.dbg = (long long*)e.ws((size_t)tot * 64 * 2);
    e.fz_dbg = d.dbg; e.fz_dbg_ctas = tot;
  }
  const long long* dbg_buf = d.dbg;
  const int nt = (int)tasks.size(), tiles = tot;
  const int smem = fz::smem_fixed() + genf * 4;
  OSRL_REQUIRE(smem <= 200 * 1024, "fused kernel: generation tables do not fit shared memory");
  Engine* ep = &e;
  const int var = fz_variant(tasks[0]);
  const FzKern kern = fz_kernel(var >> 2, (var >> 1) & 1, var & 1);
  static const char* an[4] = {"a_kc", "a_mc", "first", "lastd"};
  const std::string name = std::string("k_fz<") + an[var >> 2] + ((var >> 1) & 1 ? ",b_kc" : ",b_nc") + (var & 1 ? ",red>" : ">");
  (void)fused;
  if (dbg_buf) e.fz_dbg_all.push_back({(long long*)dbg_buf, tot, name});
  p.add(name, bytes, flops, true, [=](cudaStream_t s) {
    kern<<<tiles, fz::THREADS, smem, s>>>(d, nt);
    ep->launches++;
  });
}
static void emit_tiled(Engine& e, Program& p, std::vector<GemmTask> tasks);
void emit_gemm(Engine& e, Program& p, const std::vector<GemmTask>& tasks_in) {
  if (tasks_in.empty()) return;
  std::vector<GemmTask> tasks;
  {
    std::vector<GemmTask> thin;
    std::vector<FzTask> fused;
    for (auto t : tasks_in) {
      if (t.fz_pending > 0) {   // the last layer of a fused network: move its epilogue into the reduce epilogue
        OSRL_REQUIRE(t.fz_pending <= (int)e.fz_pending.size(), "bad fused-network marker");
        FzTask f = e.fz_pending[t.fz_pending - 1];
        OSRL_REQUIRE(!t.dact && !t.colsum && !t.mmask, "unsupported epilogue on a fused last layer");
        f.ract = t.act; f.rscale = t.scale; f.rresid = t.resid; f.ldrr = t.ldr;
        f.rclamp = t.clamp; f.rlo = t.lo; f.rhi = t.hi; f.raux = t.aux; f.ldraux = t.ldaux;
        f.rout = t.C; f.ldro = t.ldc;
        fused.push_back(f);
        continue;
      }
      OSRL_REQUIRE(t.M > 0 && t.N > 0 && t.K > 0, "empty gemm task");
      t.thin = thin_kind(t);
      (t.thin ? thin : tasks).push_back(t);
    }
    if (!fused.empty()) emit_fz(e, p, fused);
    if (!thin.empty()) emit_thin(e, p, thin);
    if (tasks.empty()) return;
  }
  for (auto& t : tasks) {
    // 16-byte cp.async needs base, leading dimension and the contiguous extent 4-float aligned
    t.a_vec = ((uintptr_t)t.A % 16 == 0) && (t.lda % 4 == 0) && ((t.a_kc ? t.K : t.M) % 4 == 0);
    t.b_vec = ((uintptr_t)t.B % 16 == 0) && (t.ldb % 4 == 0) && ((t.b_kc ? t.K : t.N) % 4 == 0);
    // split-K for weight gradients over many rows (CDT: K = 81,920 tokens onto a 128x384 tile grid): the
    // splits add their partials atomically into a pre-zeroed C (contiguous [M, N]) and bias gradient
    t.ksplit = 1;
    t.klen = (t.K + 63) / 64 * 64;
    const bool plain = !t.bias && !t.resid && !t.dact && !t.aux && !t.clamp && t.act == ACT_NONE && t.scale == 1.f;
    if (plain && t.K >= 8192 && t.ldc == t.N) {
      // (fused tcgen05 kernel: 128 x 64 tiles, one CTA per SM -> two waves' worth of splits)
      const int tiles = fz_on() ? ((t.M + 127) / 128) * ((t.N + 63) / 64) : ((t.M + 63) / 64) * ((t.N + 63) / 64);
      int want = std::max(1, (fz_on() ? 296 : 444) / std::max(1, tiles));
      want = std::min(want, t.K / 1024);
      if (want > 1) {
        t.klen = ((t.K + want - 1) / want + 63) / 64 * 64;
        t.ksplit = (t.K + t.klen - 1) / t.klen;
      }
    }
  }
  for (auto& t : tasks)
    if (t.ksplit > 1) {   // zero the accumulation targets ahead of the launch
      float* C = t.C;
      float* cs = t.colsum;
      const size_t cb = (size_t)t.M * t.N * sizeof(float), sb = (size_t)t.M * sizeof(float);
      p.add("memset", 0.0, 0.0, false, [=](cudaStream_t s) {
        cudaMemsetAsync(C, 0, cb, s);
        if (cs) cudaMemsetAsync(cs, 0, sb, s);
      });
    }
  if (fz_on()) {   // every tiled problem the fused kernel's plain mode covers goes to tcgen05
    std::vector<FzTask> fzt;
    std::vector<GemmTask> rest, packed;
    for (auto& t : tasks) {
      // large no-grad passes keep round 1's split: thin first layer -> packed tf32 images -> gemm_tc5 (two CTAs per SM)
      const Engine::PackReg* reg = nullptr;
      for (auto& r : e.pack_regs) {
        const ptrdiff_t off = t.A - r.C;
        if (t.a_kc && off >= 0 && off < r.N && t.lda == r.ldc && t.M == r.M) reg = &r;
      }
      if (reg) {
        OSRL_REQUIRE(tc5_eligible(t) && t.K == reg->gcols && (t.A - reg->C) % reg->gcols == 0,
                     "a GEMM reads activations that only exist as packed images");
        const int grp = (int)((t.A - reg->C) / reg->gcols);
        t.a_hi = reg->hi + (size_t)grp * reg->gstride;
        t.a_lo = reg->lo + (size_t)grp * reg->gstride;
        t.pk_ks = reg->ks;
        packed.push_back(t);
        continue;
      }
      // (operand rows are fetched with 16-byte copies: unaligned layouts, e.g. K = 41 first layers, stay on mma.sync)
      if ((t.colsum && t.a_kc) || !t.a_vec || !t.b_vec || t.K % 4 != 0) rest.push_back(t);
      else fzt.push_back(fz_from_gemm(t));
    }
    if (!packed.empty()) emit_tc5(e, p, packed, true);
    emit_fz(e, p, fzt);
    {   // GELU-forward / dropout-multiplier epilogues (CDT): round 1's tcgen05 kernel; only unaligned leftovers see mma.sync
      std::vector<GemmTask> big, small;
      for (auto& t : rest) (tc5_eligible(t) ? big : small).push_back(t);
      if (!big.empty()) emit_tc5(e, p, big, false);
      rest = small;
    }
    if (rest.empty()) return;
    tasks = rest;
  } else if (gemm_mode() == "tc5") {   // large forward layers -> tcgen05 kernel, the rest stays on mma.sync
    std::vector<GemmTask> big, packed, rest;
    for (auto& t : tasks) {
      const Engine::PackReg* reg = nullptr;   // were these activations announced as packed hi/lo images?
      for (auto& r : e.pack_regs) {
        const ptrdiff_t off = t.A - r.C;
        if (t.a_kc && off >= 0 && off < r.N && t.lda == r.ldc && t.M == r.M) reg = &r;
      }
      if (reg && tc5_eligible(t) && t.K == reg->gcols && (t.A - reg->C) % reg->gcols == 0) {
        const int grp = (int)((t.A - reg->C) / reg->gcols);
        t.a_hi = reg->hi + (size_t)grp * reg->gstride;
        t.a_lo = reg->lo + (size_t)grp * reg->gstride;
        t.pk_ks = reg->ks;
        packed.push_back(t);
        continue;
      }
      OSRL_REQUIRE(!(reg && reg->dead), "a GEMM reads activations that only exist as packed images");
      (tc5_eligible(t) ? big : rest).push_back(t);
    }
    if (!packed.empty()) emit_tc5(e, p, packed, true);
    if (!big.empty()) emit_tc5(e, p, big, false);
    if (rest.empty()) return;
    tasks = rest;
  }
  emit_tiled(e, p, tasks);
}
static void emit_tiled(Engine& e, Program& p, std::vector<GemmTask> tasks) {
  if (split_packs(tasks, [&](std::vector<GemmTask> part) { emit_tiled(e, p, part); })) return;
  // largest tile shape that still yields >= ~1 wave-fraction of CTAs (148 SMs)
  int cfg = 2;
  if (count_tiles(tasks, 128, 64, false) >= 120) cfg = 0;
  else if (count_tiles(tasks, 64, 64, false) >= 96) cfg = 1;
  static const int bm[3] = {128, 64, 32}, bn[3] = {64, 64, 32};
  const int tiles = count_tiles(tasks, bm[cfg], bn[cfg], true);
  const TaskPack d = make_pack(tasks);
  const int nt = (int)tasks.size();
  Engine* ep = &e;
  double bytes = 0.0, flops = 0.0;
  for (auto& t : tasks) {
    bytes += 4.0 * ((double)t.M * t.K + (double)t.K * t.N + (double)t.M * t.N);
    flops += 2.0 * (double)t.M * t.N * t.K;
  }
  static const char* names[3] = {"k_gemm_tasks<128,64,16,8,4,4>", "k_gemm_tasks<64,64,32,4,4,4>",
                                 "k_gemm_tasks<32,32,32,2,2,6>"};
  static const char* mnames[3] = {"k_gemm_mma<128,64,16,4,2,4>", "k_gemm_mma<64,64,32,2,4,4>",
                                  "k_gemm_mma<32,32,32,2,2,6>"};
  const bool mma = use_mma();
  int kmax = 0;
  for (auto& t : tasks) kmax = std::max(kmax, t.ksplit > 1 ? t.klen : t.K);
  const bool kgroups = cfg > 0 && kmax <= 1024 && tiles <= 2 * 148;
  bool full = false;
  for (auto& t : tasks) full = full || t.ksplit > 1 || t.act == ACT_GELU || t.dact == ACT_GELU || t.mmask != nullptr;
  p.add(mma ? mnames[cfg] : names[cfg], bytes, flops, true, [=](cudaStream_t s) {
#define OSRL_DISPATCH(FULL_)                                                      \
    if (mma) {                                                                    \
      if (cfg == 0) launch_mma<OSRL_MMA_CFG0, FULL_>(d, nt, tiles, s);            \
      else if (cfg == 1 && kgroups) launch_mma<OSRL_MMA_CFG1K, FULL_>(d, nt, tiles, s); \
      else if (cfg == 1) launch_mma<OSRL_MMA_CFG1, FULL_>(d, nt, tiles, s);       \
      else if (kgroups) launch_mma<OSRL_MMA_CFG2K, FULL_>(d, nt, tiles, s);       \
      else launch_mma<OSRL_MMA_CFG2, FULL_>(d, nt, tiles, s);                     \
    } else {                                                                      \
      if (cfg == 0) launch_gemm<OSRL_GEMM_CFG0, FULL_>(d, nt, tiles, s);          \
      else if (cfg == 1) launch_gemm<OSRL_GEMM_CFG1, FULL_>(d, nt, tiles, s);     \
      else launch_gemm<OSRL_GEMM_CFG2, FULL_>(d, nt, tiles, s);                   \
    }
    if (full) { OSRL_DISPATCH(true) } else { OSRL_DISPATCH(false) }
#undef OSRL_DISPATCH
    ep->launches++;
  });
}
CopyTask copy_cols(float* dst, int ldd, int dcol0, const float* src, int lds, int scol0, int rows, int cols, int row_div,
                   int row_mod) {
  CopyTask t;
  memset(&t, 0, sizeof(t));
  t.dst = dst + dcol0; t.ldd = ldd;
  t.src = src + scol0; t.lds = lds;
  t.rows = rows; t.cols = cols;
  t.row_div = row_div; t.row_mod = row_mod;
  t.mul = 1.f;
  return t;
}
void emit_copy(Engine& e, Program& p, const std::vector<CopyTask>& tasks) {
  if (tasks.empty()) return;
  CopyTask* d = e.upload(tasks);
  long long mx = 1;
  for (auto& t : tasks) mx = std::max(mx, (long long)t.rows * t.cols);
  const int bx = (int)std::min<long long>((mx + 255) / 256, 148 * 4);
  const int ny = (int)tasks.size();
  Engine* ep = &e;
  double bytes = 0.0;
  for (auto& t : tasks) bytes += 8.0 * (double)t.rows * t.cols;
  p.add("k_copy_tasks", bytes, 0.0, true, [=](cudaStream_t s) {
    k_copy_tasks<<<dim3(bx, ny), 256, 0, s>>>(d);
    ep->launches++;
  });
}
void emit_adam(Engine& e, Program& p, int group, int64_t begin, int64_t end, bool polyak, const float* clip_coef) {
  const Group& g = e.plan.groups[group];
  const int64_t n4 = (end - begin) / 4;
  const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 148 * 8);
  Engine* ep = &e;
  const float tau = e.plan.cfg.tau;
  // read p,g,m,v + write p,m,v = 28 B/param; Polyak adds read+write of the target = 8 B/param
  const double bytes = (double)(end - begin) * (polyak ? 36.0 : 28.0);
  // data parallel, peer mode: the all-reduce of this range (the DP_GRAD emit_allreduce just before) happens here
  int slot = -1;
  if (e.world > 1 && !clip_coef) {
    slot = (int)e.dp_slot_group.size();
    OSRL_REQUIRE(slot < DP_MAX_SLOT, "too many data-parallel reduce sites");
    e.dp_slot_group.push_back(group);
  }
  p.add(polyak ? "k_adam+polyak" : "k_adam", bytes, 0.0, true, [=](cudaStream_t s) {
    if (ep->peer_on && slot >= 0)
      // <= 148 blocks: with the two graph branches' reduce kernels both resident no waiting block can keep a
      // signalling block off the machine
      k_dp_adam<<<std::min(blocks, 148), 256, 0, s>>>(ep->peers, slot, begin, ep->P + begin, ep->M + begin, ep->V + begin,
                                                      ep->T + begin, n4, ep->ds, group, (float)g.beta1, (float)g.beta2,
                                                      (float)(1.0 - g.beta1), (float)(1.0 - g.beta2), g.eps, g.wd, tau,
                                                      polyak ? 1 : 0);
    else
      k_adam<<<blocks, 256, 0, s>>>(ep->P + begin, ep->G + begin, ep->M + begin, ep->V + begin, ep->T + begin, n4, ep->ds,
                                    group, (float)g.beta1, (float)g.beta2, (float)(1.0 - g.beta1), (float)(1.0 - g.beta2),
                                    g.eps, g.wd, tau, polyak ? 1 : 0, 1.f, clip_coef);
    ep->launches++;
  });
}

// ------------------------------------------------------------------ NCCL (dlopen'ed; only needed when world > 1)
namespace nccl {
struct Uid { char internal[128]; };  // ncclUniqueId
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*CommDestroy_t)(void*);
typedef const char* (*GetErrorString_t)(int);
static void* lib = nullptr;
static int (*GetUniqueId)(Uid*) = nullptr;
static int (*CommInitRank)(void**, int, Uid, int) = nullptr;
static AllReduce_t AllReduce = nullptr;
static CommDestroy_t CommDestroy = nullptr;
static GetErrorString_t GetErrorString = nullptr;
static int (*CommSplit)(void*, int, int, void**, void*) = nullptr;   // optional (NCCL >= 2.18)
static int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;   // optional: peer-memory setup
static void load() {
  if (lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) throw Err(OSRL_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
  GetUniqueId = (int (*)(Uid*))dlsym(lib, "ncclGetUniqueId");
  CommInitRank = (int (*)(void**, int, Uid, int))dlsym(lib, "ncclCommInitRank");
  AllReduce = (AllReduce_t)dlsym(lib, "ncclAllReduce");
  CommDestroy = (CommDestroy_t)dlsym(lib, "ncclCommDestroy");
  GetErrorString = (GetErrorString_t)dlsym(lib, "ncclGetErrorString");
  CommSplit = (int (*)(void*, int, int, void**, void*))dlsym(lib, "ncclCommSplit");
  AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
  if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) throw Err(OSRL_ERR_NCCL, "libnccl lacks required symbols");
}
static void check(int r, const char* what) {
  if (r != 0) throw Err(OSRL_ERR_NCCL, std::string(what) + ": " + (GetErrorString ? GetErrorString(r) : "nccl error"));
}
}  // namespace nccl

void emit_allreduce(Engine& e, Program& p, float* buf, int64_t count, bool f64, int kind) {
  if (e.world <= 1) return;
  Engine* ep = &e;
  const bool side = (&p == &e.pa);   // the pipelined VAE branch reduces on its own communicator: the two branches'
                                     // collectives may then be in flight at the same time
  int sslot = -1;
  if (kind == DP_SCALAR) {
    OSRL_REQUIRE(!f64 && count <= DP_SCAL_N && e.dp_scal_slots < DP_MAX_SLOT, "scalar exchange: at most 8 fp32 values, 32 sites");
    sslot = e.dp_scal_slots++;
  }
  p.add("allreduce", 4.0 * (double)count, 0.0, false, [=](cudaStream_t s) {
    if (ep->peer_on && kind == DP_GRAD) return;   // fused into the Adam launch that follows
    if (ep->peer_on && kind == DP_SCALAR) {
      k_dp_scalar<<<1, 32, 0, s>>>(ep->peers, sslot, buf, (int)count);
      ep->launches++;
      return;
    }
    void* comm = side ? ep->comm2 : ep->comm;
    if (!comm) throw Err(OSRL_ERR_STATE, "world_size > 1 but osrl_comm_init was not called");
    nccl::check(nccl::AllReduce(buf, buf, (size_t)count, f64 ? /*ncclFloat64*/ 8 : /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, s),
                "ncclAllReduce");
  });
}

// ------------------------------------------------------------------ ensemble helpers
EnsBuf ens_alloc(Engine& e, const EnsLay& l, int rows) {
  EnsBuf b;
  b.rows = rows;
  for (int hw : l.h) b.h.push_back(e.ws((size_t)rows * l.n * hw));
  b.q = e.ws((size_t)rows * l.n);
  return b;
}
// member i of an ensemble as three plain layers (fused path)
static void ens_member(const EnsLay& l, int i, Lin& l0, Lin& l1, Lin& l2) {
  l0.in = l.in; l0.out = l.h[0];
  l0.w = l.first.w + (int64_t)i * l.h[0] * l.in;
  l0.b = l.first.b + (int64_t)i * l.h[0];
  l1 = l.mid[0][i];
  l2.in = l.h[1]; l2.out = 1;
  l2.w = l.w_last + (int64_t)i * l.h[1];
  l2.b = l.b_last + i;
}
static bool ens_fz_ok(const EnsLay& l) {
  if (!fz_on() || l.h.size() != 2) return false;
  Lin l0, l1, l2;
  ens_member(l, 0, l0, l1, l2);
  return fz_mlp_ok(l0, l1, l2);
}
static bool ens_fz_mid_ok(const EnsLay& l) {   // wide first layers (CPQ: obs 33 + act 8): only the first layer stays apart
  if (!fz_on() || l.h.size() != 2) return false;
  Lin l0, l1, l2;
  ens_member(l, 0, l0, l1, l2);
  return fz_mid_ok(l0, l1, l2);
}
void ens_fwd(std::vector<Stage>& st, const EnsLay& l, const float* W, const float* X, int ldx, int rows, EnsBuf& buf,
             bool nograd) {
  const int nh = (int)l.h.size();
  OSRL_REQUIRE((int)st.size() >= nh + 1, "ens_fwd: not enough stages");
  if (!ens_fz_ok(l) && ens_fz_mid_ok(l) && !fz_unfuse_first(rows, nograd)) {
    // first layer as one stacked GEMM, then per member: middle layer with the Q head in its reduce epilogue
    st[0].tasks.push_back(task_fwd(X, ldx, rows, W, l.first, buf.h[0], l.n * l.h[0], ACT_RELU));
    for (int i = 0; i < l.n; ++i) {
      Lin l0, l1, l2;
      ens_member(l, i, l0, l1, l2);
      st[1].fz.push_back(fz_fwd2(buf.h[0] + (size_t)i * l.h[0], l.n * l.h[0], rows, W, l1, l2, ACT_RELU,
                                 nograd ? nullptr : buf.h[1] + (size_t)i * l.h[1], l.n * l.h[1], buf.q + i, l.n));
    }
    return;
  }
  if (ens_fz_ok(l) && !fz_unfuse_first(rows, nograd)) {   // one fused launch: first layer generated, middle layer on tcgen05, Q head in the reduce epilogue
    for (int i = 0; i < l.n; ++i) {
      Lin l0, l1, l2;
      ens_member(l, i, l0, l1, l2);
      st[0].fz.push_back(fz_fwd3(X, ldx, rows, W, l0, l1, l2, ACT_RELU,
                                 nograd ? nullptr : buf.h[0] + (size_t)i * l.h[0], l.n * l.h[0],
                                 nograd ? nullptr : buf.h[1] + (size_t)i * l.h[1], l.n * l.h[1], buf.q + i, l.n));
    }
    return;
  }
  GemmTask first = task_fwd(X, ldx, rows, W, l.first, buf.h[0], l.n * l.h[0], ACT_RELU);
  if (nograd && nh > 1) { first.pk_gcols = l.h[0]; first.c_dead = 1; }
  st[0].tasks.push_back(first);
  for (int k = 1; k < nh; ++k)
    for (int i = 0; i < l.n; ++i)
      st[k].tasks.push_back(task_fwd(buf.h[k - 1] + (size_t)i * l.h[k - 1], l.n * l.h[k - 1], rows, W, l.mid[k - 1][i],
                                     buf.h[k] + (size_t)i * l.h[k], l.n * l.h[k], ACT_RELU));
  for (int i = 0; i < l.n; ++i) {
    Lin last;
    last.in = l.h.back(); last.out = 1;
    last.w = l.w_last + (int64_t)i * l.h.back();
    last.b = l.b_last + i;
    st[nh].tasks.push_back(task_fwd(buf.h[nh - 1] + (size_t)i * l.h.back(), l.n * l.h.back(), rows, W, last,
                                    buf.q + i, l.n, ACT_NONE));
  }
}
void ens_bwd(std::vector<Stage>& st, const EnsLay& l, const float* W, float* Gsec, const float* X, int ldx, int rows,
             const EnsBuf& act, EnsBuf& grad, const float* dq, float* dX, int lddx, int xcol0, int xcols) {
  const int nh = (int)l.h.size();
  OSRL_REQUIRE((int)st.size() >= nh + 1, "ens_bwd: not enough stages");
  if (ens_fz_mid_ok(l) && (!dX || xcols <= fz::RED_MAX)) {
    // stage 0: per member, last-layer dgrad generated -> middle-layer dgrad on tcgen05 -> (input gradient summed over
    // the ensemble in the reduce epilogue); stage 1: weight gradients (middle layers on tcgen05, thin ones beside)
    const int grp = dX ? fz_new_group() : 0;
    for (int i = 0; i < l.n; ++i) {
      Lin l0, l1, l2;
      ens_member(l, i, l0, l1, l2);
      FzTask t = fz_bwd_mid(dq + i, l.n, rows, W, l1, l2, ACT_RELU, act.h[0] + (size_t)i * l.h[0], l.n * l.h[0],
                            act.h[1] + (size_t)i * l.h[1], l.n * l.h[1],
                            Gsec ? grad.h[1] + (size_t)i * l.h[1] : nullptr, l.n * l.h[1],
                            Gsec ? grad.h[0] + (size_t)i * l.h[0] : nullptr, l.n * l.h[0]);
      if (dX) fz_add_dx(t, grp, i, l.n, W, l0, xcol0, xcols, dX, lddx);
      st[0].fz.push_back(t);
      if (Gsec) {
        st[1].fz.push_back(fz_wgrad(grad.h[1] + (size_t)i * l.h[1], l.n * l.h[1], act.h[0] + (size_t)i * l.h[0],
                                    l.n * l.h[0], rows, Gsec, l1));
        st[1].tasks.push_back(task_wgrad(dq + i, l.n, act.h[1] + (size_t)i * l.h[1], l.n * l.h[1], rows, Gsec, l2));
      }
    }
    if (Gsec) st[1].tasks.push_back(task_wgrad(grad.h[0], l.n * l.h[0], X, ldx, rows, Gsec, l.first));
    return;
  }
  // last layer (1 unit): dH = dq (x) w_last, masked by relu'(H)
  for (int i = 0; i < l.n; ++i) {
    Lin last;
    last.in = l.h.back(); last.out = 1;
    last.w = l.w_last + (int64_t)i * l.h.back();
    last.b = l.b_last + i;
    const int ldh = l.n * l.h.back();
    if (Gsec) st[0].tasks.push_back(task_wgrad(dq + i, l.n, act.h[nh - 1] + (size_t)i * l.h.back(), ldh, rows, Gsec, last));
    st[0].tasks.push_back(task_dgrad(dq + i, l.n, rows, W, last, grad.h[nh - 1] + (size_t)i * l.h.back(), ldh,
                                     act.h[nh - 1] + (size_t)i * l.h.back(), ldh, ACT_RELU));
  }
  for (int k = nh - 1; k >= 1; --k) {
    const int s = nh - k;
    for (int i = 0; i < l.n; ++i) {
      const Lin& m = l.mid[k - 1][i];
      const int ldo = l.n * l.h[k], ldi = l.n * l.h[k - 1];
      if (Gsec)
        st[s].tasks.push_back(task_wgrad(grad.h[k] + (size_t)i * l.h[k], ldo, act.h[k - 1] + (size_t)i * l.h[k - 1], ldi,
                                         rows, Gsec, m));
      st[s].tasks.push_back(task_dgrad(grad.h[k] + (size_t)i * l.h[k], ldo, rows, W, m,
                                       grad.h[k - 1] + (size_t)i * l.h[k - 1], ldi,
                                       act.h[k - 1] + (size_t)i * l.h[k - 1], ldi, ACT_RELU));
    }
  }
  if (Gsec) st[nh].tasks.push_back(task_wgrad(grad.h[0], l.n * l.h[0], X, ldx, rows, Gsec, l.first));
  if (dX) st[nh].tasks.push_back(task_dgrad(grad.h[0], l.n * l.h[0], rows, W, l.first, dX, lddx, nullptr, 0, 0, xcol0, xcols));
}

// ------------------------------------------------------------------ engine life cycle
// dropout probability of a noise slot ("drop_*" slots hold multipliers, see osrl_noise in the header); 0 = normals
static float slot_drop_p(const osrl_config& c, const std::string& name) {
  if (name.rfind("drop_emb", 0) == 0) return c.embedding_dropout;
  if (name.rfind("drop_attn", 0) == 0) return c.attention_dropout;
  if (name.rfind("drop_res", 0) == 0) return c.residual_dropout;
  return 0.f;
}
// a re-upload (curriculum, re-scaled rewards) replaces the resident copy: drop the old one instead of keeping it until
// osrl_engine_destroy.  Called after a device synchronisation.
static void release_alloc(Engine& e, void* p) {
  if (!p) return;
  auto it = std::find(e.allocs.begin(), e.allocs.end(), p);
  if (it != e.allocs.end()) e.allocs.erase(it);
  cudaFree(p);
}
static void drop_sampled_graphs(Engine& e) {   // they bake the dataset pointers
  for (cudaGraphExec_t* g : {&e.g_sampled, &e.g_pro, &e.g_mid, &e.g_last, &e.g_xbody, &e.g_xpro, &e.g_xmid, &e.g_xlast})
    if (*g) { cudaGraphExecDestroy(*g); *g = nullptr; }
}
static void free_all(Engine* e) {
  if (e->peer_on) {   // peers may still be summing this rank's last gradients out of its memory
    k_dp_drain<<<1, 32>>>(e->d_peers);
    cudaDeviceSynchronize();
  }
  if (e->g_body) cudaGraphExecDestroy(e->g_body);
  drop_sampled_graphs(*e);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (float* hp : {e->x_ring, e->x_st_side, e->x_st_main})
    if (hp) cudaFreeHost(hp);
  for (void* q : e->ipc_open) cudaIpcCloseMemHandle(q);
  for (int i = 0; i < 2; ++i) {
    if (e->stats_pinned[i]) cudaFreeHost(e->stats_pinned[i]);
    if (e->stats_ev[i]) cudaEventDestroy(e->stats_ev[i]);
  }
  for (auto& ps : e->par_streams)
    for (cudaStream_t x : ps.second) cudaStreamDestroy(x);
  for (cudaEvent_t ev : e->par_events) cudaEventDestroy(ev);
  for (void* p : e->allocs) {
    // Exported over cudaIpc: a peer process may not have closed its mapping yet (engines are destroyed without a
    // cross-rank barrier -- a collective in a destructor can deadlock on destruction order), and freeing memory that is
    // still imported elsewhere is undefined.  The two exported blocks (gradient section, 21 KB of flags) stay allocated
    // until the process exits.
    if (e->peer_on && (p == (void*)e->G || p == (void*)e->dp_flags)) continue;
    cudaFree(p);
  }
  if (e->comm2 && nccl::CommDestroy) nccl::CommDestroy(e->comm2);
  if (e->comm && nccl::CommDestroy) nccl::CommDestroy(e->comm);
}

// Second build of the step as two concurrent halves (see Engine::pa / pm).  Data parallel: the VAE branch all-reduces
// on a second communicator (ncclCommSplit of the first, osrl_comm_init).  OSRL_PIPELINE=0 keeps osrl_steps() sequential.
static void build_pipelined(Engine& e, const std::vector<int>& vae_slots) {
  const char* v = getenv("OSRL_PIPELINE");
  if (v && v[0] == '0') return;
  const osrl_config& c = e.plan.cfg;
  const int o = c.obs_dim, a = c.act_dim, B = e.B;
  e.nb_obs = e.ws((size_t)B * o); e.nb_nobs = e.ws((size_t)B * o); e.nb_act = e.ws((size_t)B * a);
  e.nb_rew = e.ws(B); e.nb_cost = e.ws(B); e.nb_done = e.ws(B);
  e.nb_idx = (int64_t*)e.ws((size_t)B * 2);
  e.Psnap = e.ws((size_t)e.plan.nP);
  std::vector<NoiseSlot> sv, sr;
  for (int i = 0; i < (int)e.noise_buf.size(); ++i) {
    const bool is_vae = std::find(vae_slots.begin(), vae_slots.end(), i) != vae_slots.end();
    NoiseSlot s{e.noise_buf[i], (long long)e.plan.noise[i].second, i, 1, slot_drop_p(c, e.plan.noise[i].first)};
    NoiseSlot off = s;
    off.enabled = 0;
    sv.push_back(is_vae ? s : off);
    sr.push_back(is_vae ? off : s);
  }
  e.d_slots_vae = e.upload(sv);
  e.d_slots_rest = e.upload(sr);
  OSRL_CUDA(cudaStreamCreateWithFlags(&e.side_stream, cudaStreamNonBlocking));
  OSRL_CUDA(cudaEventCreateWithFlags(&e.ev_fork, cudaEventDisableTiming));
  OSRL_CUDA(cudaEventCreateWithFlags(&e.ev_join, cudaEventDisableTiming));
  switch (c.algo) {
    case OSRL_ALGO_BCQL: build_bcql(e, 1); build_bcql(e, 2); break;
    case OSRL_ALGO_CPQ: build_cpq(e, 1); build_cpq(e, 2); break;
    case OSRL_ALGO_BEARL: build_bearl(e, 1); build_bearl(e, 2); break;
    default: return;
  }
  e.pipelined = true;
  e.side_stat_mask = 1u;   // "loss/loss_vae" (stat 0 of BCQ-Lag, CPQ and BEAR-Lag) is the VAE branch's
}
static void build_program(Engine& e) {
  switch (e.plan.cfg.algo) {
    case OSRL_ALGO_BC: build_bc(e); break;
    case OSRL_ALGO_BCQL: {
      build_bcql(e);
      const char* v = getenv("OSRL_PIPELINE_DECODE");   // VAE branch also runs the step's two VAE decodes
      e.decode_side = v ? v[0] != '0' : true;
      if (e.decode_side) build_pipelined(e, {0, 1, 2, 3});
      else build_pipelined(e, {0});
      break;
    }
    case OSRL_ALGO_CPQ: build_cpq(e); build_pipelined(e, {0}); break;
    case OSRL_ALGO_BEARL: build_bearl(e); build_pipelined(e, {0}); break;
    case OSRL_ALGO_CDT: build_cdt(e); break;
    case OSRL_ALGO_COPTIDICE: build_coptidice(e); break;
    default: throw Err(OSRL_ERR_UNSUPPORTED, "algorithm not supported");
  }
}

static Engine* create(const osrl_config& cfg, int device) {
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    throw Err(OSRL_ERR_CUDA, std::string("no CUDA device available (osrl_b200 has no CPU fallback): ") +
                                 cudaGetErrorString(ce));
  OSRL_REQUIRE(device >= 0 && device < ndev, "bad device ordinal");
  OSRL_REQUIRE(cfg.batch_size > 0, "batch_size must be positive");
  OSRL_CUDA(cudaSetDevice(device));
  Engine* e = new Engine();
  try {
    e->plan = make_plan(cfg);
    e->device = device;
    e->world = cfg.world_size;
    e->rank = cfg.rank;
    e->B = cfg.batch_size;
    const size_t n = (size_t)e->plan.nP;
    e->P = e->ws(n); e->T = e->ws(n); e->G = e->ws(n); e->M = e->ws(n); e->V = e->ws(n);
    void* dsp = nullptr;
    OSRL_CUDA(cudaMalloc(&dsp, sizeof(DevState)));
    OSRL_CUDA(cudaMemset(dsp, 0, sizeof(DevState)));
    e->allocs.push_back(dsp);
    e->ds = (DevState*)dsp;
    if (cfg.algo == OSRL_ALGO_CDT) {
      DevState h;
      memset(&h, 0, sizeof(h));
      h.log_temperature = log((double)cfg.init_temperature);
      OSRL_CUDA(cudaMemcpy(e->ds, &h, sizeof(h), cudaMemcpyHostToDevice));
    }
    std::vector<AdamGroupCfg> gc;
    for (auto& g : e->plan.groups) gc.push_back({(double)g.lr, g.beta1, g.beta2, g.warmup});
    e->d_groups = e->upload(gc);
    const int o = cfg.obs_dim, a = cfg.act_dim, B = e->B;
    e->b_obs = e->ws((size_t)B * o); e->b_nobs = e->ws((size_t)B * o); e->b_act = e->ws((size_t)B * a);
    e->b_rew = e->ws(B); e->b_cost = e->ws(B); e->b_done = e->ws(B);
    e->b_idx = (int64_t*)e->ws((size_t)B * 2);
    if (cfg.algo == OSRL_ALGO_COPTIDICE) {
      OSRL_REQUIRE(cfg.observations_std && cfg.actions_std, "COptiDICE needs observations_std / actions_std");
      e->b_init = e->ws(B);
      e->cop_obs_std = e->upload(std::vector<float>(cfg.observations_std, cfg.observations_std + o));
      e->cop_act_std = e->upload(std::vector<float>(cfg.actions_std, cfg.actions_std + a));
      DevState hst;
      memset(&hst, 0, sizeof(hst));
      hst.cop_tau = hst.cop_lmbda = 1.f;   // torch.ones(1) (coptidice.py:104-105)
      OSRL_CUDA(cudaMemcpy(e->ds, &hst, sizeof(hst), cudaMemcpyHostToDevice));
      e->plan.cfg.observations_std = e->plan.cfg.actions_std = nullptr;   // (host pointers are not kept)
    }
    if (cfg.algo == OSRL_ALGO_CDT) {
      const size_t BT = (size_t)B * cfg.seq_len;
      e->s_states = e->ws(BT * o); e->s_actions = e->ws(BT * a); e->s_returns = e->ws(BT); e->s_ctg = e->ws(BT);
      e->s_mask = e->ws(BT); e->s_costs = e->ws(BT);
      e->s_ts = (long long*)e->ws(BT * 2);
      e->s_traj = (int*)e->ws(B); e->s_start = (int*)e->ws(B);
    }
    std::vector<NoiseSlot> slots;
    int si = 0;
    for (auto& ns : e->plan.noise) {
      float* buf = e->ws((size_t)ns.second);
      e->noise_buf.push_back(buf);
      slots.push_back({buf, (long long)ns.second, si++, 1, slot_drop_p(cfg, ns.first)});
    }
    e->d_slots_all = e->upload(slots);
    e->d_slots_dyn = e->upload(slots);
    e->stats = e->ws(std::max<size_t>(16, e->plan.stat_names.size()));
    OSRL_CUDA(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    prepare_kernels();
    build_program(*e);
    OSRL_CUDA(cudaDeviceSynchronize());
  } catch (...) {
    free_all(e);
    delete e;
    throw;
  }
  return e;
}

static cudaEvent_t par_event(Engine& e) {
  if (e.par_events_used == e.par_events.size()) {
    cudaEvent_t ev = nullptr;
    OSRL_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e.par_events.push_back(ev);
  }
  return e.par_events[e.par_events_used++];
}
static cudaStream_t par_stream(Engine& e, cudaStream_t of, size_t idx) {
  for (auto& ps : e.par_streams)
    if (ps.first == of) {
      while (ps.second.size() <= idx) {
        cudaStream_t x = nullptr;
        OSRL_CUDA(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        ps.second.push_back(x);
      }
      return ps.second[idx];
    }
  e.par_streams.push_back({of, {}});
  return par_stream(e, of, idx);
}
// Run a program on stream s.  Consecutive ops of one par group (Program::begin_par) are independent: the first stays
// on s, the others go to helper streams forked from / joined to s with events -- inside a capture that makes them
// parallel branches of the graph.  OSRL_PAR=0 keeps everything serial.
static void run_ops(Engine& e, const Program& p, cudaStream_t s) {
  static const bool par_on = [] { const char* v = getenv("OSRL_PAR"); return !(v && v[0] == '0'); }();
  for (size_t i = 0; i < p.ops.size();) {
    const int g = p.meta[i].par;
    size_t j = i + 1;
    if (g && par_on)
      while (j < p.ops.size() && p.meta[j].par == g) ++j;
    if (j - i == 1) { p.ops[i](s); i = j; continue; }
    cudaEvent_t fork = par_event(e);
    OSRL_CUDA(cudaEventRecord(fork, s));
    std::vector<cudaEvent_t> joins;
    for (size_t k = i + 1; k < j; ++k) {
      cudaStream_t sk = par_stream(e, s, k - i - 1);
      OSRL_CUDA(cudaStreamWaitEvent(sk, fork, 0));
      p.ops[k](sk);
      cudaEvent_t jn = par_event(e);
      OSRL_CUDA(cudaEventRecord(jn, sk));
      joins.push_back(jn);
    }
    p.ops[i](s);
    for (cudaEvent_t jn : joins) OSRL_CUDA(cudaStreamWaitEvent(s, jn, 0));
    i = j;
  }
}
static void prologue(Engine& e, cudaStream_t s, unsigned mask = 0xffffffffu) {
  unsigned slots = 0;
  if (e.peer_on)
    for (size_t i = 0; i < e.dp_slot_group.size(); ++i)
      if ((mask >> e.dp_slot_group[i]) & 1u) slots |= 1u << i;
  k_prologue<<<1, 32, 0, s>>>(e.ds, e.d_groups, (int)e.plan.groups.size(), mask, e.peer_on ? e.d_peers : nullptr, slots);
  e.launches++;
}
static void epilogue(Engine& e, cudaStream_t s, int mode = 0, bool hostq = false) {
  if (hostq) k_epilogue<<<1, 32, 0, s>>>(e.ds, mode, e.xq, e.stats, (int)e.plan.stat_names.size());
  else k_epilogue<<<1, 32, 0, s>>>(e.ds, mode);
  e.launches++;
}
// front of a step fed from the host-batch queue (osrl_steps_host): which = 1 reads the VAE branch's batch into the
// "next" buffers, 0 the main branch's (or the whole step's) into the current ones
static void unpack_front(Engine& e, cudaStream_t s, int which, const NoiseSlot* slots, const unsigned long long* counter) {
  const osrl_config& c = e.plan.cfg;
  const bool bc = c.algo == OSRL_ALGO_BC;
  if (which)
    k_unpack_batch<<<24, 256, 0, s>>>(e.xq, 1, e.B, c.obs_dim, c.act_dim, e.nb_obs, e.nb_nobs, e.nb_act, e.nb_rew, e.nb_cost,
                                      e.nb_done, nullptr);
  else
    k_unpack_batch<<<24, 256, 0, s>>>(e.xq, 0, e.B, c.obs_dim, c.act_dim, e.b_obs, bc ? nullptr : e.b_nobs, e.b_act,
                                      bc ? nullptr : e.b_rew, bc ? nullptr : e.b_cost, bc ? nullptr : e.b_done,
                                      c.algo == OSRL_ALGO_COPTIDICE ? e.b_init : nullptr);
  e.launches++;
  if (!e.noise_buf.empty()) {
    k_noise_fill<<<dim3(64, (unsigned)e.noise_buf.size()), 256, 0, s>>>(slots, (int)e.noise_buf.size(), c.seed, counter,
                                                                         (uint32_t)e.rank);
    e.launches++;
  }
}
static void launch_seq_gather(Engine& e, cudaStream_t s, const int* traj_in, const int* start_in, int rows,
                              float* states, float* actions, float* returns, float* ctg, long long* ts, float* mask,
                              float* costs, int* traj_out, int* start_out) {
  const osrl_config& c = e.plan.cfg;
  const int T = c.seq_len;
  k_seq_gather<<<(rows * T + 7) / 8, 256, 0, s>>>(e.sq_rows, e.sq_off, e.sq_ntraj, e.sq_stride, c.obs_dim, c.act_dim, T,
                                                  e.sq_prob, e.sq_alias, traj_in, start_in, c.seed, e.ds, (uint32_t)e.rank,
                                                  rows, states, actions, returns, ctg, ts, mask, costs, traj_out,
                                                  start_out);
  e.launches++;
}
static void sample_front(Engine& e, cudaStream_t s) {
  const osrl_config& c = e.plan.cfg;
  if (c.algo == OSRL_ALGO_CDT) {
    launch_seq_gather(e, s, nullptr, nullptr, e.B, e.s_states, e.s_actions, e.s_returns, e.s_ctg, e.s_ts, e.s_mask,
                      e.s_costs, e.s_traj, e.s_start);
    if (!e.noise_buf.empty()) {   // dropout multipliers
      k_noise_fill<<<dim3(64, (unsigned)e.noise_buf.size()), 256, 0, s>>>(e.d_slots_all, (int)e.noise_buf.size(), c.seed,
                                                                           &e.ds->step, (uint32_t)e.rank);
      e.launches++;
    }
    return;
  }
  const int warps_per_block = 8;
  k_sample_gather<<<(e.B + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, s>>>(
      e.ds_rows, e.ds_n, e.ds_stride, c.obs_dim, c.act_dim, nullptr, c.seed, &e.ds->step, (uint32_t)e.rank, e.B, e.b_obs,
      e.b_nobs, e.b_act, e.b_rew, e.b_cost, e.b_done, e.b_idx, e.b_init);
  e.launches++;
  if (!e.noise_buf.empty()) {
    k_noise_fill<<<dim3(64, (unsigned)e.noise_buf.size()), 256, 0, s>>>(e.d_slots_all, (int)e.noise_buf.size(), c.seed,
                                                                         &e.ds->step, (uint32_t)e.rank);
    e.launches++;
  }
}

static cudaGraphExec_t capture(Engine& e, bool sampled, bool hostq = false) {
  cudaStream_t s = e.cap_stream;
  const int64_t before = e.launches;
  OSRL_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  cudaGraph_t g = nullptr;
  try {
    if (hostq) unpack_front(e, s, 0, e.d_slots_all, &e.ds->step);
    else if (sampled) sample_front(e, s);
    prologue(e, s);
    run_ops(e, e.body, s);
    epilogue(e, s, 0, hostq);
  } catch (...) {
    cudaStreamEndCapture(s, &g);
    if (g) cudaGraphDestroy(g);
    e.launches = before;
    throw;
  }
  OSRL_CUDA(cudaStreamEndCapture(s, &g));
  e.launches = before;  // capture does not execute
  cudaGraphExec_t x = nullptr;
  cudaError_t ce = cudaGraphInstantiate(&x, g, 0);
  cudaGraphDestroy(g);
  if (ce != cudaSuccess) throw Err(OSRL_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
  return x;
}
// ---- pipelined graphs.  side = VAE update of the step whose index is ds->vae_step, on the NEXT minibatch;
// main = the rest of step ds->step on the current minibatch, VAE weights from the snapshot taken before the fork.
static void side_ops(Engine& e, cudaStream_t s, bool hostq = false) {
  const osrl_config& c = e.plan.cfg;
  if (hostq) {
    unpack_front(e, s, 1, e.d_slots_vae, &e.ds->vae_step);
  } else {
    k_sample_gather<<<(e.B + 7) / 8, 256, 0, s>>>(e.ds_rows, e.ds_n, e.ds_stride, c.obs_dim, c.act_dim, nullptr, c.seed,
                                                 &e.ds->vae_step, (uint32_t)e.rank, e.B, e.nb_obs, e.nb_nobs, e.nb_act,
                                                 e.nb_rew, e.nb_cost, e.nb_done, e.nb_idx);
    k_noise_fill<<<dim3(64, (unsigned)e.noise_buf.size()), 256, 0, s>>>(e.d_slots_vae, (int)e.noise_buf.size(), c.seed,
                                                                         &e.ds->vae_step, (uint32_t)e.rank);
    e.launches += 2;
  }
  prologue(e, s, 1u << e.plan.g_vae);
  run_ops(e, e.pa, s);
  epilogue(e, s, 2, hostq);
}
static void main_ops(Engine& e, cudaStream_t s, bool hostq = false) {
  const osrl_config& c = e.plan.cfg;
  if (hostq) {
    unpack_front(e, s, 0, e.d_slots_rest, &e.ds->step);
  } else {
    k_sample_gather<<<(e.B + 7) / 8, 256, 0, s>>>(e.ds_rows, e.ds_n, e.ds_stride, c.obs_dim, c.act_dim, nullptr, c.seed,
                                                 &e.ds->step, (uint32_t)e.rank, e.B, e.b_obs, e.b_nobs, e.b_act, e.b_rew,
                                                 e.b_cost, e.b_done, e.b_idx);
    k_noise_fill<<<dim3(64, (unsigned)e.noise_buf.size()), 256, 0, s>>>(e.d_slots_rest, (int)e.noise_buf.size(), c.seed,
                                                                         &e.ds->step, (uint32_t)e.rank);
    e.launches += 2;
  }
  prologue(e, s, ~(1u << e.plan.g_vae));
  run_ops(e, e.pm, s);
  epilogue(e, s, 1, hostq);
}
static void snapshot_vae(Engine& e, cudaStream_t s) {
  const Group& g = e.plan.groups[e.plan.g_vae];
  OSRL_CUDA(cudaMemcpyAsync(e.Psnap + g.begin, e.P + g.begin, (size_t)(g.end - g.begin) * sizeof(float),
                            cudaMemcpyDeviceToDevice, s));
  for (auto& hd : e.handoff)   // what the VAE branch produced for the step the main branch is about to run
    OSRL_CUDA(cudaMemcpyAsync(hd.cur, hd.next, hd.bytes, cudaMemcpyDeviceToDevice, s));
}
// which: 0 = first VAE update alone, 1 = steady state (fork / join), 2 = last step's remainder alone
static cudaGraphExec_t capture_pipelined(Engine& e, int which, bool hostq = false) {
  cudaStream_t s = e.cap_stream, s2 = e.side_stream;
  const int64_t before = e.launches;
  OSRL_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  cudaGraph_t g = nullptr;
  try {
    if (which == 0) {
      side_ops(e, s, hostq);
    } else {
      snapshot_vae(e, s);
      if (which == 1) {
        OSRL_CUDA(cudaEventRecord(e.ev_fork, s));
        OSRL_CUDA(cudaStreamWaitEvent(s2, e.ev_fork, 0));
        side_ops(e, s2, hostq);
        OSRL_CUDA(cudaEventRecord(e.ev_join, s2));
      }
      main_ops(e, s, hostq);
      if (which == 1) OSRL_CUDA(cudaStreamWaitEvent(s, e.ev_join, 0));
    }
  } catch (...) {
    cudaStreamEndCapture(s, &g);
    if (g) cudaGraphDestroy(g);
    e.launches = before;
    throw;
  }
  OSRL_CUDA(cudaStreamEndCapture(s, &g));
  e.launches = before;
  cudaGraphExec_t x = nullptr;
  cudaError_t ce = cudaGraphInstantiate(&x, g, 0);
  cudaGraphDestroy(g);
  if (ce != cudaSuccess) throw Err(OSRL_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
  return x;
}
static int kernels_per_step(const Engine& e, bool sampled) {
  return e.body.kernels + 2 + (sampled ? (1 + (e.noise_buf.empty() ? 0 : 1)) : 0);
}

// Vose alias table for Categorical(p): slot s keeps itself with probability prob[s], else alias[s]
static void build_alias(const std::vector<double>& p, std::vector<float>& prob, std::vector<int>& alias) {
  const int n = (int)p.size();
  prob.assign(n, 1.f);
  alias.resize(n);
  std::vector<double> q(n);
  std::vector<int> small, large;
  for (int i = 0; i < n; ++i) {
    q[i] = p[i] * n;
    alias[i] = i;
    (q[i] < 1.0 ? small : large).push_back(i);
  }
  while (!small.empty() && !large.empty()) {
    const int s_ = small.back(), l = large.back();
    small.pop_back();
    prob[s_] = (float)q[s_];
    alias[s_] = l;
    q[l] = (q[l] + q[s_]) - 1.0;
    if (q[l] < 1.0) { large.pop_back(); small.push_back(l); }
  }
  for (int i : small) prob[i] = 1.f;
  for (int i : large) prob[i] = 1.f;
}

}  // namespace osrl

// ====================================================================== C ABI
using namespace osrl;

// noise of an explicit-batch step: provided slots are copied in, the others generated on the device
static void stage_noise_impl(Engine& e, const osrl_noise* nz, cudaStream_t s) {
  const int ns = (int)e.noise_buf.size();
  if (!ns) return;
  int provided = 0;
  for (int i = 0; i < ns; ++i) {
    const float* src = nz ? nz->slot[i] : nullptr;
    if (!src) continue;
    ++provided;
    OSRL_CUDA(cudaMemcpyAsync(e.noise_buf[i], src, (size_t)e.plan.noise[i].second * sizeof(float),
                              nz->on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
  }
  if (provided == ns) return;
  const NoiseSlot* slots = e.d_slots_all;  // every slot generated on the device
  if (provided > 0) {                      // mixed: only the missing slots
    std::vector<NoiseSlot> dyn;
    for (int i = 0; i < ns; ++i)
      dyn.push_back({e.noise_buf[i], (long long)e.plan.noise[i].second, i, (nz && nz->slot[i]) ? 0 : 1,
                     slot_drop_p(e.plan.cfg, e.plan.noise[i].first)});
    OSRL_CUDA(cudaMemcpyAsync(e.d_slots_dyn, dyn.data(), dyn.size() * sizeof(NoiseSlot), cudaMemcpyHostToDevice, s));
    OSRL_CUDA(cudaStreamSynchronize(s));
    slots = e.d_slots_dyn;
  }
  k_noise_fill<<<dim3(64, (unsigned)ns), 256, 0, s>>>(slots, ns, e.plan.cfg.seed, &e.ds->step, (uint32_t)e.rank);
  e.launches++;
}

#define OSRL_TRY try {
#define OSRL_CATCH                                   \
  }                                                  \
  catch (const Err& x) {                             \
    g_err = x.what();                                \
    return x.code;                                   \
  }                                                  \
  catch (const std::exception& x) {                  \
    g_err = x.what();                                \
    return OSRL_ERR_STATE;                           \
  }                                                  \
  return OSRL_OK;

struct osrl_engine {
  Engine* e;
};

static void fill_desc(const ParamEntry& pe, osrl_param_desc* d, const Engine* e) {
  memset(d, 0, sizeof(*d));
  strncpy(d->name, pe.name.c_str(), sizeof(d->name) - 1);
  d->rows = pe.rows; d->cols = pe.cols; d->offset = pe.offset; d->section = pe.section; d->group = pe.group;
  d->ptr = e ? ((pe.section == 1 ? e->T : e->P) + pe.offset) : nullptr;
}

extern "C" {

int osrl_abi_version(void) { return OSRL_ABI_VERSION; }
const char* osrl_last_error(void) { return g_err.c_str(); }

int osrl_plan(const osrl_config* cfg, osrl_param_desc* out, int cap, int* n) {
  OSRL_TRY
  OSRL_REQUIRE(cfg && n, "null argument");
  osrl_config c = *cfg;
  if (c.batch_size <= 0) c.batch_size = 1;
  if (c.world_size <= 0) c.world_size = 1;
  Plan p = make_plan(c);
  *n = (int)p.table.size();
  if (out)
    for (int i = 0; i < *n && i < cap; ++i) fill_desc(p.table[i], &out[i], nullptr);
  OSRL_CATCH
}

int osrl_engine_create(const osrl_config* cfg, int device, osrl_engine** out) {
  OSRL_TRY
  OSRL_REQUIRE(cfg && out, "null argument");
  osrl_config c = *cfg;
  if (c.world_size <= 0) { c.world_size = 1; c.rank = 0; }
  Engine* e = create(c, device);
  *out = new osrl_engine{e};
  OSRL_CATCH
}

void osrl_engine_destroy(osrl_engine* h) {
  if (!h) return;
  cudaSetDevice(h->e->device);
  cudaDeviceSynchronize();
  free_all(h->e);
  delete h->e;
  delete h;
}

int osrl_param_table(osrl_engine* h, osrl_param_desc* out, int cap, int* n) {
  OSRL_TRY
  OSRL_REQUIRE(h && n, "null argument");
  *n = (int)h->e->plan.table.size();
  if (out)
    for (int i = 0; i < *n && i < cap; ++i) fill_desc(h->e->plan.table[i], &out[i], h->e);
  OSRL_CATCH
}

static int64_t entry_count(const ParamEntry& pe) { return pe.rows * (pe.cols ? pe.cols : 1); }

int osrl_param_set(osrl_engine* h, int index, const float* host, int64_t count) {
  OSRL_TRY
  OSRL_REQUIRE(h && host, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(index >= 0 && index < (int)e.plan.table.size(), "bad parameter index");
  const ParamEntry& pe = e.plan.table[index];
  OSRL_REQUIRE(count == entry_count(pe), "parameter element count mismatch");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaMemcpy((pe.section == 1 ? e.T : e.P) + pe.offset, host, count * sizeof(float), cudaMemcpyHostToDevice));
  OSRL_CATCH
}
int osrl_param_get(osrl_engine* h, int index, float* host, int64_t count) {
  OSRL_TRY
  OSRL_REQUIRE(h && host, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(index >= 0 && index < (int)e.plan.table.size(), "bad parameter index");
  const ParamEntry& pe = e.plan.table[index];
  OSRL_REQUIRE(count == entry_count(pe), "parameter element count mismatch");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  OSRL_CUDA(cudaMemcpy(host, (pe.section == 1 ? e.T : e.P) + pe.offset, count * sizeof(float), cudaMemcpyDeviceToHost));
  OSRL_CATCH
}
int osrl_sync_targets(osrl_engine* h) {
  OSRL_TRY
  OSRL_REQUIRE(h, "null argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  for (auto& g : e.plan.groups)
    if (g.has_target)
      OSRL_CUDA(cudaMemcpy(e.T + g.begin, e.P + g.begin, (g.end - g.begin) * sizeof(float), cudaMemcpyDeviceToDevice));
  OSRL_CATCH
}

int osrl_buffer_upload(osrl_engine* h, const osrl_dataset_view* v) {
  OSRL_TRY
  OSRL_REQUIRE(h && v, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(v->n > 0 && v->observations && v->next_observations && v->actions && v->rewards && v->costs,
               "dataset view incomplete");
  OSRL_REQUIRE(v->done || (v->terminals && v->timeouts), "need done or terminals+timeouts");
  OSRL_CUDA(cudaSetDevice(e.device));
  const int o = e.plan.cfg.obs_dim, a = e.plan.cfg.act_dim;
  const bool cop = e.plan.cfg.algo == OSRL_ALGO_COPTIDICE;
  if (cop) OSRL_REQUIRE(v->is_init, "COptiDICE datasets carry is_init (TransitionDataset(state_init=True))");
  const int stride = (2 * o + a + 3 + (cop ? 1 : 0) + 3) / 4 * 4;
  std::vector<float> packed((size_t)v->n * stride, 0.f);
  const float rs = v->reward_scale, cs = v->cost_scale;
  for (int64_t i = 0; i < v->n; ++i) {
    float* r = packed.data() + (size_t)i * stride;
    memcpy(r, v->observations + (size_t)i * o, o * sizeof(float));
    memcpy(r + o, v->next_observations + (size_t)i * o, o * sizeof(float));
    memcpy(r + 2 * o, v->actions + (size_t)i * a, a * sizeof(float));
    r[2 * o + a] = v->rewards[i] * rs;   // dataset.py:836 (float32 * weak python float -> float32)
    r[2 * o + a + 1] = v->costs[i] * cs; // dataset.py:837
    r[2 * o + a + 2] = v->done ? v->done[i] : ((v->terminals[i] || v->timeouts[i]) ? 1.f : 0.f);  // :815-816
    if (cop) r[2 * o + a + 3] = v->is_init[i];
  }
  if (e.ds_rows) {
    OSRL_CUDA(cudaDeviceSynchronize());
    drop_sampled_graphs(e);
    release_alloc(e, e.ds_rows);
    e.ds_rows = nullptr;
  }
  void* d = nullptr;
  OSRL_CUDA(cudaMalloc(&d, packed.size() * sizeof(float)));
  e.allocs.push_back(d);
  OSRL_CUDA(cudaMemcpy(d, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  e.ds_rows = (float*)d;
  e.ds_n = v->n;
  e.ds_stride = stride;
  drop_sampled_graphs(e);
  OSRL_CATCH
}

int osrl_seq_buffer_upload(osrl_engine* h, const osrl_seq_dataset_view* v) {
  OSRL_TRY
  OSRL_REQUIRE(h && v, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.plan.cfg.algo == OSRL_ALGO_CDT, "trajectory buffers belong to CDT engines");
  OSRL_REQUIRE(v->n > 0 && v->n_traj > 0 && v->observations && v->actions && v->returns && v->cost_returns &&
                   v->costs && v->traj_offsets,
               "trajectory view incomplete");
  OSRL_REQUIRE(v->traj_offsets[0] == 0 && v->traj_offsets[v->n_traj] == v->n, "traj_offsets must span [0, n]");
  OSRL_CUDA(cudaSetDevice(e.device));
  const int o = e.plan.cfg.obs_dim, a = e.plan.cfg.act_dim;
  const int stride = (o + a + 3 + 3) / 4 * 4;
  std::vector<float> packed((size_t)v->n * stride, 0.f);
  for (int64_t i = 0; i < v->n; ++i) {
    float* r = packed.data() + (size_t)i * stride;
    memcpy(r, v->observations + (size_t)i * o, o * sizeof(float));
    memcpy(r + o, v->actions + (size_t)i * a, a * sizeof(float));
    r[o + a] = v->returns[i] * v->reward_scale;          // dataset.py:762 (float32 * weak python float)
    r[o + a + 1] = v->cost_returns[i] * v->cost_scale;   // :763
    r[o + a + 2] = v->costs[i];
  }
  std::vector<double> p(v->n_traj, 1.0 / v->n_traj);
  if (v->sample_prob) {
    double tot = 0;
    for (int i = 0; i < v->n_traj; ++i) { OSRL_REQUIRE(v->sample_prob[i] >= 0, "negative sample_prob"); tot += v->sample_prob[i]; }
    OSRL_REQUIRE(tot > 0, "sample_prob sums to zero");
    for (int i = 0; i < v->n_traj; ++i) p[i] = v->sample_prob[i] / tot;
  }
  for (int i = 0; i < v->n_traj; ++i) OSRL_REQUIRE(v->traj_offsets[i + 1] > v->traj_offsets[i], "empty trajectory");
  std::vector<float> prob;
  std::vector<int> alias;
  build_alias(p, prob, alias);
  std::vector<long long> off(v->traj_offsets, v->traj_offsets + v->n_traj + 1);
  if (e.sq_rows) {
    OSRL_CUDA(cudaDeviceSynchronize());
    drop_sampled_graphs(e);
    for (void* q : {(void*)e.sq_rows, (void*)e.sq_off, (void*)e.sq_prob, (void*)e.sq_alias, (void*)e.sq_first_ret,
                    (void*)e.sq_first_cret})
      release_alloc(e, q);
    e.sq_rows = nullptr; e.sq_first_ret = e.sq_first_cret = nullptr;
  }
  e.sq_rows = e.upload(packed);
  e.sq_off = e.upload(off);
  e.sq_prob = e.upload(prob);
  e.sq_alias = e.upload(alias);
  e.sq_ntraj = (int)v->n_traj;
  e.sq_stride = stride;
  e.ds_rows = e.sq_rows;  // marks "a resident dataset exists" for osrl_steps
  drop_sampled_graphs(e);
  OSRL_CATCH
}

// ---- trajectory preprocessing on the device (preproc_kernels.cuh)
static void seq_install_prob(Engine& e, const double* prob, int n) {
  std::vector<double> p(n, 1.0 / n);
  if (prob) {
    double tot = 0;
    for (int i = 0; i < n; ++i) { OSRL_REQUIRE(prob[i] >= 0, "negative sample_prob"); tot += prob[i]; }
    OSRL_REQUIRE(tot > 0, "sample_prob sums to zero");
    for (int i = 0; i < n; ++i) p[i] = prob[i] / tot;
  }
  std::vector<float> pr;
  std::vector<int> alias;
  build_alias(p, pr, alias);
  if (e.sq_prob) { release_alloc(e, e.sq_prob); release_alloc(e, e.sq_alias); }
  e.sq_prob = e.upload(pr);
  e.sq_alias = e.upload(alias);
}

int osrl_seq_preprocess(osrl_engine* h, const osrl_dataset_view* v, int cost_reverse, int64_t* n_traj_out,
                        int64_t* n_used_out) {
  OSRL_TRY
  OSRL_REQUIRE(h && v && n_traj_out, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.plan.cfg.algo == OSRL_ALGO_CDT, "trajectory buffers belong to CDT engines");
  OSRL_REQUIRE(v->n > 0 && v->observations && v->actions && v->rewards && v->costs && (v->terminals || v->timeouts),
               "flat dataset view incomplete (observations, actions, rewards, costs, terminals / timeouts)");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  const int o = e.plan.cfg.obs_dim, a = e.plan.cfg.act_dim;
  const int stride = (o + a + 3 + 3) / 4 * 4;
  const long long n = v->n;
  // staging copies of the flat arrays (freed before returning)
  std::vector<void*> tmp;
  auto stage = [&](const void* src, size_t bytes) -> void* {
    if (!src) return nullptr;
    void* d = nullptr;
    OSRL_CUDA(cudaMalloc(&d, bytes));
    tmp.push_back(d);
    OSRL_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    return d;
  };
  struct Cleanup { std::vector<void*>& t; ~Cleanup() { for (void* q : t) cudaFree(q); } } cleanup{tmp};
  const float* d_obs = (const float*)stage(v->observations, (size_t)n * o * sizeof(float));
  const float* d_act = (const float*)stage(v->actions, (size_t)n * a * sizeof(float));
  const float* d_rew = (const float*)stage(v->rewards, (size_t)n * sizeof(float));
  const float* d_cost = (const float*)stage(v->costs, (size_t)n * sizeof(float));
  const uint8_t* d_term = (const uint8_t*)stage(v->terminals, (size_t)n);
  const uint8_t* d_tout = (const uint8_t*)stage(v->timeouts, (size_t)n);
  long long* d_off = nullptr;
  OSRL_CUDA(cudaMalloc((void**)&d_off, (size_t)(n + 1) * sizeof(long long)));
  tmp.push_back(d_off);
  EpisodeCount* d_cnt = nullptr;
  OSRL_CUDA(cudaMalloc((void**)&d_cnt, sizeof(EpisodeCount)));
  tmp.push_back(d_cnt);
  k_episode_offsets<<<1, 1024>>>(d_term, d_tout, n, d_off, d_cnt);
  k_episode_finish<<<1, 1>>>(d_off, d_cnt);
  e.launches += 2;
  EpisodeCount cnt;
  OSRL_CUDA(cudaMemcpy(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
  OSRL_REQUIRE(cnt.n_traj > 0, "no finished episode in the dataset (terminals | timeouts never set)");
  OSRL_REQUIRE(cnt.n_traj < (1ll << 31), "too many episodes");
  // the resident buffer
  if (e.sq_rows) {
    drop_sampled_graphs(e);
    for (void* q : {(void*)e.sq_rows, (void*)e.sq_off, (void*)e.sq_prob, (void*)e.sq_alias, (void*)e.sq_first_ret,
                    (void*)e.sq_first_cret})
      release_alloc(e, q);
    e.sq_rows = nullptr; e.sq_prob = nullptr; e.sq_alias = nullptr; e.sq_first_ret = e.sq_first_cret = nullptr;
  }
  void* rows = nullptr;
  OSRL_CUDA(cudaMalloc(&rows, (size_t)cnt.n_used * stride * sizeof(float)));
  e.allocs.push_back(rows);
  void* off = nullptr;
  OSRL_CUDA(cudaMalloc(&off, (size_t)(cnt.n_traj + 1) * sizeof(long long)));
  e.allocs.push_back(off);
  OSRL_CUDA(cudaMemcpy(off, d_off, (size_t)(cnt.n_traj + 1) * sizeof(long long), cudaMemcpyDeviceToDevice));
  e.sq_first_ret = e.ws((size_t)cnt.n_traj);
  e.sq_first_cret = e.ws((size_t)cnt.n_traj);
  k_seq_pack<<<148 * 8, 256>>>(d_obs, d_act, d_cost, cnt.n_used, o, a, stride, cost_reverse ? 1 : 0, (float*)rows);
  k_episode_suffix<<<(unsigned)((cnt.n_traj + 127) / 128), 128>>>(d_rew, (const long long*)off, cnt.n_traj, o, a, stride,
                                                                  v->reward_scale, v->cost_scale, (float*)rows,
                                                                  e.sq_first_ret, e.sq_first_cret);
  e.launches += 2;
  OSRL_CUDA(cudaGetLastError());
  OSRL_CUDA(cudaDeviceSynchronize());
  e.sq_rows = (float*)rows;
  e.sq_off = (long long*)off;
  e.sq_ntraj = (int)cnt.n_traj;
  e.sq_stride = stride;
  seq_install_prob(e, nullptr, e.sq_ntraj);   // uniform until osrl_seq_set_sample_prob
  e.ds_rows = e.sq_rows;
  drop_sampled_graphs(e);
  *n_traj_out = cnt.n_traj;
  if (n_used_out) *n_used_out = cnt.n_used;
  OSRL_CATCH
}

int osrl_seq_episode_info(osrl_engine* h, float* first_return, float* first_cost_return, int64_t* offsets, int cap) {
  OSRL_TRY
  OSRL_REQUIRE(h, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.sq_rows && e.sq_first_ret, "no preprocessed trajectory buffer (osrl_seq_preprocess)");
  OSRL_REQUIRE(cap >= e.sq_ntraj, "buffers too small");
  OSRL_CUDA(cudaSetDevice(e.device));
  if (first_return) OSRL_CUDA(cudaMemcpy(first_return, e.sq_first_ret, e.sq_ntraj * sizeof(float), cudaMemcpyDeviceToHost));
  if (first_cost_return)
    OSRL_CUDA(cudaMemcpy(first_cost_return, e.sq_first_cret, e.sq_ntraj * sizeof(float), cudaMemcpyDeviceToHost));
  if (offsets) {
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    OSRL_CUDA(cudaMemcpy(offsets, e.sq_off, (size_t)(e.sq_ntraj + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  OSRL_CATCH
}

int osrl_seq_set_sample_prob(osrl_engine* h, const double* prob, int n) {
  OSRL_TRY
  OSRL_REQUIRE(h, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.sq_rows, "no trajectory buffer");
  OSRL_REQUIRE(!prob || n == e.sq_ntraj, "sample_prob length != number of trajectories");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  drop_sampled_graphs(e);
  seq_install_prob(e, prob, e.sq_ntraj);
  OSRL_CATCH
}

int osrl_seq_alias_table(osrl_engine* h, float* prob_out, int32_t* alias_out, int cap) {
  OSRL_TRY
  OSRL_REQUIRE(h && prob_out && alias_out, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.sq_rows && cap >= e.sq_ntraj, "no trajectory buffer / buffer too small");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaMemcpy(prob_out, e.sq_prob, e.sq_ntraj * sizeof(float), cudaMemcpyDeviceToHost));
  OSRL_CUDA(cudaMemcpy(alias_out, e.sq_alias, e.sq_ntraj * sizeof(int), cudaMemcpyDeviceToHost));
  OSRL_CATCH
}

int osrl_seq_gather(osrl_engine* h, const int32_t* traj_idx, const int32_t* start_idx, int n, osrl_seq_batch* out,
                    void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && traj_idx && start_idx && out, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.sq_rows, "no trajectory buffer uploaded");
  OSRL_REQUIRE(!out->on_host, "osrl_seq_gather writes device buffers");
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  void *dt = nullptr, *dsx = nullptr;
  OSRL_CUDA(cudaMalloc(&dt, (size_t)n * sizeof(int)));
  OSRL_CUDA(cudaMalloc(&dsx, (size_t)n * sizeof(int)));
  OSRL_CUDA(cudaMemcpyAsync(dt, traj_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
  OSRL_CUDA(cudaMemcpyAsync(dsx, start_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
  launch_seq_gather(e, s, (const int*)dt, (const int*)dsx, n, (float*)out->states, (float*)out->actions,
                    (float*)out->returns, (float*)out->costs_return, (long long*)out->time_steps, (float*)out->mask,
                    (float*)out->costs, nullptr, nullptr);
  OSRL_CUDA(cudaStreamSynchronize(s));
  cudaFree(dt);
  cudaFree(dsx);
  OSRL_CATCH
}

int osrl_last_sequences(osrl_engine* h, int32_t* traj_out, int32_t* start_out, int cap) {
  OSRL_TRY
  OSRL_REQUIRE(h && traj_out && start_out && cap >= h->e->B, "bad argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.s_traj, "not a CDT engine");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  OSRL_CUDA(cudaMemcpy(traj_out, e.s_traj, e.B * sizeof(int), cudaMemcpyDeviceToHost));
  OSRL_CUDA(cudaMemcpy(start_out, e.s_start, e.B * sizeof(int), cudaMemcpyDeviceToHost));
  OSRL_CATCH
}

int osrl_gather(osrl_engine* h, const int64_t* idx, int n, int idx_on_host, osrl_batch* out, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && idx && out, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.ds_rows, "no dataset uploaded");
  OSRL_REQUIRE(n >= 0, "negative count");
  if (n == 0) return OSRL_OK;
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int o = e.plan.cfg.obs_dim, a = e.plan.cfg.act_dim;
  const int64_t* didx = idx;
  void* tmp_idx = nullptr;
  if (idx_on_host) {
    for (int i = 0; i < n; ++i) OSRL_REQUIRE(idx[i] >= 0 && idx[i] < e.ds_n, "index out of range");
    OSRL_CUDA(cudaMalloc(&tmp_idx, (size_t)n * sizeof(int64_t)));
    OSRL_CUDA(cudaMemcpyAsync(tmp_idx, idx, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    didx = (const int64_t*)tmp_idx;
  }
  float* dst[6] = {(float*)out->observations, (float*)out->next_observations, (float*)out->actions,
                   (float*)out->rewards,      (float*)out->costs,             (float*)out->done};
  const size_t cnt[6] = {(size_t)n * o, (size_t)n * o, (size_t)n * a, (size_t)n, (size_t)n, (size_t)n};
  float* dev[6];
  void* tmp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int k = 0; k < 6; ++k) {
    if (!dst[k]) { dev[k] = nullptr; continue; }
    if (out->on_host) { OSRL_CUDA(cudaMalloc(&tmp[k], cnt[k] * sizeof(float))); dev[k] = (float*)tmp[k]; }
    else dev[k] = dst[k];
  }
  k_sample_gather<<<(n + 7) / 8, 256, 0, s>>>(e.ds_rows, e.ds_n, e.ds_stride, o, a, didx, 0, &e.ds->step, 0, n, dev[0], dev[1],
                                              dev[2], dev[3], dev[4], dev[5], nullptr);
  e.launches++;
  OSRL_CUDA(cudaGetLastError());
  if (out->on_host)
    for (int k = 0; k < 6; ++k)
      if (dst[k]) OSRL_CUDA(cudaMemcpyAsync(dst[k], dev[k], cnt[k] * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out->on_host || tmp_idx) OSRL_CUDA(cudaStreamSynchronize(s));
  for (int k = 0; k < 6; ++k)
    if (tmp[k]) cudaFree(tmp[k]);
  if (tmp_idx) cudaFree(tmp_idx);
  OSRL_CATCH
}

int osrl_step(osrl_engine* h, const osrl_batch* b, const osrl_noise* nz, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && b, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.plan.cfg.algo != OSRL_ALGO_CDT, "CDT steps take a sequence batch: use osrl_step_seq");
  OSRL_REQUIRE(b->rows == e.B, "batch rows != engine batch_size");
  OSRL_REQUIRE(b->observations && b->actions, "observations/actions required");
  const bool bc = e.plan.cfg.algo == OSRL_ALGO_BC;
  if (!bc) OSRL_REQUIRE(b->next_observations && b->rewards && b->costs && b->done, "incomplete transition batch");
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int o = e.plan.cfg.obs_dim, a = e.plan.cfg.act_dim, B = e.B;
  const cudaMemcpyKind kind = b->on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  OSRL_CUDA(cudaMemcpyAsync(e.b_obs, b->observations, (size_t)B * o * sizeof(float), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.b_act, b->actions, (size_t)B * a * sizeof(float), kind, s));
  if (!bc) {
    OSRL_CUDA(cudaMemcpyAsync(e.b_nobs, b->next_observations, (size_t)B * o * sizeof(float), kind, s));
    OSRL_CUDA(cudaMemcpyAsync(e.b_rew, b->rewards, (size_t)B * sizeof(float), kind, s));
    OSRL_CUDA(cudaMemcpyAsync(e.b_cost, b->costs, (size_t)B * sizeof(float), kind, s));
    OSRL_CUDA(cudaMemcpyAsync(e.b_done, b->done, (size_t)B * sizeof(float), kind, s));
  }
  if (e.plan.cfg.algo == OSRL_ALGO_COPTIDICE) {
    OSRL_REQUIRE(b->is_init, "COptiDICE batches carry is_init (coptidice.py:126-127)");
    OSRL_CUDA(cudaMemcpyAsync(e.b_init, b->is_init, (size_t)B * sizeof(float), kind, s));
  }
  stage_noise_impl(e, nz, s);
  if (!e.g_body) e.g_body = capture(e, false);
  OSRL_CUDA(cudaGraphLaunch(e.g_body, s));
  e.launches += kernels_per_step(e, false);
  OSRL_CATCH
}

int osrl_step_seq(osrl_engine* h, const osrl_seq_batch* b, const osrl_noise* nz, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && b, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.plan.cfg.algo == OSRL_ALGO_CDT, "osrl_step_seq is the CDT entry point");
  OSRL_REQUIRE(b->rows == e.B && b->seq_len == e.plan.cfg.seq_len, "batch rows / seq_len differ from the engine config");
  OSRL_REQUIRE(b->states && b->actions && b->returns && b->costs_return && b->time_steps && b->mask && b->costs,
               "incomplete sequence batch");
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t BT = (size_t)e.B * b->seq_len;
  const int o = e.plan.cfg.obs_dim, a = e.plan.cfg.act_dim;
  const cudaMemcpyKind kind = b->on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  OSRL_CUDA(cudaMemcpyAsync(e.s_states, b->states, BT * o * sizeof(float), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.s_actions, b->actions, BT * a * sizeof(float), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.s_returns, b->returns, BT * sizeof(float), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.s_ctg, b->costs_return, BT * sizeof(float), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.s_ts, b->time_steps, BT * sizeof(long long), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.s_mask, b->mask, BT * sizeof(float), kind, s));
  OSRL_CUDA(cudaMemcpyAsync(e.s_costs, b->costs, BT * sizeof(float), kind, s));
  stage_noise_impl(e, nz, s);
  if (!e.g_body) e.g_body = capture(e, false);
  OSRL_CUDA(cudaGraphLaunch(e.g_body, s));
  e.launches += kernels_per_step(e, false);
  OSRL_CATCH
}

int osrl_steps(osrl_engine* h, int k, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && k >= 0, "bad argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(e.ds_rows, "osrl_steps needs a resident dataset (osrl_buffer_upload / osrl_seq_buffer_upload)");
  if (e.plan.cfg.algo == OSRL_ALGO_CDT) OSRL_REQUIRE(e.sq_rows, "CDT needs osrl_seq_buffer_upload");
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (e.pipelined && k >= 2 && (e.world == 1 || e.comm2)) {
    // VAE update of step s+1 overlapped with the rest of step s; same kernels on the same data in the same
    // per-parameter order as the sequential graph, so the state after k steps is bit-identical to k x osrl_steps(1)
    if (!e.g_pro) e.g_pro = capture_pipelined(e, 0);
    if (!e.g_mid) e.g_mid = capture_pipelined(e, 1);
    if (!e.g_last) e.g_last = capture_pipelined(e, 2);
    OSRL_CUDA(cudaGraphLaunch(e.g_pro, s));
    for (int i = 0; i + 1 < k; ++i) OSRL_CUDA(cudaGraphLaunch(e.g_mid, s));
    OSRL_CUDA(cudaGraphLaunch(e.g_last, s));
    e.launches += (int64_t)k * (e.pa.kernels + e.pm.kernels + 8);
  } else {
    if (!e.g_sampled) e.g_sampled = capture(e, true);
    for (int i = 0; i < k; ++i) OSRL_CUDA(cudaGraphLaunch(e.g_sampled, s));
    e.launches += (int64_t)k * kernels_per_step(e, true);
  }
  OSRL_CATCH
}

// k steps on k explicit HOST minibatches (stacked [k][B][...] arrays), the batched form of osrl_step: the batches are
// packed into a pinned ring the step graphs read directly (mapped memory: one PCIe read per step, no copy-engine
// round trip on the stream), the per-step stats come back the same way, and for the VAE algorithms the steps are
// pipelined exactly like osrl_steps().  Bit-identical to k x osrl_step(batch_j, NULL noise).
int osrl_steps_host(osrl_engine* h, const osrl_batch* stk, int k, float* stats_out, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && stk && k >= 0, "bad argument");
  Engine& e = *h->e;
  const osrl_config& c = e.plan.cfg;
  OSRL_REQUIRE(c.algo != OSRL_ALGO_CDT, "CDT steps take sequence batches: use osrl_step_seq");
  OSRL_REQUIRE(stk->rows == e.B, "batch rows != engine batch_size");
  OSRL_REQUIRE(stk->on_host, "osrl_steps_host takes host batches (device batches: osrl_step)");
  OSRL_REQUIRE(stk->observations && stk->actions, "observations/actions required");
  const bool bc = c.algo == OSRL_ALGO_BC, cop = c.algo == OSRL_ALGO_COPTIDICE;
  if (!bc) OSRL_REQUIRE(stk->next_observations && stk->rewards && stk->costs && stk->done, "incomplete transition batch");
  if (cop) OSRL_REQUIRE(stk->is_init, "COptiDICE batches carry is_init (coptidice.py:126-127)");
  if (k == 0) return OSRL_OK;
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int o = c.obs_dim, a = c.act_dim, B = e.B, ns = (int)e.plan.stat_names.size();
  OSRL_REQUIRE(ns <= HQ_STAT_LD, "too many stats");
  const size_t slot = (size_t)B * (2 * o + a + 4);
  if (e.x_pending) { OSRL_CUDA(cudaStreamSynchronize(s)); e.x_pending = false; }
  if (k > e.x_cap) {
    OSRL_CUDA(cudaDeviceSynchronize());
    for (float** hp : {&e.x_ring, &e.x_st_side, &e.x_st_main})
      if (*hp) { cudaFreeHost(*hp); *hp = nullptr; }
    e.x_cap = 0;
    const int cap = std::max(k, 64);
    OSRL_CUDA(cudaHostAlloc((void**)&e.x_ring, cap * slot * sizeof(float), cudaHostAllocMapped));
    OSRL_CUDA(cudaHostAlloc((void**)&e.x_st_side, (size_t)cap * HQ_STAT_LD * sizeof(float), cudaHostAllocMapped));
    OSRL_CUDA(cudaHostAlloc((void**)&e.x_st_main, (size_t)cap * HQ_STAT_LD * sizeof(float), cudaHostAllocMapped));
    e.x_cap = cap;
  }
  if (!e.xq) { OSRL_CUDA(cudaMalloc((void**)&e.xq, sizeof(HostQueue))); e.allocs.push_back(e.xq); }
  HostQueue hq;
  memset(&hq, 0, sizeof(hq));
  OSRL_CUDA(cudaHostGetDevicePointer((void**)&hq.ring, e.x_ring, 0));
  OSRL_CUDA(cudaHostGetDevicePointer((void**)&hq.st_side, e.x_st_side, 0));
  OSRL_CUDA(cudaHostGetDevicePointer((void**)&hq.st_main, e.x_st_main, 0));
  hq.slot_floats = (long long)slot;
  OSRL_CUDA(cudaMemcpyAsync(e.xq, &hq, sizeof(hq), cudaMemcpyHostToDevice, s));   // pageable source: staged before return
  auto pack = [&](int j) {
    float* d = e.x_ring + (size_t)j * slot;
    auto put = [&](const float* src, size_t per) {
      if (src) memcpy(d, src + (size_t)j * per, per * sizeof(float));
      d += per;
    };
    put(stk->observations, (size_t)B * o); put(stk->next_observations, (size_t)B * o); put(stk->actions, (size_t)B * a);
    put(stk->rewards, B); put(stk->costs, B); put(stk->done, B); put(stk->is_init, B);
  };
  const bool pipe = e.pipelined && k >= 2 && (e.world == 1 || e.comm2);
  if (pipe) {
    if (!e.g_xpro) e.g_xpro = capture_pipelined(e, 0, true);
    if (!e.g_xmid) e.g_xmid = capture_pipelined(e, 1, true);
    if (!e.g_xlast) e.g_xlast = capture_pipelined(e, 2, true);
    pack(0);
    OSRL_CUDA(cudaGraphLaunch(e.g_xpro, s));
    for (int j = 1; j < k; ++j) {
      pack(j);   // the GPU is at most at batch j-1: the host packs ahead of it
      OSRL_CUDA(cudaGraphLaunch(e.g_xmid, s));
    }
    OSRL_CUDA(cudaGraphLaunch(e.g_xlast, s));
    e.launches += (int64_t)k * (e.pa.kernels + e.pm.kernels + 8);
  } else {
    if (!e.g_xbody) e.g_xbody = capture(e, false, true);
    for (int j = 0; j < k; ++j) {
      pack(j);
      OSRL_CUDA(cudaGraphLaunch(e.g_xbody, s));
    }
    e.launches += (int64_t)k * kernels_per_step(e, true);
  }
  e.x_last_k = k;
  e.x_last_pipelined = pipe;
  e.x_pending = true;
  if (stats_out) {
    OSRL_CUDA(cudaStreamSynchronize(s));
    e.x_pending = false;
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < ns; ++i)
        stats_out[(size_t)j * ns + i] = (pipe && (e.side_stat_mask >> i & 1u)) ? e.x_st_side[(size_t)j * HQ_STAT_LD + i]
                                                                               : e.x_st_main[(size_t)j * HQ_STAT_LD + i];
  }
  OSRL_CATCH
}

int osrl_stat_names(osrl_engine* h, const char** names, int cap, int* n) {
  OSRL_TRY
  OSRL_REQUIRE(h && n, "null argument");
  *n = (int)h->e->plan.stat_names.size();
  if (names)
    for (int i = 0; i < *n && i < cap; ++i) names[i] = h->e->plan.stat_names[i].c_str();
  OSRL_CATCH
}
int osrl_stats(osrl_engine* h, float* host_out, int cap, int* n, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_out && n, "null argument");
  Engine& e = *h->e;
  *n = (int)e.plan.stat_names.size();
  OSRL_REQUIRE(cap >= *n, "stats buffer too small");
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  OSRL_CUDA(cudaMemcpyAsync(host_out, e.stats, *n * sizeof(float), cudaMemcpyDeviceToHost, s));
  OSRL_CUDA(cudaStreamSynchronize(s));
  if (e.peer_on) {   // a peer that never arrived (crashed rank, diverged control flow): fail loudly instead of training on
    unsigned to = 0;
    OSRL_CUDA(cudaMemcpy(&to, &e.dp_flags->timeout, sizeof(to), cudaMemcpyDeviceToHost));
    if (to) throw Err(OSRL_ERR_NCCL, "data-parallel peer exchange timed out waiting for another rank");
  }
  OSRL_CATCH
}

int osrl_stats_lagged(osrl_engine* h, float* host_out, int cap, int* n, int* valid, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_out && n && valid, "null argument");
  Engine& e = *h->e;
  *n = (int)e.plan.stat_names.size();
  OSRL_REQUIRE(cap >= *n, "stats buffer too small");
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (!e.stats_pinned[0])
    for (int i = 0; i < 2; ++i) {
      OSRL_CUDA(cudaMallocHost((void**)&e.stats_pinned[i], 16 * sizeof(float)));
      OSRL_CUDA(cudaEventCreateWithFlags(&e.stats_ev[i], cudaEventDisableTiming));
    }
  const int cur = e.stats_slot, prev = cur ^ 1;
  OSRL_CUDA(cudaMemcpyAsync(e.stats_pinned[cur], e.stats, *n * sizeof(float), cudaMemcpyDeviceToHost, s));
  OSRL_CUDA(cudaEventRecord(e.stats_ev[cur], s));
  *valid = e.stats_calls > 0;
  if (*valid) {
    OSRL_CUDA(cudaEventSynchronize(e.stats_ev[prev]));   // (recorded one step ago: no wait in steady state)
    memcpy(host_out, e.stats_pinned[prev], *n * sizeof(float));
  }
  e.stats_slot = prev;
  e.stats_calls++;
  OSRL_CATCH
}

// ---- resumable checkpoint blob: header | P | T | M | V | DevState
struct StateHeader { char magic[8]; int32_t abi, algo; int64_t nP; int32_t dev_state_bytes, reserved; };
int osrl_state_size(osrl_engine* h, int64_t* bytes) {
  OSRL_TRY
  OSRL_REQUIRE(h && bytes, "null argument");
  *bytes = (int64_t)sizeof(StateHeader) + 4 * h->e->plan.nP * (int64_t)sizeof(float) + (int64_t)sizeof(DevState);
  OSRL_CATCH
}
int osrl_state_save(osrl_engine* h, void* host_buf, int64_t cap) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_buf, "null argument");
  Engine& e = *h->e;
  const int64_t sec = e.plan.nP * (int64_t)sizeof(float);
  OSRL_REQUIRE(cap >= (int64_t)sizeof(StateHeader) + 4 * sec + (int64_t)sizeof(DevState), "state buffer too small");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  StateHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, "OSRLB200", 8);
  hd.abi = OSRL_ABI_VERSION; hd.algo = e.plan.cfg.algo; hd.nP = e.plan.nP; hd.dev_state_bytes = (int32_t)sizeof(DevState);
  char* o = (char*)host_buf;
  memcpy(o, &hd, sizeof(hd)); o += sizeof(hd);
  for (float* src : {e.P, e.T, e.M, e.V}) { OSRL_CUDA(cudaMemcpy(o, src, sec, cudaMemcpyDeviceToHost)); o += sec; }
  OSRL_CUDA(cudaMemcpy(o, e.ds, sizeof(DevState), cudaMemcpyDeviceToHost));
  OSRL_CATCH
}
int osrl_state_load(osrl_engine* h, const void* host_buf, int64_t bytes) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_buf, "null argument");
  Engine& e = *h->e;
  const int64_t sec = e.plan.nP * (int64_t)sizeof(float);
  OSRL_REQUIRE(bytes >= (int64_t)sizeof(StateHeader), "state blob truncated");
  StateHeader hd;
  memcpy(&hd, host_buf, sizeof(hd));
  OSRL_REQUIRE(memcmp(hd.magic, "OSRLB200", 8) == 0, "not an osrl_b200 state blob");
  OSRL_REQUIRE(hd.abi == OSRL_ABI_VERSION && hd.algo == e.plan.cfg.algo && hd.nP == e.plan.nP &&
                   hd.dev_state_bytes == (int32_t)sizeof(DevState),
               "state blob was written by a different engine configuration / ABI");
  OSRL_REQUIRE(bytes >= (int64_t)sizeof(StateHeader) + 4 * sec + (int64_t)sizeof(DevState), "state blob truncated");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  const char* o = (const char*)host_buf + sizeof(hd);
  for (float* dst : {e.P, e.T, e.M, e.V}) { OSRL_CUDA(cudaMemcpy(dst, o, sec, cudaMemcpyHostToDevice)); o += sec; }
  OSRL_CUDA(cudaMemcpy(e.ds, o, sizeof(DevState), cudaMemcpyHostToDevice));
  OSRL_CATCH
}

static const char* kScalarNames[] = {"step", "pid_error_old", "pid_error_integral", "log_alpha", "n_train_steps",
                                     "log_temperature", "adam_t0", "adam_t1", "adam_t2", "adam_t3", "tau", "lmbda"};
static const int kNumScalars = 12;
int osrl_scalar_names(osrl_engine* h, const char** names, int cap, int* n) {
  OSRL_TRY
  OSRL_REQUIRE(h && n, "null argument");
  *n = kNumScalars;
  if (names)
    for (int i = 0; i < *n && i < cap; ++i) names[i] = kScalarNames[i];
  OSRL_CATCH
}
int osrl_scalars_get(osrl_engine* h, double* out, int cap, int* n) {
  OSRL_TRY
  OSRL_REQUIRE(h && out && n && cap >= kNumScalars, "bad argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  DevState d;
  OSRL_CUDA(cudaMemcpy(&d, e.ds, sizeof(d), cudaMemcpyDeviceToHost));
  *n = kNumScalars;
  out[10] = d.cop_tau; out[11] = d.cop_lmbda;   // COptiDICE: raw tau / lmbda (coptidice.py:104-105)
  out[0] = (double)d.step; out[1] = d.pid_e_old; out[2] = d.pid_e_int; out[3] = d.log_alpha;
  out[4] = d.n_train_steps; out[5] = d.log_temperature;
  for (int i = 0; i < 4; ++i) out[6 + i] = d.adam_t[i];
  OSRL_CATCH
}
int osrl_scalars_set(osrl_engine* h, const double* in, int n) {
  OSRL_TRY
  OSRL_REQUIRE(h && in && (n == 10 || n == kNumScalars), "bad argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  DevState d;
  OSRL_CUDA(cudaMemcpy(&d, e.ds, sizeof(d), cudaMemcpyDeviceToHost));
  d.step = d.vae_step = (unsigned long long)in[0]; d.pid_e_old = (float)in[1]; d.pid_e_int = (float)in[2];
  d.log_alpha = (float)in[3]; d.n_train_steps = (int)in[4]; d.log_temperature = in[5];
  for (int i = 0; i < 4; ++i) d.adam_t[i] = (int)in[6 + i];
  if (n == kNumScalars) { d.cop_tau = (float)in[10]; d.cop_lmbda = (float)in[11]; }
  OSRL_CUDA(cudaMemcpy(e.ds, &d, sizeof(d), cudaMemcpyHostToDevice));
  OSRL_CATCH
}

int osrl_noise_layout(osrl_engine* h, const char** names, int64_t* counts, int cap, int* n) {
  OSRL_TRY
  OSRL_REQUIRE(h && n, "null argument");
  *n = (int)h->e->plan.noise.size();
  for (int i = 0; i < *n && i < cap; ++i) {
    if (names) names[i] = h->e->plan.noise[i].first.c_str();
    if (counts) counts[i] = h->e->plan.noise[i].second;
  }
  OSRL_CATCH
}
int osrl_last_indices(osrl_engine* h, int64_t* host_out, int cap) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_out && cap >= h->e->B, "bad argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  OSRL_CUDA(cudaMemcpy(host_out, e.b_idx, (size_t)e.B * sizeof(int64_t), cudaMemcpyDeviceToHost));
  OSRL_CATCH
}
int osrl_last_noise(osrl_engine* h, int slot, float* host_out, int64_t cap) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_out, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(slot >= 0 && slot < (int)e.noise_buf.size(), "bad noise slot");
  OSRL_REQUIRE(cap >= e.plan.noise[slot].second, "noise buffer too small");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  OSRL_CUDA(cudaMemcpy(host_out, e.noise_buf[slot], (size_t)e.plan.noise[slot].second * sizeof(float),
                       cudaMemcpyDeviceToHost));
  OSRL_CATCH
}

int osrl_debug_linear(osrl_engine* h, const char* impl, int M, int N, int K, const float* A, const float* W,
                      const float* bias, int act, float* C) {
  OSRL_TRY
  OSRL_REQUIRE(h && impl && A && W && C && M > 0 && N > 0 && K > 0, "bad argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  const size_t before = e.allocs.size();
  float* dA = e.ws((size_t)M * K); float* dW = e.ws((size_t)N * K); float* db = e.ws(N); float* dC = e.ws((size_t)M * N);
  OSRL_CUDA(cudaMemcpy(dA, A, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice));
  OSRL_CUDA(cudaMemcpy(dW, W, (size_t)N * K * sizeof(float), cudaMemcpyHostToDevice));
  if (bias) OSRL_CUDA(cudaMemcpy(db, bias, (size_t)N * sizeof(float), cudaMemcpyHostToDevice));
  Lin l;
  l.w = 0; l.b = 0; l.in = K; l.out = N;
  GemmTask t = task_fwd(dA, K, M, dW, l, dC, N, act);
  t.bias = bias ? db : nullptr;
  Program prog;
  setenv("OSRL_GEMM", impl, 1);
  try { emit_gemm(e, prog, {t}); } catch (...) { unsetenv("OSRL_GEMM"); throw; }
  unsetenv("OSRL_GEMM");
  for (auto& op : prog.ops) op(e.cap_stream);
  OSRL_CUDA(cudaStreamSynchronize(e.cap_stream));
  OSRL_CUDA(cudaGetLastError());
  OSRL_CUDA(cudaMemcpy(C, dC, (size_t)M * N * sizeof(float), cudaMemcpyDeviceToHost));
  while (e.allocs.size() > before) { cudaFree(e.allocs.back()); e.allocs.pop_back(); }
  OSRL_CATCH
}

int osrl_debug_gemm(osrl_engine* h, const char* impl, int M, int N, int K, const float* A, int a_kc, const float* B,
                    int b_kc, float* C, float* colsum) {
  OSRL_TRY
  OSRL_REQUIRE(h && impl && A && B && C && M > 0 && N > 0 && K > 0, "bad argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  const size_t before = e.allocs.size();
  float* dA = e.ws((size_t)M * K); float* dB = e.ws((size_t)N * K); float* dC = e.ws((size_t)M * N); float* dS = e.ws(M);
  OSRL_CUDA(cudaMemcpy(dA, A, (size_t)M * K * sizeof(float), cudaMemcpyHostToDevice));
  OSRL_CUDA(cudaMemcpy(dB, B, (size_t)N * K * sizeof(float), cudaMemcpyHostToDevice));
  GemmTask t = blank_task();
  t.A = dA; t.a_kc = a_kc; t.lda = a_kc ? K : M;
  t.B = dB; t.b_kc = b_kc; t.ldb = b_kc ? K : N;
  t.C = dC; t.ldc = N;
  t.M = M; t.N = N; t.K = K;
  t.colsum = colsum ? dS : nullptr;
  Program prog;
  setenv("OSRL_GEMM", impl, 1);
  try { emit_gemm(e, prog, {t}); } catch (...) { unsetenv("OSRL_GEMM"); throw; }
  unsetenv("OSRL_GEMM");
  for (auto& op : prog.ops) op(e.cap_stream);
  OSRL_CUDA(cudaStreamSynchronize(e.cap_stream));
  OSRL_CUDA(cudaGetLastError());
  if (const char* reps_s = getenv("OSRL_DEBUG_TIME")) {   // kernel tuning aid: mean device time of the launch(es)
    const int reps = std::max(1, atoi(reps_s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, e.cap_stream);
    for (int r = 0; r < reps; ++r)
      for (auto& op : prog.ops) op(e.cap_stream);
    cudaEventRecord(e1, e.cap_stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[osrl_debug_gemm] %s %dx%dx%d a_kc=%d b_kc=%d: %.2f us/launch (%zu ops, %d back-to-back reps)\n", impl, M,
            N, K, a_kc, b_kc, ms * 1e3f / reps, prog.ops.size(), reps);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  if (e.fz_dbg) {
    std::vector<long long> h((size_t)e.fz_dbg_ctas * 64);
    OSRL_CUDA(cudaMemcpy(h.data(), e.fz_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int c = 0; c < std::min(e.fz_dbg_ctas, 3); ++c) {
      const long long* r = h.data() + (size_t)c * 64;
      fprintf(stderr, "[fz timeline cta %d] (cycles since start) setup %lld tables %lld |", c, r[1] - r[0], r[2] - r[0]);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " st%d %lld..%lld", k, r[8 + 2 * k] - r[0], r[9 + 2 * k] - r[0]);
      fprintf(stderr, " | loop_end %lld acc %lld ph1 %lld ph2 %lld ph3 %lld end %lld\n", r[3] - r[0], r[4] - r[0], r[5] - r[0],
              r[6] - r[0], r[7] - r[0], r[63] - r[0]);
      fprintf(stderr, "[fz timeline cta %d] mma:", c);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " k%d %lld..%lld", k, r[32 + 2 * k] - r[0], r[33 + 2 * k] - r[0]);
      fprintf(stderr, "\n");
    }
    e.fz_dbg = nullptr;
  }
  OSRL_CUDA(cudaMemcpy(C, dC, (size_t)M * N * sizeof(float), cudaMemcpyDeviceToHost));
  if (colsum) OSRL_CUDA(cudaMemcpy(colsum, dS, (size_t)M * sizeof(float), cudaMemcpyDeviceToHost));
  while (e.allocs.size() > before) { cudaFree(e.allocs.back()); e.allocs.pop_back(); }
  OSRL_CATCH
}

int osrl_debug_read(osrl_engine* h, int section, int64_t offset, int64_t count, float* host_out) {
  OSRL_TRY
  OSRL_REQUIRE(h && host_out, "null argument");
  Engine& e = *h->e;
  float* secs[5] = {e.P, e.T, e.G, e.M, e.V};
  OSRL_REQUIRE(section >= 0 && section < 5, "bad section");
  OSRL_REQUIRE(offset >= 0 && count >= 0 && offset + count <= e.plan.nP, "range outside the arena section");
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  OSRL_CUDA(cudaMemcpy(host_out, secs[section] + offset, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
  OSRL_CATCH
}

int osrl_profile(osrl_engine* h, int reps, int* n_ops, const char** names, double* ms, double* bytes, double* flops,
                 int cap, void* stream) {
  OSRL_TRY
  OSRL_REQUIRE(h && n_ops && reps >= 1, "bad argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)e.body.ops.size();
  *n_ops = n;
  if (!ms) return OSRL_OK;
  OSRL_REQUIRE(cap >= n, "profile buffers too small");
  // Time the launches INSIDE a captured graph: the step program in launch order (parallel branches serialised) with an
  // event-record node between consecutive launches, replayed like the real step -- so a launch's time carries the
  // graph's kernel-to-kernel edge, not the ~6 us CPU launch floor of eager timing.  OSRL_PROFILE_EAGER=1 (or a driver
  // that refuses to time event nodes) falls back to eager launches with an event pair around each.
  std::vector<double> acc(n, 0.0);
  bool done = false;
  cudaError_t why = cudaSuccess;
  if (!getenv("OSRL_PROFILE_EAGER")) {
    // (event-record nodes cannot be timed with cudaEventElapsedTime -- "invalid argument" -- so the graph carries
    // one-thread timestamp kernels (%globaltimer) between the launches; stamp[0] -> stamp[1] has nothing in between
    // and measures the cost of a stamp node itself, which is subtracted)
    cudaStream_t cs = e.cap_stream;
    const int64_t before = e.launches;
    unsigned long long* stamps = (unsigned long long*)e.ws((size_t)2 * (n + 3));
    cudaGraph_t g = nullptr;
    cudaGraphExec_t x = nullptr;
    bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      try {
        prologue(e, cs);
        k_stamp<<<1, 1, 0, cs>>>(stamps, 0);
        for (int i = 0; i < n; ++i) {
          k_stamp<<<1, 1, 0, cs>>>(stamps, i + 1);
          e.body.ops[i](cs);
        }
        k_stamp<<<1, 1, 0, cs>>>(stamps, n + 1);
        epilogue(e, cs);
      } catch (...) { ok = false; }
      const cudaError_t ce = cudaStreamEndCapture(cs, &g);
      if (ce != cudaSuccess) { why = ce; ok = false; }
    }
    e.launches = before;
    if (ok && g) {
      const cudaError_t ce = cudaGraphInstantiate(&x, g, 0);
      if (ce != cudaSuccess) { why = ce; ok = false; }
    } else ok = false;
    if (g) cudaGraphDestroy(g);
    std::vector<unsigned long long> hs(n + 2);
    for (int r = 0; ok && r < reps + 2; ++r) {
      ok = cudaGraphLaunch(x, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
      if (!ok || r < 2) continue;   // warm-up replays
      ok = cudaMemcpy(hs.data(), stamps, hs.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess;
      const double floor_ms = (double)(hs[1] - hs[0]) * 1e-6;
      for (int i = 0; ok && i < n; ++i) acc[i] += std::max(0.0, (double)(hs[i + 2] - hs[i + 1]) * 1e-6 - floor_ms);
    }
    if (x) cudaGraphExecDestroy(x);
    done = ok;
    if (!done) {
      fprintf(stderr, "[osrl_profile] in-graph timing unavailable (%s), falling back to eager launches\n",
              cudaGetErrorString(why != cudaSuccess ? why : cudaGetLastError()));
      std::fill(acc.begin(), acc.end(), 0.0);
    }
  }
  if (!done) {
    std::vector<cudaEvent_t> ev2(2 * n);
    for (auto& x : ev2) OSRL_CUDA(cudaEventCreate(&x));
    for (int r = 0; r < reps; ++r) {
      prologue(e, s);
      for (int i = 0; i < n; ++i) {
        OSRL_CUDA(cudaEventRecord(ev2[2 * i], s));
        e.body.ops[i](s);
        OSRL_CUDA(cudaEventRecord(ev2[2 * i + 1], s));
      }
      epilogue(e, s);
      OSRL_CUDA(cudaStreamSynchronize(s));
      for (int i = 0; i < n; ++i) {
        float t = 0.f;
        OSRL_CUDA(cudaEventElapsedTime(&t, ev2[2 * i], ev2[2 * i + 1]));
        acc[i] += t;
      }
    }
    for (auto& x : ev2) cudaEventDestroy(x);
  }
  e.profile_in_graph = done;
  for (int i = 0; i < n; ++i) {
    ms[i] = acc[i] / reps;
    if (names) names[i] = e.body.meta[i].name.c_str();
    if (bytes) bytes[i] = e.body.meta[i].bytes;
    if (flops) flops[i] = e.body.meta[i].flops;
  }
  OSRL_CATCH
}

int osrl_debug_fz_timelines(osrl_engine* h) {
  OSRL_TRY
  OSRL_REQUIRE(h, "null argument");
  Engine& e = *h->e;
  OSRL_CUDA(cudaSetDevice(e.device));
  OSRL_CUDA(cudaDeviceSynchronize());
  int idx = 0;
  for (auto& dbg : e.fz_dbg_all) {
    std::vector<long long> hbuf((size_t)dbg.ctas * 64);
    OSRL_CUDA(cudaMemcpy(hbuf.data(), dbg.buf, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long worst = 0;
    int wc = 0;
    for (int c = 0; c < dbg.ctas; ++c) {
      const long long d = hbuf[(size_t)c * 64 + 63] - hbuf[(size_t)c * 64];
      if (d > worst) { worst = d; wc = c; }
    }
    const long long* r = hbuf.data() + (size_t)wc * 64;
    fprintf(stderr, "[fz %2d %-22s ctas %4d] slowest cta %d: setup %lld tables %lld |", idx++, dbg.name.c_str(), dbg.ctas, wc,
            r[1] - r[0], r[2] - r[0]);
    for (int k = 0; k < 8; k += 1) fprintf(stderr, " %lld..%lld", r[8 + 2 * k] - r[0], r[9 + 2 * k] - r[0]);
    fprintf(stderr, " | loop_end %lld acc %lld ph1 %lld ph2 %lld ph3 %lld end %lld\n", r[3] - r[0], r[4] - r[0], r[5] - r[0],
            r[6] - r[0], r[7] - r[0], r[63] - r[0]);
  }
  OSRL_CATCH
}

int osrl_profile_was_in_graph(osrl_engine* h) { return h && h->e->profile_in_graph ? 1 : 0; }
int64_t osrl_launch_count(osrl_engine* h) { return h ? h->e->launches : 0; }
int osrl_launches_per_step(osrl_engine* h) { return h ? kernels_per_step(*h->e, true) : 0; }

int osrl_comm_unique_id(char out[128]) {
  OSRL_TRY
  OSRL_REQUIRE(out, "null argument");
  nccl::load();
  nccl::Uid id;
  nccl::check(nccl::GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out, id.internal, 128);
  OSRL_CATCH
}
// Map every peer's gradient section and flag block (dp_peer.cuh).  The cudaIpc handles travel over the communicator
// that was just created; peer mode is switched on only if EVERY rank mapped every peer (a rank that cannot -- no P2P
// path, IPC disabled -- votes it down for all, and the step keeps its NCCL collectives).  OSRL_DP=nccl skips it.
static void peer_setup(Engine& e) {
  const char* v = getenv("OSRL_DP");
  const bool want = !(v && strcmp(v, "nccl") == 0) && e.plan.cfg.algo != OSRL_ALGO_CDT && e.world <= DP_MAX_WORLD &&
                    nccl::AllGather != nullptr;
  struct Rec { cudaIpcMemHandle_t g, f; int ok, pad[3]; };
  const int W = e.world;
  Rec mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = want ? 1 : 0;
  if (want) {
    void* fl = nullptr;
    if (cudaMalloc(&fl, sizeof(DpFlags)) == cudaSuccess) {
      e.allocs.push_back(fl);
      e.dp_flags = (DpFlags*)fl;
      OSRL_CUDA(cudaMemset(fl, 0, sizeof(DpFlags)));
    } else mine.ok = 0;
    if (mine.ok && (cudaIpcGetMemHandle(&mine.g, e.G) != cudaSuccess || cudaIpcGetMemHandle(&mine.f, e.dp_flags) != cudaSuccess))
      mine.ok = 0;
    cudaGetLastError();
  }
  // marks the peers must find through their mappings (a handle that resolves to the wrong base would otherwise go
  // unnoticed): the flag block's magic word and, until set-up ends, the first gradient word
  uint32_t g0_saved = 0;
  const uint32_t mark = 0xD9000000u | (uint32_t)e.rank;
  if (mine.ok) {
    OSRL_CUDA(cudaMemcpy(&g0_saved, e.G, 4, cudaMemcpyDeviceToHost));
    OSRL_CUDA(cudaMemcpy(e.G, &mark, 4, cudaMemcpyHostToDevice));
    OSRL_CUDA(cudaMemcpy(&e.dp_flags->magic, &mark, 4, cudaMemcpyHostToDevice));
  }
  const bool marked = mine.ok != 0;
  if (!nccl::AllGather) return;   // every rank loads the same library: the same decision everywhere
  Rec* d = nullptr;
  OSRL_CUDA(cudaMalloc((void**)&d, sizeof(Rec) * (W + 1)));
  OSRL_CUDA(cudaMemcpy(d + W, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
  std::vector<Rec> all(W);
  auto gather = [&]() {
    nccl::check(nccl::AllGather(d + W, d, sizeof(Rec), /*ncclInt8*/ 0, e.comm, (cudaStream_t)0), "ncclAllGather");
    OSRL_CUDA(cudaStreamSynchronize(0));
    OSRL_CUDA(cudaMemcpy(all.data(), d, sizeof(Rec) * W, cudaMemcpyDeviceToHost));
  };
  gather();
  bool ok = true;
  for (auto& r : all) ok = ok && r.ok;
  DpPeers pr;
  memset(&pr, 0, sizeof(pr));
  pr.world = W; pr.rank = e.rank;
  if (ok) {
    for (int r = 0; r < W && ok; ++r) {
      if (r == e.rank) { pr.G[r] = e.G; pr.flags[r] = e.dp_flags; continue; }
      void *pg = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&pg, all[r].g, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      e.ipc_open.push_back(pg);
      if (cudaIpcOpenMemHandle(&pf, all[r].f, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      e.ipc_open.push_back(pf);
      pr.G[r] = (const float*)pg; pr.flags[r] = (DpFlags*)pf;
      uint32_t seen_g = 0, seen_f = 0;
      if (cudaMemcpy(&seen_g, pg, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(&seen_f, &((DpFlags*)pf)->magic, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
          seen_g != (0xD9000000u | (uint32_t)r) || seen_f != (0xD9000000u | (uint32_t)r))
        ok = false;
    }
    cudaGetLastError();
  }
  // second round: did every rank manage to map everybody?
  mine.ok = ok ? 1 : 0;
  OSRL_CUDA(cudaMemcpy(d + W, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
  gather();
  for (auto& r : all) ok = ok && r.ok;
  cudaFree(d);
  if (marked) OSRL_CUDA(cudaMemcpy(e.G, &g0_saved, 4, cudaMemcpyHostToDevice));   // (after every rank has looked)
  if (!ok) {
    for (void* q : e.ipc_open) cudaIpcCloseMemHandle(q);
    e.ipc_open.clear();
    cudaGetLastError();
    if (want && e.rank == 0) fprintf(stderr, "osrl_b200: peer-memory data parallelism unavailable, using NCCL collectives\n");
    return;
  }
  e.peers = pr;
  OSRL_CUDA(cudaMalloc((void**)&e.d_peers, sizeof(DpPeers)));
  e.allocs.push_back(e.d_peers);
  OSRL_CUDA(cudaMemcpy(e.d_peers, &pr, sizeof(pr), cudaMemcpyHostToDevice));
  e.peer_on = true;
}

int osrl_dp_mode(osrl_engine* h) {
  if (!h || h->e->world <= 1) return 0;
  return h->e->peer_on ? 2 : (h->e->comm ? 1 : 0);
}

int osrl_comm_init(osrl_engine* h, const char id[128], int world_size, int rank) {
  OSRL_TRY
  OSRL_REQUIRE(h && id, "null argument");
  Engine& e = *h->e;
  OSRL_REQUIRE(world_size == e.world && rank == e.rank, "world_size/rank differ from the engine config");
  OSRL_CUDA(cudaSetDevice(e.device));
  nccl::load();
  nccl::Uid u;
  memcpy(u.internal, id, 128);
  nccl::check(nccl::CommInitRank(&e.comm, world_size, u, rank), "ncclCommInitRank");
  if (e.pipelined && nccl::CommSplit)   // communicator of the pipelined VAE branch (absent: osrl_steps stays sequential)
    nccl::check(nccl::CommSplit(e.comm, 0, rank, &e.comm2, nullptr), "ncclCommSplit");
  peer_setup(e);
  OSRL_CATCH
}

}  // extern "C"
