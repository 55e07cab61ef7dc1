/*
 * osrl_b200 -- C ABI of the B200-native OSRL training-step engine (libosrl_b200.so).
 *
 * The reference (liuzuxin/OSRL) is pure Python; its "FFI" for the per-step hot path is
 * the Python surface osrl.algorithms.*Trainer.train_one_step / osrl.common.*Dataset.
 * Every entry point below names the reference interface it replaces (paths relative to
 * the reference root).  Plain pointers and sizes only; no torch types.  All functions
 * return 0 on success or a negative code; osrl_last_error() holds the message
 * (thread-local).  One engine per (process, GPU); entry points are not re-entrant per
 * engine.  The library never falls back to the CPU: without a CUDA device
 * osrl_engine_create fails with OSRL_ERR_CUDA.
 */
#ifndef OSRL_B200_H_
#define OSRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSRL_ABI_VERSION 3

enum { OSRL_OK = 0, OSRL_ERR_ARG = -1, OSRL_ERR_CUDA = -2, OSRL_ERR_STATE = -3, OSRL_ERR_NCCL = -4,
       OSRL_ERR_UNSUPPORTED = -5 };

/* algorithm ids: osrl/algorithms/{bc,bcql,cpq,bearl,cdt,coptidice}.py */
enum { OSRL_ALGO_BC = 0, OSRL_ALGO_BCQL = 1, OSRL_ALGO_CPQ = 2, OSRL_ALGO_BEARL = 3, OSRL_ALGO_CDT = 4,
       OSRL_ALGO_COPTIDICE = 5 };

#define OSRL_MAX_HIDDEN 4

/* Constructor + trainer hyper-parameters of one algorithm:
 *   BC.__init__ bc.py:26-32, BCTrainer.__init__ bc.py:81-98
 *   BCQL.__init__ bcql.py:44-62, BCQLTrainer.__init__ bcql.py:262-281
 *   CPQ.__init__ cpq.py:38-54,  CPQTrainer.__init__ cpq.py:272-292
 *   BEARL.__init__ bearl.py:46-68, BEARLTrainer.__init__ bearl.py:369-387
 *   CDT.__init__ cdt.py:45-70, CDTTrainer.__init__ cdt.py:291-341
 *   COptiDICE.__init__ coptidice.py:68-123, COptiDICETrainer.__init__ coptidice.py:267-283
 * Fields an algorithm does not have are ignored. */
typedef struct osrl_config {
  int32_t algo;
  int32_t obs_dim, act_dim;
  float max_action;
  int32_t n_a_hidden, a_hidden[OSRL_MAX_HIDDEN];
  int32_t n_c_hidden, c_hidden[OSRL_MAX_HIDDEN];
  int32_t vae_hidden;
  int32_t sample_action_num;
  float gamma, tau, phi, lmbda, beta;
  float pid_kp, pid_ki, pid_kd;
  int32_t num_q, num_qc;
  float cost_limit;
  int32_t episode_len;
  float qc_scalar;                                            /* CPQ */
  float mmd_sigma, target_mmd_thresh;                         /* BEAR-Lag */
  int32_t num_samples_mmd_match, mmd_kernel /*0 gaussian, 1 laplacian*/, start_update_policy_step;
  float actor_lr, critic_lr, vae_lr, alpha_lr;
  /* CDT */
  int32_t seq_len, embedding_dim, num_layers, num_heads;
  float attention_dropout, residual_dropout, embedding_dropout;
  int32_t use_rew, use_cost, cost_transform, stochastic;
  float init_temperature, target_entropy;
  float learning_rate, weight_decay, adam_beta1, adam_beta2, clip_grad;
  int32_t lr_warmup_steps;
  float loss_cost_weight, loss_state_weight;
  /* engine */
  int32_t batch_size;  /* rows per rank per step */
  uint64_t seed;       /* Philox key of the on-device sampler / noise */
  int32_t world_size, rank;
  /* COptiDICE (ABI 3).  The two std vectors are HOST pointers, read once by osrl_engine_create. */
  int32_t f_type;                 /* 0 chi2, 1 softchi, 2 kl (get_f_div_fn, coptidice.py:15-38) */
  float init_state_propotion, alpha, cost_ub_epsilon;
  int32_t num_nu, num_chi;
  float scalar_lr;                /* Adam lr of tau and lmbda (coptidice.py:233-234) */
  const float* observations_std;  /* [obs_dim] */
  const float* actions_std;       /* [act_dim] */
} osrl_config;

/* One parameter tensor of model.state_dict() (names/shapes identical to the reference's,
 * e.g. "critic.q1_nets.0.2.weight"); `ptr` is a device pointer into the engine's arena,
 * valid until osrl_engine_destroy (NULL from osrl_plan). */
typedef struct osrl_param_desc {
  char name[96];
  int64_t rows, cols;      /* cols == 0: 1-D tensor of `rows` elements */
  int64_t offset;          /* float offset inside its arena section */
  int32_t section;         /* 0 trained parameter, 1 target ("*_old") copy, 2 buffer */
  int32_t group;           /* optimiser group index, -1 if none */
  float* ptr;
} osrl_param_desc;

/* Host-side view of a DSRL transition dataset: what TransitionDataset.__init__ receives
 * (dataset.py:803-820).  `done` may be NULL, then terminals|timeouts is used. */
typedef struct osrl_dataset_view {
  int64_t n;
  const float* observations;      /* [n, obs_dim] */
  const float* next_observations; /* [n, obs_dim] */
  const float* actions;           /* [n, act_dim] */
  const float* rewards;           /* [n] */
  const float* costs;             /* [n] */
  const float* done;              /* [n] float or NULL */
  const uint8_t* terminals;       /* [n] or NULL */
  const uint8_t* timeouts;        /* [n] or NULL */
  float reward_scale, cost_scale;
  const float* is_init;           /* [n] or NULL: TransitionDataset(state_init=True) (dataset.py:817-820), COptiDICE */
} osrl_dataset_view;

/* One collated transition minibatch = the positional arguments of
 * BCQLTrainer.train_one_step (bcql.py:283-284); BC uses observations+actions only
 * (bc.py:103).  Row-major float32; `on_host` != 0 means host pointers (copied H2D on the
 * given stream inside the call), else device pointers. */
typedef struct osrl_batch {
  int32_t rows;
  int32_t on_host;
  const float* observations;
  const float* next_observations;
  const float* actions;
  const float* rewards;
  const float* costs;
  const float* done;
  const float* is_init;           /* COptiDICE only: 7th element of its batch (coptidice.py:126-127); else NULL */
} osrl_batch;

/* Noise replay: raw standard-normal draws in the order the reference consumes them
 * (SURVEY.md Appendix B).  Slot names/sizes come from osrl_noise_layout.  NULL slots (or a
 * NULL osrl_noise) are generated on the device with Philox4x32-10.
 * CDT dropout (net.py:404-414, cdt.py:87,222): slots named "drop_*" hold the dropout multipliers themselves,
 * 0 or 1/(1-p) per element -- "drop_emb" [B,4T,E], per block i "drop_attn<i>" [B,H,4T,4T] (attention weights),
 * "drop_res<i>a" [B,4T,E] (after the attention projection), "drop_res<i>b" [B,4T,E] (after the MLP). */
#define OSRL_MAX_NOISE 32
typedef struct osrl_noise {
  int32_t on_host;
  const float* slot[OSRL_MAX_NOISE];
} osrl_noise;

/* One collated SequenceDataset minibatch = the positional arguments of CDTTrainer.train_one_step
 * (cdt.py:343-344): states [B,T,o], actions [B,T,a], returns [B,T], costs_return [B,T], time_steps [B,T]
 * int64, mask [B,T] (float32 here; the reference yields float64), episode_cost [B] (unused without
 * cost_prefix, may be NULL), costs [B,T]. */
typedef struct osrl_seq_batch {
  int32_t rows, seq_len;
  int32_t on_host;
  const float* states;
  const float* actions;
  const float* returns;
  const float* costs_return;
  const int64_t* time_steps;
  const float* mask;
  const float* episode_cost;
  const float* costs;
} osrl_seq_batch;

/* Host-side view of a trajectory dataset: what SequenceDataset holds after process_sequence_dataset
 * (dataset.py:137-183): per-transition arrays in trajectory order + CSR offsets, per-transition suffix sums
 * (returns / cost_returns), and the trajectory sampling distribution (compute_cost_sample_prob, dataset.py:439-459;
 * NULL = uniform). */
typedef struct osrl_seq_dataset_view {
  int64_t n, n_traj;
  const float* observations;   /* [n, obs_dim] */
  const float* actions;        /* [n, act_dim] */
  const float* returns;        /* [n] reward-to-go, unscaled */
  const float* cost_returns;   /* [n] cost-to-go, unscaled */
  const float* costs;          /* [n] */
  const int64_t* traj_offsets; /* [n_traj + 1] */
  const double* sample_prob;   /* [n_traj] or NULL */
  float reward_scale, cost_scale;
} osrl_seq_dataset_view;

typedef struct osrl_engine osrl_engine;

int osrl_abi_version(void);
const char* osrl_last_error(void);

/* Parameter plan without touching CUDA (names, shapes, offsets). */
int osrl_plan(const osrl_config* cfg, osrl_param_desc* out, int cap, int* n);

/* Replaces <Algo>.__init__ + <Algo>Trainer.__init__ + setup_optimizers.  Parameters start
 * at zero; fill them through the osrl_param_table pointers (the Python shim copies the
 * reference-order torch initialisation in), then call osrl_sync_targets (deepcopy of
 * actor/critic/cost_critic, bcql.py:100-105). */
int osrl_engine_create(const osrl_config* cfg, int device, osrl_engine** out);
void osrl_engine_destroy(osrl_engine* e);
int osrl_param_table(osrl_engine* e, osrl_param_desc* out, int cap, int* n);
int osrl_param_set(osrl_engine* e, int index, const float* host, int64_t count);
int osrl_param_get(osrl_engine* e, int index, float* host, int64_t count);
int osrl_sync_targets(osrl_engine* e);

/* Replaces TransitionDataset.__init__ (dataset.py:803-820): packs the dataset once into
 * HBM as rows [obs | next_obs | act | r*reward_scale | c*cost_scale | done]. */
int osrl_buffer_upload(osrl_engine* e, const osrl_dataset_view* view);
/* Replaces TransitionDataset.__prepare_sample + collate (dataset.py:832-842): gathers
 * `n` rows by index into six device/host outputs (bit-exact row copies). */
int osrl_gather(osrl_engine* e, const int64_t* idx, int n, int idx_on_host, osrl_batch* out /* writable ptrs */,
                void* stream);

/* Replaces SequenceDataset.__init__ residency (dataset.py:668-747) / __prepare_sample (dataset.py:749-775): packs
 * the trajectories once into HBM; osrl_seq_gather builds [n, T, .] windows for explicit (trajectory, start) pairs
 * (bit-exact, zero padded, mask, time_steps) into DEVICE buffers; osrl_seq_alias_table / osrl_last_sequences
 * expose the sampler's alias table and the pairs the last osrl_steps() step drew. */
int osrl_seq_buffer_upload(osrl_engine* e, const osrl_seq_dataset_view* view);
int osrl_seq_gather(osrl_engine* e, const int32_t* traj_idx, const int32_t* start_idx, int n, osrl_seq_batch* out,
                    void* stream);
int osrl_seq_alias_table(osrl_engine* e, float* prob_out, int32_t* alias_out, int cap);
/* Replaces process_sequence_dataset (dataset.py:137-183: a Python loop over every transition) and its two
 * discounted_cumsum calls (dataset.py:19-27, gamma = 1) on the device: `flat` is the raw DSRL dictionary (observations,
 * actions, rewards, costs, terminals and/or timeouts; next_observations / done / is_init ignored).  Episodes end where
 * terminals | timeouts is set, a trailing unfinished episode is dropped (as the reference does); reward-to-go and
 * cost-to-go are summed per episode backwards with one fp32 add per step -- the reference's order, bit-identical --
 * and the packed trajectory buffer osrl_seq_gather / osrl_steps read is left resident with a uniform episode
 * distribution.  cost_reverse: costs become 1 - cost (dataset.py:164-165).  reward_scale / cost_scale of the view are
 * applied to the stored returns as osrl_seq_buffer_upload does (dataset.py:762-763).
 * osrl_seq_episode_info: per-episode UNSCALED first return / cost return and the episode offsets [n_traj + 1] -- what
 * the caller needs for cost_sample / pf_sample (dataset.py:439-459, 390-430); any pointer may be NULL.
 * osrl_seq_set_sample_prob: Categorical over the episodes (normalised here; NULL = uniform). */
int osrl_seq_preprocess(osrl_engine* e, const osrl_dataset_view* flat, int cost_reverse, int64_t* n_traj_out,
                        int64_t* n_used_out);
int osrl_seq_episode_info(osrl_engine* e, float* first_return, float* first_cost_return, int64_t* offsets, int cap);
int osrl_seq_set_sample_prob(osrl_engine* e, const double* prob, int n);
int osrl_last_sequences(osrl_engine* e, int32_t* traj_out, int32_t* start_out, int cap);

/* Replaces <Algo>Trainer.train_one_step (bc.py:103-109, bcql.py:283-306, cpq.py:294-313,
 * bearl.py:389-412): all sub-updates + Polyak for one minibatch, no host sync. */
int osrl_step(osrl_engine* e, const osrl_batch* batch, const osrl_noise* noise, void* stream);
/* Replaces CDTTrainer.train_one_step (cdt.py:343-418): forward, losses, backward, clip_grad_norm_, AdamW with
 * LR warm-up, temperature Adam -- one sequence minibatch, no host sync. */
int osrl_step_seq(osrl_engine* e, const osrl_seq_batch* batch, const osrl_noise* noise_or_null, void* stream);
/* k steps with minibatches drawn on the device from the resident dataset (replaces the loop
 * body train_bcql.py:142-148 including DataLoader draw and .to(device)).  For the VAE algorithms (BCQ-Lag, CPQ,
 * BEAR-Lag) and k >= 2 the steps are software-pipelined: the VAE update of step s+1 (bcql.py:122-132), which depends
 * on nothing the rest of step s writes, runs on a second graph branch while the critic / actor updates of step s
 * read a snapshot of the VAE weights.  Every parameter sees the same kernels on the same data in the same order, so
 * the state after the call is bit-identical to k calls with k = 1 (tests/test_gpu_parity.py).  OSRL_PIPELINE=0
 * disables it. */
int osrl_steps(osrl_engine* e, int k, void* stream);
/* k steps on k explicit HOST minibatches: the batched form of osrl_step for a caller that keeps its own data loader
 * (the loop body train_bcql.py:142-148 with the reference's DataLoader left in place, k iterations per call).
 * `stacked` holds the k batches back to back per field (observations = [k][rows][obs_dim] floats, rewards = [k][rows],
 * ...; rows = batch_size, on_host = 1; pageable or pinned).  The library packs each batch into a pinned ring that
 * the step graph reads in place over PCIe (mapped memory), so the transfer of batch j+1 overlaps the compute of batch
 * j and no copy sits on the stream between steps; noise is Philox on the device (as osrl_step with a NULL noise);
 * the VAE algorithms are pipelined as in osrl_steps.  `stats_out` = [k][n_stats] floats (order of osrl_stat_names):
 * the call then returns after the k-th step has finished; NULL = return as soon as everything is queued (the caller's
 * buffers are already consumed).  State after the call is bit-identical to k x osrl_step (tests/test_gpu_parity.py). */
int osrl_steps_host(osrl_engine* e, const osrl_batch* stacked, int k, float* stats_out, void* stream);

/* Stats of the most recent step, in the order of osrl_stat_names (logger.store keys,
 * bcql.py:131,154,178,205-207).  Synchronises the stream. */
int osrl_stat_names(osrl_engine* e, const char** names, int cap, int* n);
int osrl_stats(osrl_engine* e, float* host_out, int cap, int* n, void* stream);

/* Scalar training state the reference keeps outside state_dict (PID error_old /
 * error_integral net.py:373-374, log_alpha cpq.py:93, step counters). */
int osrl_scalar_names(osrl_engine* e, const char** names, int cap, int* n);
int osrl_scalars_get(osrl_engine* e, double* host_out, int cap, int* n);
int osrl_scalars_set(osrl_engine* e, const double* host_in, int n);

/* Noise slots (name, float count) of this engine's algorithm at its batch size. */
int osrl_noise_layout(osrl_engine* e, const char** names, int64_t* counts, int cap, int* n);
/* Debug/parity: the sampled indices and the noise the last osrl_steps() step consumed. */
int osrl_last_indices(osrl_engine* e, int64_t* host_out, int cap);
int osrl_last_noise(osrl_engine* e, int slot, float* host_out, int64_t cap);

/* Debug/unit test: C[M,N] = act(A[M,K] * W[N,K]^T + bias) through one GEMM implementation
 * ("ffma" CUDA cores, "mma" 3xTF32 mma.sync, "tc5" tcgen05/TMEM where eligible); host pointers. */
int osrl_debug_linear(osrl_engine* e, const char* impl, int M, int N, int K, const float* A, const float* W,
                      const float* bias, int act, float* C);

/* Debug/unit test: C[M,N] = op(A) * op(B) through one GEMM implementation, any operand layout the step uses:
 * a_kc=1: A is [M,K] row-major (k contiguous), a_kc=0: A is [K,M]; b_kc=1: B is [N,K], b_kc=0: B is [K,N].
 * colsum (optional, [M]) receives sum_k A(i,k) -- the bias-gradient by-product of a weight-gradient GEMM.
 * Host pointers.  (Forward = 1,1; dgrad = 1,0; wgrad = 0,0.) */
int osrl_debug_gemm(osrl_engine* e, const char* impl, int M, int N, int K, const float* A, int a_kc, const float* B,
                    int b_kc, float* C, float* colsum);

/* Debug/parity: read `count` floats at float offset `offset` of an arena section
 * (0 params, 1 targets, 2 gradients of the last step, 3 Adam m, 4 Adam v). */
int osrl_debug_read(osrl_engine* e, int section, int64_t offset, int64_t count, float* host_out);

/* Lagged, sync-free statistics (the reference syncs per scalar with .item(), bcql.py:131-207):
 * enqueues an async copy of the CURRENT stats into an internal pinned buffer on `stream` and returns the stats of the
 * PREVIOUS call (whose copy has long completed) without synchronising the stream.  *valid = 0 on the first call. */
int osrl_stats_lagged(osrl_engine* e, float* host_out, int cap, int* n, int* valid, void* stream);

/* Resumable checkpoint (the reference saves {"model_state": state_dict} only, train_bcql.py:108-109; optimiser
 * moments, Polyak targets, PID / dual variables, Adam step counts and the Philox step counters are lost).
 * osrl_state_save writes one self-describing blob (header, P | T | M | V arena sections, device state) into a host
 * buffer of osrl_state_size bytes; osrl_state_load restores it into an engine built from the same osrl_config --
 * training then continues bit-identically. */
int osrl_state_size(osrl_engine* e, int64_t* bytes);
int osrl_state_save(osrl_engine* e, void* host_buf, int64_t cap);
int osrl_state_load(osrl_engine* e, const void* host_buf, int64_t bytes);

/* Per-launch timing of one step.  The step program is captured in launch order into a CUDA graph with an
 * event-record node between consecutive launches and replayed `reps` times (so a launch is timed as it runs
 * inside the step graph, not with the CPU launch floor of eager launches); returns, per launch, its name,
 * mean milliseconds and its algorithmic bytes / flops.  Call with ms == NULL to query the launch count.
 * osrl_profile_was_in_graph() tells whether the last call used the graph (1) or fell back to eager launches. */
int osrl_profile(osrl_engine* e, int reps, int* n_ops, const char** names, double* ms, double* bytes,
                 double* flops, int cap, void* stream);
int osrl_profile_was_in_graph(osrl_engine* e);

/* Number of kernels launched by this engine so far / per step. */
int64_t osrl_launch_count(osrl_engine* e);
int osrl_launches_per_step(osrl_engine* e);

/* Data-parallel: one engine per rank (one process per GPU); each optimiser update uses the gradient summed over the
 * ranks (losses are pre-scaled by 1/world), so N ranks x B rows equal one step on the concatenated batch.  id is an
 * ncclUniqueId made by rank 0.  osrl_comm_init creates the NCCL communicator(s) and then tries to map every peer's
 * gradient section over NVLink (cudaIpc): if every rank succeeds, BC / BCQ-Lag / CPQ / BEAR-Lag steps exchange
 * gradients through peer memory inside the Adam kernel (ordered sum: bit-identical replicas) and NCCL is no longer on
 * the step's path; otherwise (or with OSRL_DP=nccl, and always for CDT) the step graph holds ncclAllReduce nodes.
 * osrl_dp_mode: 0 = single GPU, 1 = NCCL collectives, 2 = peer memory. */
int osrl_comm_unique_id(char out[128]);
int osrl_comm_init(osrl_engine* e, const char id[128], int world_size, int rank);
int osrl_dp_mode(osrl_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* OSRL_B200_H_ */
