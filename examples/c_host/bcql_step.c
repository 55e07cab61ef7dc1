/* A C host for libosrl_b200.so: the boundary is a plain C ABI (include/osrl_b200.h), so the step that replaces
 * BCQLTrainer.train_one_step (osrl/algorithms/bcql.py:283-306) can be driven without Python.
 *
 *   gcc -std=c99 -I include examples/c_host/bcql_step.c -L osrl_b200 -losrl_b200 -Wl,-rpath,$PWD/osrl_b200 -o bcql_step
 *
 * Prints the parameter table (works without a GPU: osrl_plan), then creates an engine, takes three steps on a random
 * host minibatch and prints the logged statistics.  Exit code 3 = no CUDA device (the library has no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "osrl_b200.h"

static float frand(void) { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }

int main(void) {
  osrl_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.algo = OSRL_ALGO_BCQL;
  cfg.obs_dim = 8; cfg.act_dim = 2; cfg.max_action = 1.f;
  cfg.n_a_hidden = 2; cfg.a_hidden[0] = cfg.a_hidden[1] = 256;
  cfg.n_c_hidden = 2; cfg.c_hidden[0] = cfg.c_hidden[1] = 256;
  cfg.vae_hidden = 400; cfg.sample_action_num = 10;
  cfg.gamma = 0.99f; cfg.tau = 0.005f; cfg.phi = 0.05f; cfg.lmbda = 0.75f; cfg.beta = 0.5f;
  cfg.pid_kp = 0.1f; cfg.pid_ki = 0.003f; cfg.pid_kd = 0.001f;
  cfg.num_q = 2; cfg.num_qc = 2; cfg.cost_limit = 10.f; cfg.episode_len = 300;
  cfg.actor_lr = 1e-3f; cfg.critic_lr = 1e-3f; cfg.vae_lr = 1e-3f;
  cfg.batch_size = 256; cfg.seed = 1; cfg.world_size = 1; cfg.rank = 0;

  int n = 0;
  if (osrl_plan(&cfg, NULL, 0, &n) != OSRL_OK) { fprintf(stderr, "osrl_plan: %s\n", osrl_last_error()); return 1; }
  osrl_param_desc* table = (osrl_param_desc*)calloc((size_t)n, sizeof(*table));
  if (osrl_plan(&cfg, table, n, &n) != OSRL_OK) { fprintf(stderr, "osrl_plan: %s\n", osrl_last_error()); return 1; }
  long long total = 0;
  for (int i = 0; i < n; ++i) total += (long long)table[i].rows * table[i].cols;
  printf("ABI %d: %d state_dict tensors, %lld parameters; first = %s [%lld x %lld]\n", osrl_abi_version(), n, total,
         table[0].name, (long long)table[0].rows, (long long)table[0].cols);

  osrl_engine* eng = NULL;
  if (osrl_engine_create(&cfg, 0, &eng) != OSRL_OK) {
    fprintf(stderr, "osrl_engine_create: %s\n", osrl_last_error());
    return 3;
  }
  const int B = cfg.batch_size, o = cfg.obs_dim, a = cfg.act_dim;
  float* buf = (float*)malloc(sizeof(float) * (size_t)B * (2 * o + a + 3));
  float *obs = buf, *nobs = obs + B * o, *act = nobs + B * o, *rew = act + B * a, *cost = rew + B, *done = cost + B;
  for (int i = 0; i < B * (2 * o + a); ++i) buf[i] = frand();
  for (int i = 0; i < B; ++i) { rew[i] = frand(); cost[i] = frand() > 0.8f; done[i] = 0.f; }
  osrl_batch batch;
  memset(&batch, 0, sizeof(batch));
  batch.rows = B; batch.on_host = 1;
  batch.observations = obs; batch.next_observations = nobs; batch.actions = act;
  batch.rewards = rew; batch.costs = cost; batch.done = done;
  for (int s = 0; s < 3; ++s)
    if (osrl_step(eng, &batch, NULL /* Philox noise on the device */, NULL /* default stream */) != OSRL_OK) {
      fprintf(stderr, "osrl_step: %s\n", osrl_last_error());
      return 1;
    }
  const char* names[16];
  float stats[16];
  int ns = 0;
  osrl_stat_names(eng, names, 16, &ns);
  if (osrl_stats(eng, stats, 16, &ns, NULL) != OSRL_OK) { fprintf(stderr, "osrl_stats: %s\n", osrl_last_error()); return 1; }
  for (int i = 0; i < ns; ++i) printf("%s = %g\n", names[i], stats[i]);
  printf("launches per step: %d\n", osrl_launches_per_step(eng));
  osrl_engine_destroy(eng);
  free(buf);
  free(table);
  return 0;
}
