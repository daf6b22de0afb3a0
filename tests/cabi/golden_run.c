/* golden_run.c — a plain C caller of the C ABI (include/lbft.h), the way a Rust `extern "C"` shim would call it.
 * Reproduces librabft-v2/tests/simulated_run.rs:45-94 (seed 52 / 3 nodes, seed 48 / 8 nodes) and exits non-zero on
 * any mismatch.  Built and run by tests/test_gpu_cabi_c.py with:  gcc golden_run.c -I include -L csrc -llbft_b200 */
#include <stdio.h>
#include <string.h>

#include "lbft.h"

static int run(uint64_t seed, uint32_t nodes, const uint32_t* want_counts, const uint64_t* want_states) {
  lbft_config c;
  memset(&c, 0, sizeof c);
  c.struct_size = sizeof c;
  c.num_instances = 1;
  c.num_nodes = nodes;
  c.delay_kind = LBFT_DELAY_LOGNORMAL;
  c.seeds = &seed;
  c.max_clock = 1000;            /* loop_until(GlobalTime(1000)) */
  c.delay_mean = 10.0;           /* RandomDelay::new(10.0, 4.0) */
  c.delay_variance = 4.0;
  c.target_commit_interval = 100000; /* NodeConfig of simulated_run.rs:30-36 */
  c.delta = 20;
  c.gamma = 2.0;
  c.lambda = 0.5;
  c.commands_per_epoch = 30000;
  lbft_sim* sim = NULL;
  if (lbft_create(&c, &sim) != LBFT_OK) { fprintf(stderr, "create: %s\n", lbft_last_error()); return 1; }
  if (lbft_run(sim) != LBFT_OK) { fprintf(stderr, "run: %s\n", lbft_last_error()); return 1; }
  uint32_t counts[64];
  uint64_t states[64];
  if (lbft_commit_counts(sim, counts) != LBFT_OK || lbft_last_states(sim, states) != LBFT_OK) return 1;
  int bad = 0;
  for (uint32_t n = 0; n < nodes; n++) {
    if (counts[n] != want_counts[n] || states[n] != want_states[n]) {
      fprintf(stderr, "node %u: %u commits, state %llu (want %u, %llu)\n", n, counts[n], (unsigned long long)states[n],
              want_counts[n], (unsigned long long)want_states[n]);
      bad = 1;
    }
  }
  lbft_commit log[64];
  size_t len = 0;
  if (lbft_commit_log(sim, 0, 0, log, 64, &len) != LBFT_OK || len != want_counts[0]) bad = 1;
  printf("seed %llu, %u nodes: %u commits on node 0, first command (proposer %u, index %u, time %lld)\n",
         (unsigned long long)seed, nodes, counts[0], log[0].proposer, log[0].index, (long long)log[0].time);
  lbft_destroy(sim);
  return bad;
}

int main(void) {
  const uint32_t c3[3] = {27, 27, 27};
  const uint64_t s3[3] = {11134312813757838303ULL, 11134312813757838303ULL, 11134312813757838303ULL};
  const uint32_t c8[8] = {28, 28, 28, 28, 28, 28, 28, 30};
  const uint64_t a = 12785928431398617538ULL, b = 4890275890002623733ULL;
  const uint64_t s8[8] = {a, a, a, a, a, a, a, b};
  if (lbft_abi_version() != LBFT_ABI_VERSION) return 2;
  int rc = run(52, 3, c3, s3) | run(48, 8, c8, s8);
  puts(rc ? "MISMATCH" : "golden runs reproduced through the C ABI");
  return rc;
}
