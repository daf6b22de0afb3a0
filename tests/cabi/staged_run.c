/* staged_run.c — a plain C caller of the opt-in entry points of the C ABI (include/lbft.h): resumable runs
 * (lbft_run_until = loop_until called again, simulator.rs:380), a checkpoint carried into a second handle
 * (lbft_snapshot_*) and DataWriter's round switches (lbft_round_switches, data_writer.rs:34-50).
 * Seed 52 / 3 nodes of librabft-v2/tests/simulated_run.rs:45-66, stopped at 500 and continued to 1000.  Prints the
 * results in a form tests/test_gpu_cabi_c.py compares with the oracle; exits non-zero if the two handles disagree.
 * Built with:  gcc staged_run.c -I include -L csrc -llbft_b200 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lbft.h"

#define NODES 3
#define CHECK(call)                                                          \
  do {                                                                       \
    if ((call) != LBFT_OK) {                                                 \
      fprintf(stderr, "%s: %s\n", #call, lbft_last_error());                 \
      return 1;                                                              \
    }                                                                        \
  } while (0)

static lbft_sim* make(const uint64_t* seed) {
  lbft_config c;
  memset(&c, 0, sizeof c);
  c.struct_size = sizeof c;
  c.num_instances = 1;
  c.num_nodes = NODES;
  c.delay_kind = LBFT_DELAY_LOGNORMAL;
  c.seeds = seed;
  c.max_clock = 1000; /* the horizon: the largest clock lbft_run_until will be given */
  c.delay_mean = 10.0;
  c.delay_variance = 4.0;
  c.target_commit_interval = 100000;
  c.delta = 20;
  c.gamma = 2.0;
  c.lambda = 0.5;
  c.commands_per_epoch = 30000;
  c.flags = LBFT_FLAG_RESUMABLE | LBFT_FLAG_ROUND_SWITCHES;
  lbft_sim* sim = NULL;
  if (lbft_create(&c, &sim) != LBFT_OK) {
    fprintf(stderr, "create: %s\n", lbft_last_error());
    return NULL;
  }
  return sim;
}

int main(void) {
  const uint64_t seed_a = 52, seed_b = 53;
  lbft_sim *a = make(&seed_a), *b = make(&seed_b); /* b's own seed is never used: it continues a's checkpoint */
  if (!a || !b) return 1;
  CHECK(lbft_run_until(a, 500));
  uint32_t mid[NODES];
  CHECK(lbft_commit_counts(a, mid));
  size_t bytes = 0;
  CHECK(lbft_snapshot_size(a, &bytes));
  void* snap = malloc(bytes);
  if (!snap) return 1;
  CHECK(lbft_snapshot_save(a, snap, bytes));
  CHECK(lbft_snapshot_load(b, snap, bytes));
  free(snap);
  CHECK(lbft_run_until(a, 1000));
  CHECK(lbft_run_until(b, 1000));
  uint32_t ca[NODES], cb[NODES];
  uint64_t sa[NODES], sb[NODES];
  CHECK(lbft_commit_counts(a, ca));
  CHECK(lbft_commit_counts(b, cb));
  CHECK(lbft_last_states(a, sa));
  CHECK(lbft_last_states(b, sb));
  size_t na = 0, nb = 0;
  CHECK(lbft_round_switches(a, 0, NULL, 0, &na));
  lbft_round_switch* wa = malloc((na + 1) * sizeof *wa);
  lbft_round_switch* wb = malloc((na + 1) * sizeof *wb);
  if (!wa || !wb) return 1;
  CHECK(lbft_round_switches(a, 0, wa, na, &na));
  CHECK(lbft_round_switches(b, 0, wb, na, &nb));
  int bad = na != nb || memcmp(ca, cb, sizeof ca) || memcmp(sa, sb, sizeof sa) || memcmp(wa, wb, na * sizeof *wa);
  printf("snapshot_bytes %zu\n", bytes);
  printf("counts_at_500 %u %u %u\n", mid[0], mid[1], mid[2]);
  printf("counts %u %u %u\n", ca[0], ca[1], ca[2]);
  printf("states %llu %llu %llu\n", (unsigned long long)sa[0], (unsigned long long)sa[1], (unsigned long long)sa[2]);
  printf("switches %zu\n", na);
  for (size_t i = 0; i < na; i++) printf("switch %u %u %lld\n", wa[i].node, wa[i].round, (long long)wa[i].time);
  puts(bad ? "MISMATCH between the original and the restored handle" : "restored handle tracks the original");
  free(wa);
  free(wb);
  lbft_destroy(a);
  lbft_destroy(b);
  return bad;
}
