// oracle_selftest.cpp — the reference's scripted unit tests, transcribed against the oracle
// (TEST INFRASTRUCTURE).  Sources:
//   librabft-v2/src/unit_tests/record_store_tests.rs:7-292  (SharedRecordStore fixture + 7 tests)
//   librabft-v2/src/unit_tests/node_tests.rs:9-76           (Block + QC insertion moves the hqc)
//   bft-lib/src/unit_tests/simulated_context_tests.rs:38-129 (happened_before, compute/commit/epoch)
//   bft-lib/src/unit_tests/configuration_tests.rs:6-47       (count, pick_author KAT, quorum table)
//   bft-lib/src/unit_tests/simulator_tests.rs:6-12           (GlobalTime <-> NodeTime)
//   bft-lib/src/unit_tests/base_type_tests.rs:6-9            (Round + usize)
//   librabft-v2/src/unit_tests/util_tests.rs:6-21            (is_power2_minus1, merge_sort)
// Not transcribed, nothing of theirs is on the path: data_sync_tests.rs (serde_json round trips of the three message types —
// the oracle's messages are plain structs, never serialised), record_tests.rs (SignedValue / hash of a Block — the oracle
// identifies records by creation id), pacemaker_tests.rs (empty).
#include <cstdarg>
#include <string>

#include "lbft_oracle.hpp"

using namespace lbft_oracle;

namespace {
struct Report {
  std::string text;
  int failures = 0;
  void check(bool ok, const char* what, int line) {
    if (!ok) {
      failures++;
      char b[256];
      snprintf(b, sizeof b, "FAIL line %d: %s\n", line, what);
      text += b;
    }
  }
};
#define CHECK(cond) rep.check((cond), #cond, __LINE__)

// record_store_tests.rs:7-104
struct SharedRecordStore {
  LedgerIntern ledger;
  IdSource ids;
  Counters counters;
  std::map<Author, SimulatedContext> contexts;
  RecordStoreState store;
  SharedRecordStore(size_t num_nodes, uint64_t epoch_ttl) {
    for (size_t i = 0; i < num_nodes; i++) contexts.emplace((Author)i, SimulatedContext((Author)i, num_nodes, epoch_ttl, &ledger));
    StateId state = contexts.at(0).last_committed_state();
    store = RecordStoreState(QC_INITIAL, state, 0, contexts.at(0).configuration(state), &ids, &counters);
  }
  void create_timeout(Author a, Round r) { store.create_timeout(a, r, contexts.at(a)); }
  void propose_block(Author a, QcId prev, NodeTime clock) { store.propose_block(contexts.at(a), prev, clock); }
  bool create_vote(Author a, BlockId b) { return store.create_vote(contexts.at(a), b); }
  Author leader(Round r) const { return store.leader(r); }
  bool check_for_new_quorum_certificate() {
    Author a = leader(store.current_round());
    return store.check_for_new_quorum_certificate(contexts.at(a));
  }
  bool make_round(NodeTime clock) {
    Author a = leader(store.current_round());
    QcId prev = store.highest_quorum_certificate_hash();
    store.propose_block(contexts.at(a), prev, clock);
    BlockId proposed = *store.current_proposed_block;
    uint64_t threshold = store.configuration.quorum_threshold();
    bool ok = true;
    for (uint64_t i = 0; i < threshold; i++) ok &= create_vote((Author)i, proposed);
    ok &= check_for_new_quorum_certificate();
    return ok;
  }
  void make_tc() {
    uint64_t threshold = store.configuration.quorum_threshold();
    Round round = store.current_round();
    for (uint64_t i = 0; i < threshold; i++) create_timeout((Author)i, round);
  }
};

void test_initial_store(Report& rep) {  // :106-121
  SharedRecordStore s(2, 20);
  CHECK(s.store.blocks.size() == 0);
  CHECK(s.store.quorum_certificates.size() == 0);
  CHECK(s.store.highest_quorum_certificate_hash() == QC_INITIAL);
  CHECK(s.store.highest_quorum_certificate_round() == 0);
  CHECK(s.store.highest_timeout_certificate_round() == 0);
  CHECK(s.store.highest_committed_round() == 0);
  CHECK(s.store.current_round() == 1);
  CHECK(s.store.current_timeouts.size() == 0);
}
void test_propose_and_vote_no_qc(Report& rep) {  // :123-146
  SharedRecordStore s(2, 20);
  s.propose_block(0, QC_INITIAL, 1);
  s.propose_block(1, QC_INITIAL, 2);
  std::vector<BlockId> hashes;
  for (auto& kv : s.store.blocks) hashes.push_back(kv.first);
  CHECK(hashes.size() == 2);
  CHECK(s.create_vote(0, hashes[0]));
  CHECK(s.create_vote(0, hashes[0]));
  CHECK(s.create_vote(1, hashes[1]));
  CHECK(!s.check_for_new_quorum_certificate());
  CHECK(s.store.blocks.size() == 2);
  CHECK(s.store.quorum_certificates.size() == 0);
  CHECK(s.store.highest_quorum_certificate_hash() == QC_INITIAL);
  CHECK(s.store.highest_quorum_certificate_round() == 0);
  CHECK(s.store.highest_timeout_certificate_round() == 0);
  CHECK(s.store.highest_committed_round() == 0);
  CHECK(s.store.current_round() == 1);
  CHECK(s.store.current_timeouts.size() == 0);
}
void test_vote_with_quorum(Report& rep) {  // :148-165
  SharedRecordStore s(2, 20);
  s.propose_block(0, QC_INITIAL, 1);
  s.propose_block(1, QC_INITIAL, 2);
  CHECK(s.store.current_proposed_block.has_value());
  BlockId proposed = *s.store.current_proposed_block;
  CHECK(s.create_vote(0, proposed));
  CHECK(s.create_vote(1, proposed));
  CHECK(s.check_for_new_quorum_certificate());
  CHECK(s.store.blocks.size() == 2);
  CHECK(s.store.quorum_certificates.size() == 1);
  CHECK(s.store.highest_quorum_certificate_round() == 1);
  CHECK(s.store.highest_timeout_certificate_round() == 0);
  CHECK(s.store.highest_committed_round() == 0);
  CHECK(s.store.current_round() == 2);
  CHECK(s.store.current_timeouts.size() == 0);
}
void test_timeouts_no_tc(Report& rep) {  // :167-187
  SharedRecordStore s(2, 20);
  s.propose_block(1, QC_INITIAL, 2);
  s.create_timeout(0, 1);
  s.create_timeout(0, 1);
  s.create_timeout(1, 0);
  CHECK(s.store.blocks.size() == 1);
  CHECK(s.store.quorum_certificates.size() == 0);
  CHECK(s.store.highest_quorum_certificate_hash() == QC_INITIAL);
  CHECK(s.store.highest_quorum_certificate_round() == 0);
  CHECK(s.store.highest_timeout_certificate_round() == 0);
  CHECK(s.store.highest_committed_round() == 0);
  CHECK(s.store.current_round() == 1);
  CHECK(s.store.current_timeouts.size() == 1);
}
void test_timeouts_with_tc(Report& rep) {  // :189-217
  SharedRecordStore s(2, 20);
  s.propose_block(1, QC_INITIAL, 2);
  s.create_timeout(1, 0);
  s.create_timeout(0, 1);
  s.create_timeout(1, 1);
  s.create_timeout(1, 2);
  CHECK(s.store.blocks.size() == 1);
  CHECK(s.store.quorum_certificates.size() == 0);
  CHECK(s.store.highest_quorum_certificate_hash() == QC_INITIAL);
  CHECK(s.store.highest_quorum_certificate_round() == 0);
  CHECK(s.store.highest_timeout_certificate_round() == 1);
  CHECK(s.store.highest_committed_round() == 0);
  CHECK(s.store.current_round() == 2);
  CHECK(s.store.current_timeouts.size() == 1);
  s.create_timeout(0, 2);
  CHECK(s.store.blocks.size() == 1);
  CHECK(s.store.highest_timeout_certificate_round() == 2);
  CHECK(s.store.current_round() == 3);
  CHECK(s.store.current_timeouts.size() == 0);
}
void test_non_contiguous_qcs(Report& rep) {  // :219-234
  SharedRecordStore s(2, 20);
  CHECK(s.make_round(10));
  CHECK(s.make_round(20));
  s.make_tc();
  CHECK(s.make_round(40));
  CHECK(s.store.blocks.size() == 3);
  CHECK(s.store.quorum_certificates.size() == 3);
  CHECK(s.store.highest_quorum_certificate_round() == 4);
  CHECK(s.store.highest_timeout_certificate_round() == 3);
  CHECK(s.store.highest_committed_round() == 0);
  CHECK(s.store.current_round() == 5);
  CHECK(s.store.current_timeouts.size() == 0);
}
void test_commit(Report& rep) {  // :236-292
  SharedRecordStore s(2, 20);
  CHECK(s.make_round(10));
  s.make_tc();
  CHECK(s.make_round(30));
  CHECK(s.make_round(40));
  CHECK(s.make_round(50));
  s.make_tc();
  CHECK(s.store.blocks.size() == 4);
  CHECK(s.store.quorum_certificates.size() == 4);
  CHECK(s.store.highest_quorum_certificate_round() == 5);
  CHECK(s.store.highest_timeout_certificate_round() == 6);
  CHECK(s.store.highest_committed_round() == 3);
  CHECK(s.store.current_round() == 7);
  CHECK(s.store.current_timeouts.size() == 0);
  const QuorumCertificate* hcc = s.store.highest_commit_certificate();
  CHECK(hcc != nullptr);
  if (!hcc) return;
  CHECK(hcc->round == 5);
  CHECK(s.store.previous_round(hcc->certified_block_hash) == 4);
  CHECK(s.store.second_previous_round(hcc->certified_block_hash) == 3);
  auto commits = s.store.committed_states_after(0);
  CHECK(commits.size() == 2);
  if (commits.size() != 2) return;
  CHECK(commits[0].first == 1);
  CHECK(commits[1].first == 3);
  CHECK(hcc->committed_state.has_value() && *hcc->committed_state == commits[1].second);
}
void test_node(Report& rep) {  // node_tests.rs:9-76
  LedgerIntern ledger;
  IdSource ids;
  Counters counters;
  SimulatedContext context(0, 1, 2, &ledger);
  NodeConfig cfg;
  cfg.target_commit_interval = 0; cfg.delta = 0; cfg.gamma = 0; cfg.lambda = 0;  // NodeConfig::default()
  NodeState node1 = NodeState::make_initial_state(context, cfg, 0, &ids, &counters);
  StateId initial_state = context.last_committed_state();
  Command cmd = context.fetch();
  Block b0{ids.next_block++, cmd, 1, QC_INITIAL, 1, 0};
  std::optional<StateId> state = context.compute(initial_state, cmd, 1);
  CHECK(state.has_value());
  QuorumCertificate qc0{ids.next_qc++, 0, 1, b0.id, *state, std::nullopt, {0}, 0};
  node1.insert_network_record(0, Record::of(b0), context);
  node1.insert_network_record(0, Record::of(qc0), context);
  CHECK(node1.record_store.highest_quorum_certificate_hash() == qc0.id);
}
void test_simulated_context(Report& rep) {  // simulated_context_tests.rs:79-129
  LedgerIntern ledger;
  SimulatedContext context(0, 2, 2, &ledger);
  StateId s0 = context.last_committed_state();
  Command c1 = context.fetch(), c2 = context.fetch(), c3 = context.fetch();
  auto s1 = context.compute(s0, c1, 1);
  CHECK(s1 && context.read_epoch_id(*s1) == 0);
  auto s2 = context.compute(*s1, c2, 4);
  CHECK(s2 && context.read_epoch_id(*s2) == 1);
  auto s3 = context.compute(s0, c3, 3);
  CHECK(s3 && context.read_epoch_id(*s3) == 0);
  context.commit(*s1);
  context.commit(*s2);
  auto h = context.committed_history();
  CHECK(h.size() == 2);
  if (h.size() == 2) {
    CHECK(h[0].proposer == 0 && h[0].index == 0 && h[0].time == 1);
    CHECK(h[1].proposer == 0 && h[1].index == 1 && h[1].time == 4);
  }
  // committing a state that does not extend the last one by exactly one entry must fail (:172-174)
  bool threw = false;
  try { context.commit(*s3); } catch (const OracleError&) { threw = true; }
  CHECK(threw);
  // compute on a base that is neither pending nor the last committed state fails (:102-108,136)
  CHECK(!context.compute(s0, c3, 9).has_value());
}
void test_configuration(Report& rep) {  // configuration_tests.rs:6-47
  EpochConfiguration c({{0, 1}, {1, 2}, {2, 3}});
  CHECK(c.total_votes == 6);
  CHECK(c.weight(1) == 2);
  CHECK(c.weight(4) == 0);
  EpochConfiguration p({{0, 1}, {1, 2}, {2, 5}});
  std::map<Author, int> hits;
  for (uint64_t seed = 20; seed < 20 + p.total_votes; seed++) hits[p.pick_author(seed)]++;
  std::vector<int> r;
  for (auto& kv : hits) r.push_back(kv.second);
  std::sort(r.begin(), r.end());
  CHECK((r == std::vector<int>{1, 2, 5}));
  const uint64_t expect[6] = {1, 2, 3, 3, 4, 5};
  for (int n = 1; n <= 6; n++) {
    std::vector<std::pair<Author, uint64_t>> v;
    for (int i = 0; i < n; i++) v.push_back({i, 1});
    CHECK(EpochConfiguration(v).quorum_threshold() == expect[n - 1]);
  }
}
void test_time_conversion(Report& rep) {  // simulator_tests.rs:6-12
  const GlobalTime x = 15, start = 3;
  CHECK(to_node_time(x, start) == 12);
  CHECK(from_node_time(12, start) == x);
}
void test_round_plus_usize(Report& rep) {  // base_type_tests.rs:6-9
  CHECK(Round(3) + 4 == Round(7));
}
void test_util(Report& rep) {  // util_tests.rs:6-21
  CHECK(is_power2_minus1(1));
  CHECK(is_power2_minus1(3));
  CHECK(is_power2_minus1(7));
  CHECK(!is_power2_minus1(8));
  CHECK(!is_power2_minus1(2));
  auto cmp = [](uint64_t a, uint64_t b) { return a < b ? -1 : (a == b ? 0 : 1); };  // u64::cmp
  auto eq = [](uint64_t a, uint64_t b) { return a == b; };
  CHECK((merge_sort(std::vector<uint64_t>{0, 2, 6, 9}, std::vector<uint64_t>{2, 5, 6}, cmp, eq) == std::vector<uint64_t>{0, 2, 5, 6, 9}));
}
}  // namespace

extern "C" int lbfo_selftest(char* buf, size_t cap) {
  Report rep;
  try {
    test_initial_store(rep);
    test_propose_and_vote_no_qc(rep);
    test_vote_with_quorum(rep);
    test_timeouts_no_tc(rep);
    test_timeouts_with_tc(rep);
    test_non_contiguous_qcs(rep);
    test_commit(rep);
    test_node(rep);
    test_simulated_context(rep);
    test_configuration(rep);
    test_time_conversion(rep);
    test_round_plus_usize(rep);
    test_util(rep);
  } catch (const std::exception& e) {
    rep.failures++;
    rep.text += std::string("EXCEPTION: ") + e.what() + "\n";
  }
  if (buf && cap) {
    snprintf(buf, cap, "%s", rep.text.c_str());
  }
  return rep.failures;
}
