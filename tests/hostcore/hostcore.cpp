// hostcore.cpp — TEST INFRASTRUCTURE: compiles the device state machine (csrc/sim_core.cuh) with g++
// so that its logic can be checked against the oracle in the CPU-only container.  It is never part
// of, linked into, or reachable from the product library (librabft_simulator_b200/csrc/liblbft_b200.so).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../librabft_simulator_b200/csrc/host_setup.hpp"
#include "../../librabft_simulator_b200/csrc/sim_core.cuh"

using namespace lbft;
static thread_local std::string g_err;

template <int NMAX, int QMODE, bool REC, bool RES, bool EP, bool TDS = false>
static void run_all_ep(const Params& P, std::vector<uint32_t>& state, const double* zx, const double* zf) {
  for (uint32_t inst = 0; inst < P.num_instances; inst++) {
    uint32_t tile = inst / 32, lane = inst % 32;
    TileMem<32> mem{state.data() + (size_t)tile * P.L.total_words * 32, lane};
    std::vector<uint32_t> sk(QMODE == 2 ? (size_t)P.L.queue_cap * 32 : 1);  // stands in for the shared-memory queue
    std::vector<uint16_t> sd(QMODE == 2 ? (size_t)P.L.queue_cap * 32 : 1);
    Core<TileMem<32>, NMAX, QMODE, false, REC, RES, 1, EP, TDS> core(P, mem, zx, zf, P.delay_thr, sk.data() + lane, sd.data() + lane);
    // QMODE 2: the stand-in for the shared-memory queue is per call, like shared memory is per launch
    if (RES && (P.run_flags & 1u)) core.restore_regs();
    else core.init(P.seeds[inst]);
    core.run();
    core.finalize(inst);
    if (RES) core.save_regs();
  }
}

// (the host harness instantiates the epoch machinery for every mode, so that the CPU tests cover epochs x recording x staged
// runs; the product library has it for plain kernels only)
template <int NMAX, int QMODE, bool REC, bool RES>
static void run_all(const Params& P, std::vector<uint32_t>& state, const double* zx, const double* zf) {
  if (P.L.tds) {
    if (!REC && !RES) run_all_ep<NMAX, QMODE, false, false, false, true>(P, state, zx, zf);  // (HostSetup refuses the other combinations)
  } else if (P.L.epochs > 1) run_all_ep<NMAX, QMODE, REC, RES, true>(P, state, zx, zf);
  else run_all_ep<NMAX, QMODE, REC, RES, false>(P, state, zx, zf);
}

// The instantiations with a compile-time layout (sim_core.cuh FX, sim_params.h fixed_layout) on the host: selected the way
// lbft_api.cu select_kernel does (minus the kernel family / tile / lane-group conditions, which do not exist here), so that
// the CPU tests of those shapes exercise the very state machine the BASELINE configurations run on the GPU.
template <int NMAX, int QMODE, int FX>
static void run_all_fx(const Params& P, std::vector<uint32_t>& state, const double* zx, const double* zf) {
  for (uint32_t inst = 0; inst < P.num_instances; inst++) {
    uint32_t tile = inst / 32, lane = inst % 32;
    TileMem<32> mem{state.data() + (size_t)tile * P.L.total_words * 32, lane};
    std::vector<uint32_t> sk(QMODE == 2 ? (size_t)P.L.queue_cap * 32 : 1);
    std::vector<uint16_t> sd(QMODE == 2 ? (size_t)P.L.queue_cap * 32 : 1);
    // (calendar queue: with the occupancy words in a stand-in for shared memory, as the sparse-tile / wide kernels hold them)
    constexpr bool KS = QMODE == 3;
    std::vector<uint32_t> km(KS ? (size_t)((P.L.cal_times + 7) / 8) * 32 : 1);
    Core<TileMem<32>, NMAX, QMODE, FX, false, false, 1, false, false, KS> core(P, mem, zx, zf, P.delay_thr, sk.data() + lane, sd.data() + lane);
    core.km = km.data() + lane;
    core.init(P.seeds[inst]);
    core.run();
    core.finalize(inst);
  }
}
static int fixed_selected(const Params& P) {
  constexpr Layout kDefault4 = fixed_layout(FX_DEFAULT4), kPart7 = fixed_layout(FX_PART7), kCommittee64 = fixed_layout(FX_COMMITTEE64);
  const bool table_delay = P.delay_kind == LBFT_DELAY_LOGNORMAL && !P.delay_const && P.delay_kmax != 0;
  const bool plain_model = table_delay && P.delay_kmax + 2 <= 256 && P.silent_mask == 0;
  if (P.record_rs || P.resumable) return FX_NONE;
  if (P.L.queue_scan == 2 && plain_model && memcmp(&P.L, &kDefault4, sizeof(Layout)) == 0) return FX_DEFAULT4;
  if (P.L.queue_scan == 3 && plain_model && memcmp(&P.L, &kPart7, sizeof(Layout)) == 0) return FX_PART7;
  if (P.L.queue_scan == 3 && table_delay && memcmp(&P.L, &kCommittee64, sizeof(Layout)) == 0) return FX_COMMITTEE64;
  return FX_NONE;
}

static int run_impl(const lbft_config* c, uint32_t* commit_counts, uint64_t* last_states, uint32_t* lc_round,
                    uint32_t* counters, uint32_t* status, uint32_t* words_per_instance, std::vector<uint32_t>& state,
                    Params& P, const int64_t* stops = nullptr, size_t nstops = 0);

static size_t read_round_switches(const Layout& L, const std::vector<uint32_t>& state, uint32_t num_nodes, uint32_t instance,
                                  lbft_round_switch* out, size_t cap) {
  const uint32_t row = L.round_cap + 1, tile = instance >> 5, lane = instance & 31;
  size_t k = 0;
  for (uint32_t node = 0; node < num_nodes; node++)
    for (uint32_t r = 1; r < row; r++) {  // slot 0 is the per-node maximum, not a switch
      const uint32_t w = state[((size_t)tile * L.total_words + rs_table_base(L) + (size_t)node * row + r) * 32 + lane];
      if (!w) continue;
      if (k < cap) out[k] = lbft_round_switch{node, r, (int64_t)(w - 1u)};
      k++;
    }
  return k;
}

extern "C" {
const char* hostcore_last_error(void) { return g_err.c_str(); }

// What the host setup decided for a configuration: [0] delay_kmax (0 = exp() fallback), [1] scan queue,
// [2] round_cap, [3] queue_cap, [4] payload_cap, [5] words per instance.
int hostcore_setup_info(const lbft_config* c, uint32_t* out6) {
  HostSetup hs;
  if (!hs.build(*c)) { g_err = hs.error; return LBFT_ERR_INVALID; }
  out6[0] = hs.params.delay_kmax;
  out6[1] = hs.params.L.queue_scan;
  out6[2] = hs.params.L.round_cap;
  out6[3] = hs.params.L.queue_cap;
  out6[4] = hs.params.L.payload_cap;
  out6[5] = hs.params.L.total_words;
  return LBFT_OK;
}

// Which compile-time-layout instantiation (sim_params.h FX_*) hostcore_run takes for this configuration; 0 = generic.
int hostcore_fixed_shape(const lbft_config* c) {
  HostSetup hs;
  if (!hs.build(*c)) { g_err = hs.error; return -1; }
  return getenv("HOSTCORE_NO_FIXED") ? (int)FX_NONE : fixed_selected(hs.params);
}

// Same outputs as the product's lbft_* getters; chain_out (optional) receives, per instance,
// round_cap * 2 words of the chain table and leader_out (optional) the leader table.
int hostcore_run(const lbft_config* c, uint32_t* commit_counts, uint64_t* last_states, uint32_t* lc_round,
                 uint32_t* counters, uint32_t* status, uint32_t* words_per_instance) {
  std::vector<uint32_t> state;
  Params P;
  return run_impl(c, commit_counts, last_states, lc_round, counters, status, words_per_instance, state, P);
}

// Same contract as the product's lbft_round_switches (include/lbft.h): runs the batch with the configuration's
// flags (LBFT_FLAG_ROUND_SWITCHES required) and reads one instance's table out of the tile layout.
int hostcore_round_switches(const lbft_config* c, uint32_t instance, lbft_round_switch* out, size_t cap, size_t* n) {
  if (!(c->flags & LBFT_FLAG_ROUND_SWITCHES)) { g_err = "LBFT_FLAG_ROUND_SWITCHES not set"; return LBFT_ERR_STATE; }
  if (instance >= c->num_instances) { g_err = "instance out of range"; return LBFT_ERR_INVALID; }
  const size_t IN = (size_t)c->num_instances * c->num_nodes;
  std::vector<uint32_t> cc(IN), lc(IN), counters((size_t)c->num_instances * 12), status(c->num_instances), state;
  std::vector<uint64_t> ls(IN);
  Params P;
  int rc = run_impl(c, cc.data(), ls.data(), lc.data(), counters.data(), status.data(), nullptr, state, P);
  if (rc != LBFT_OK) return rc;
  if (n) *n = read_round_switches(P.L, state, c->num_nodes, instance, out, cap);
  return LBFT_OK;
}
// lbft_run_until called once per stop (LBFT_FLAG_RESUMABLE required); outputs describe the state at the last stop.
int hostcore_run_staged(const lbft_config* c, const int64_t* stops, size_t nstops, uint32_t* commit_counts,
                        uint64_t* last_states, uint32_t* lc_round, uint32_t* counters, uint32_t* status) {
  std::vector<uint32_t> state;
  Params P;
  return run_impl(c, commit_counts, last_states, lc_round, counters, status, nullptr, state, P, stops, nstops);
}

int hostcore_round_switches_staged(const lbft_config* c, uint32_t instance, const int64_t* stops, size_t nstops,
                                   lbft_round_switch* out, size_t cap, size_t* n) {
  if (!(c->flags & LBFT_FLAG_ROUND_SWITCHES)) { g_err = "LBFT_FLAG_ROUND_SWITCHES not set"; return LBFT_ERR_STATE; }
  if (instance >= c->num_instances) { g_err = "instance out of range"; return LBFT_ERR_INVALID; }
  const size_t IN = (size_t)c->num_instances * c->num_nodes;
  std::vector<uint32_t> cc(IN), lc(IN), counters((size_t)c->num_instances * 12), status(c->num_instances), state;
  std::vector<uint64_t> ls(IN);
  Params P;
  int rc = run_impl(c, cc.data(), ls.data(), lc.data(), counters.data(), status.data(), nullptr, state, P, stops, nstops);
  if (rc != LBFT_OK) return rc;
  if (n) *n = read_round_switches(P.L, state, c->num_nodes, instance, out, cap);
  return LBFT_OK;
}
}  // extern "C"

static int run_impl(const lbft_config* c, uint32_t* commit_counts, uint64_t* last_states, uint32_t* lc_round,
                    uint32_t* counters, uint32_t* status, uint32_t* words_per_instance, std::vector<uint32_t>& state,
                    Params& P, const int64_t* stops, size_t nstops) {
  HostSetup hs;
  if (!hs.build(*c)) { g_err = hs.error; return LBFT_ERR_INVALID; }
  P = hs.params;
  P.seeds = c->seeds;
  P.zig_x = hs.zig_x.data();
  P.zig_f = hs.zig_f.data();
  P.leader = hs.leader.data();
  P.duration = hs.duration.data();
  P.period = hs.period.data();
  P.weights = hs.weights.data();
  P.delay_thr = hs.delay_thr.empty() ? nullptr : hs.delay_thr.data();
  uint32_t tiles = (c->num_instances + 31) / 32;
  state.assign((size_t)tiles * P.L.total_words * 32, 0xdeadbeefu);
  P.state = state.data();
  P.out_commit_counts = commit_counts;
  P.out_last_state = last_states;
  P.out_lc_round = lc_round;
  P.out_counters = counters;
  P.out_status = status;
  if (words_per_instance) *words_per_instance = P.L.total_words;
  if (stops && !P.resumable) { g_err = "LBFT_FLAG_RESUMABLE not set"; return LBFT_ERR_STATE; }
  const int64_t one_stop[1] = {P.max_clock};
  if (!stops) { stops = one_stop; nstops = 1; }
  // one launch per stop over the same state, like lbft_run_until
  for (size_t stage = 0; stage < nstops; stage++) {
    if (stops[stage] < 0 || stops[stage] > P.max_clock) { g_err = "stop_clock out of range"; return LBFT_ERR_INVALID; }
    P.stop_clock = (int32_t)stops[stage];
    P.run_flags = stage > 0 ? 1u : 0u;
#define RUN(NMAX, QS)                                                                                             \
  (P.resumable ? (P.record_rs ? run_all<NMAX, QS, true, true>(P, state, P.zig_x, P.zig_f)                           \
                              : run_all<NMAX, QS, false, true>(P, state, P.zig_x, P.zig_f))                          \
               : (P.record_rs ? run_all<NMAX, QS, true, false>(P, state, P.zig_x, P.zig_f)                          \
                              : run_all<NMAX, QS, false, false>(P, state, P.zig_x, P.zig_f)))
  const int fx = getenv("HOSTCORE_NO_FIXED") ? (int)FX_NONE : fixed_selected(P);
  if (fx == FX_DEFAULT4) run_all_fx<16, 2, FX_DEFAULT4>(P, state, P.zig_x, P.zig_f);
  else if (fx == FX_PART7) run_all_fx<16, 3, FX_PART7>(P, state, P.zig_x, P.zig_f);
  else if (fx == FX_COMMITTEE64) run_all_fx<64, 3, FX_COMMITTEE64>(P, state, P.zig_x, P.zig_f);
  else if (P.L.queue_scan == 2) RUN(16, 2);
  else if (P.L.queue_scan == 1) RUN(16, 1);
  else if (P.L.queue_scan == 3) {
    if (c->num_nodes <= 16) RUN(16, 3);
    else if (c->num_nodes <= 32) RUN(32, 3);
    else RUN(64, 3);
  } else if (c->num_nodes <= 16) RUN(16, 0);
  else if (c->num_nodes <= 32) RUN(32, 0);
  else RUN(64, 0);
#undef RUN
  }
  return LBFT_OK;
}
