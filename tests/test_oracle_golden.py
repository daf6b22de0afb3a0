"""The oracle against every golden vector / known-answer test the reference holds for the hot path
(SURVEY.md §8c).  CPU only."""
import ctypes

import numpy as np

U64 = np.uint64


def test_selftest_scripted_reference_unit_tests(oracle):
    # record_store_tests.rs:106-292, node_tests.rs:9-76, simulated_context_tests.rs:79-129,
    # configuration_tests.rs:6-47 transcribed in oracle/oracle_selftest.cpp
    failures, text = oracle.selftest()
    assert failures == 0, text


def test_siphash_empty_state(oracle):
    # README.md:27 `initial_state: State(13646096770106105413)` = SipHash-1-3(k=0) of the empty Vec (len 0u64)
    assert oracle.lib.lbfo_siphash13(bytes(8), 8) == 13646096770106105413
    assert oracle.state_key([]) == 13646096770106105413


def test_golden_3_nodes(oracle):
    # librabft-v2/tests/simulated_run.rs:45-66
    r = oracle.run([52], 3, 1000)
    assert r.commit_counts.tolist() == [[27, 27, 27]]
    assert r.last_states.tolist() == [[11134312813757838303] * 3]
    assert r.status.tolist() == [1]


def test_golden_8_nodes(oracle):
    # librabft-v2/tests/simulated_run.rs:68-94
    r = oracle.run([48], 8, 1000)
    assert r.commit_counts.tolist() == [[28] * 7 + [30]]
    assert r.last_states.tolist() == [[12785928431398617538] * 7 + [4890275890002623733]]


def test_golden_logs_hash_to_states(oracle):
    for seed, n in ((52, 3), (48, 8)):
        r = oracle.run([seed], n, 1000)
        for node in range(n):
            log = oracle.commit_log([seed], n, 0, node, 1000)
            assert len(log) == r.commit_counts[0, node]
            assert oracle.state_key(log) == r.last_states[0, node]
            # command indices of one proposer increase along the log; times are the proposers' local clocks
            for p in range(n):
                idx = [i for (q, i, _) in log if q == p]
                assert idx == sorted(idx)


def test_pick_author_kat(oracle):
    # configuration_tests.rs:16-29: weights 1/2/5, seeds 20..27 -> sorted hit counts [1, 2, 5]
    w = np.array([1, 2, 5], dtype=U64)
    hits = {}
    for seed in range(20, 28):
        a = oracle.lib.lbfo_pick_author(ctypes.c_void_p(w.ctypes.data), 3, seed)
        hits[a] = hits.get(a, 0) + 1
    assert sorted(hits.values()) == [1, 2, 5]


def test_quorum_thresholds(oracle):
    # configuration_tests.rs:39-47
    for n, q in zip(range(1, 7), (1, 2, 3, 3, 4, 5)):
        w = np.ones(n, dtype=U64)
        assert oracle.lib.lbfo_quorum_threshold(ctypes.c_void_p(w.ctypes.data), n) == q
    w = np.array([1 + (i % 3) for i in range(64)], dtype=U64)  # BASELINE config 4: total 127 -> 85
    assert oracle.lib.lbfo_quorum_threshold(ctypes.c_void_p(w.ctypes.data), 64) == 85


def test_xoshiro_reference_vector(oracle):
    # Xoshiro256** from state [1,2,3,4] gives 11520, 0, 1509978240, 1215971899390074240 (rand_xoshiro's
    # own test vector); here the seeding goes through SplitMix64, so check the seeding arithmetic instead:
    out = np.zeros(4, dtype=U64)
    oracle.lib.lbfo_xoshiro_seq(0, ctypes.c_void_p(out.ctypes.data), 4)
    # SplitMix64(0) first outputs (public reference values)
    s = [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f, 0xf88bb8a8724c81ec]

    def rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & (2**64 - 1)

    exp = []
    for _ in range(4):
        exp.append((rotl((s[1] * 5) & (2**64 - 1), 7) * 9) & (2**64 - 1))
        t = (s[1] << 17) & (2**64 - 1)
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
    assert out.tolist() == exp


def test_ziggurat_table_literals(oracle):
    # First literals of rand_distr's ZIG_NORM_X / ZIG_NORM_F
    x = np.zeros(257); f = np.zeros(257)
    oracle.lib.lbfo_ziggurat_tables(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(f.ctypes.data))
    assert x[0] == float("3.910757959537090045")
    assert x[1] == float("3.654152885361008796")
    assert x[2] == float("3.449278298560964462")
    assert f[0] == float("0.000477467764586655")
    assert x[256] == 0.0 and f[256] == 1.0
    assert np.all(np.diff(x) < 0) and np.all(np.diff(f) > 0)


def test_fixed_delay_truncates_to_mean(oracle):
    # BASELINE config 1 "fixed 10 ms": LogNormal(10, variance 0) -> exp(ln 10) truncates to 10, draws still consumed
    out = np.zeros(64, dtype=np.int64)
    oracle.lib.lbfo_delay_samples(7, 10.0, 0.0, ctypes.c_void_p(out.ctypes.data), 64)
    assert set(out.tolist()) == {10}


def test_lognormal_moments(oracle):
    out = np.zeros(200000, dtype=np.float64)
    oracle.lib.lbfo_normal_samples(123, ctypes.c_void_p(out.ctypes.data), out.size)
    assert abs(out.mean()) < 0.01 and abs(out.std() - 1.0) < 0.01
    d = np.zeros(200000, dtype=np.int64)
    oracle.lib.lbfo_delay_samples(5, 10.0, 4.0, ctypes.c_void_p(d.ctypes.data), d.size)
    assert abs(d.mean() - 9.5) < 0.05  # truncation removes ~0.5 from the mean of 10


def test_shuffle_is_permutation(oracle):
    for seed in range(20):
        v = np.arange(7, dtype=np.uint32)
        oracle.lib.lbfo_shuffle(seed, ctypes.c_void_p(v.ctypes.data), 7)
        assert sorted(v.tolist()) == list(range(7))


def test_config1_fixed_delay_100_rounds(oracle):
    # BASELINE config 1: 1 instance, 3 authors, fixed 10 ms delay, >= 100 rounds
    r = oracle.run([52], 3, 3000, delay_variance=0.0)
    assert r.counters[0, 6] >= 100
    assert r.status[0] == 1
    assert len(set(r.last_states[0].tolist())) <= 2


def test_prefix_consistency_many_seeds(oracle):
    # Safety: every node's log is a prefix of the longest one (simulated_context.rs:172-174 invariant)
    seeds = list(range(200, 232))
    for n in (3, 4, 7):
        r = oracle.run(seeds, n, 1000)
        assert (r.status == 1).all()
        for i in (0, 13, 31):
            logs = [oracle.commit_log(seeds, n, i, node, 1000) for node in range(n)]
            longest = max(logs, key=len)
            for lg in logs:
                assert lg == longest[: len(lg)]


def test_native_build_matches_the_portable_one(oracle):
    """The CPU-baseline leg of bench.py times the -O3 -march=native build of the oracle; it must remain the
    same function: goldens and a seed sweep, bit for bit."""
    from tests.support import Oracle, assert_same
    native = Oracle(native=True)
    r = native.run([52], 3, 1000)
    assert r.commit_counts.tolist() == [[27, 27, 27]] and r.last_states.tolist() == [[11134312813757838303] * 3]
    r = native.run([48], 8, 1000)
    assert r.last_states.tolist() == [[12785928431398617538] * 7 + [4890275890002623733]]
    seeds = np.arange(900, 1100, dtype=np.uint64)
    for nodes, kw in ((4, {}), (7, dict(partition_windows=4, partition_max_len=150)), (4, dict(delay_kind=1, delay_lo=5, delay_hi=15))):
        assert_same(oracle.run(seeds, nodes, 1000, **kw), native.run(seeds, nodes, 1000, **kw), "native vs portable N=%d" % nodes)
