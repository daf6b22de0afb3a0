// The reference's integration test (librabft-v2/tests/simulated_run.rs:19-94) against the GPU library: the same
// parameters (delta 20, gamma 2, lambda 0.5, target_commit_interval 100000, 30000 commands per epoch, LogNormal(10, 4)
// delay, loop_until(1000)) and the same golden values — only `make_simulator` builds a `GpuSimulator`.
use bft_lib::base_types::Duration;
use bft_lib::simulated_context::State;
use bft_lib::simulator::GlobalTime;
use bft_lib_gpu::GpuSimulator;
use librabft_v2::node::NodeConfig;

fn make_simulator(seed: u64, nodes: usize) -> GpuSimulator {
    let config = NodeConfig { target_commit_interval: Duration(100_000), delta: Duration(20), gamma: 2.0, lambda: 0.5 };
    GpuSimulator::new(seed, nodes, 10.0, 4.0, config, 30_000)
}

fn run(seed: u64, nodes: usize) -> (Vec<usize>, Vec<State>) {
    let mut sim = make_simulator(seed, nodes);
    let contexts = sim.loop_until(GlobalTime(1000));
    (contexts.iter().map(|c| c.committed_history().len()).collect(), contexts.iter().map(|c| c.last_committed_state()).collect())
}

#[test]
fn three_nodes_seed_52() {
    let (lens, states) = run(52, 3);
    assert_eq!(lens, vec![27, 27, 27]);
    assert_eq!(states, vec![State(11134312813757838303); 3]);
}

#[test]
fn eight_nodes_seed_48() {
    let (lens, states) = run(48, 8);
    assert_eq!(lens, vec![28, 28, 28, 28, 28, 28, 28, 30]);
    let mut want = vec![State(12785928431398617538); 7];
    want.push(State(4890275890002623733));
    assert_eq!(states, want);
}

#[test]
fn a_batch_is_the_same_as_its_instances_one_by_one() {
    let config = NodeConfig { target_commit_interval: Duration(100_000), delta: Duration(20), gamma: 2.0, lambda: 0.5 };
    let mut batch = GpuSimulator::new_batch((40..56).collect(), 3, 10.0, 4.0, config, 30_000);
    let counts = batch.loop_until_batch(GlobalTime(1000));
    assert_eq!(counts[12], vec![27, 27, 27]); // seed 52
    let (rows, lens) = batch.commit_logs(64);
    for (i, seed) in (40u64..56).enumerate() {
        let (one, _) = run(seed, 3);
        assert_eq!(counts[i], one);
        let single = make_simulator(seed, 3).loop_until(GlobalTime(1000));
        for n in 0..3 {
            let log = single[n].committed_history();
            assert_eq!(lens[i][n] as usize, log.len());
            for (k, (cmd, time)) in log.iter().enumerate() {
                assert_eq!((rows[i][k].proposer as usize, rows[i][k].index as usize, rows[i][k].time), (cmd.proposer.0, cmd.index, time.0));
            }
        }
    }
}
