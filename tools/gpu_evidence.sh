#!/bin/bash
# ncu evidence for every kernel a BASELINE configuration selects (one GPU): a --set full capture per kernel and the launch
# list of the default bench command.  Usage (under gpurun, repo root): bash tools/gpu_evidence.sh <tag>
tag=${1:-rX}
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1"
$NCU -k regex:lbft_event_loop -o gpurun_out/${tag}_cfg3_thread python tools/profile_one.py 3 65536 thread > gpurun_out/${tag}_cfg3.log 2>&1; tail -2 gpurun_out/${tag}_cfg3.log
$NCU -k regex:lbft_wide -o gpurun_out/${tag}_cfg4_wide python tools/profile_one.py 4 8192 > gpurun_out/${tag}_cfg4.log 2>&1; tail -2 gpurun_out/${tag}_cfg4.log
$NCU -k regex:lbft_wide -o gpurun_out/${tag}_cfg5_wide python tools/profile_one.py 5 16384 > gpurun_out/${tag}_cfg5.log 2>&1; tail -2 gpurun_out/${tag}_cfg5.log
$NCU -k regex:lbft_wide -o gpurun_out/${tag}_cfg2_wide python tools/profile_one.py 2 1024 > gpurun_out/${tag}_cfg2.log 2>&1; tail -2 gpurun_out/${tag}_cfg2.log
$NCU -k regex:lbft_wide -o gpurun_out/${tag}_cfg1_wide python tools/profile_one.py 1 1 > gpurun_out/${tag}_cfg1.log 2>&1; tail -2 gpurun_out/${tag}_cfg1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_launches_bench.log 2>&1
grep -c lbft gpurun_out/${tag}_launches.csv
