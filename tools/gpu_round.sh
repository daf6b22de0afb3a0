#!/bin/bash
# One GPU session of a development round: the GPU suite, the default bench line (with the per-config block), the reference arm and
# one ncu --set full capture of the bench kernel.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-rX}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${tag}_pytest_gpu.txt
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python tools/show_bench.py gpurun_out/${tag}_bench.json
if [ "$2" != "noncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:lbft_event_loop -s 2 -c 1 -o gpurun_out/${tag}_bench_kernel \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  ls -la gpurun_out/${tag}_bench_kernel.ncu-rep
fi
