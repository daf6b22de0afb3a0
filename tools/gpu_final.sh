#!/bin/bash
# Last GPU session of a round, ordered by priority under a tight budget: the GPU suite, the bench line (e2e through run_stream),
# an A/B of alternatively built libraries (ab_libs/lib_<name>.so) on the calendar-queue configurations, and the parity +
# fuzz tests on the last of them.  Usage (repo root, under gpurun): bash tools/gpu_final.sh <tag> <variant> [<variant> ...]
tag=${1:-rX}; shift
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -2 gpurun_out/${tag}_pytest_gpu.txt
python bench.py --no-configs > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python tools/show_bench.py gpurun_out/${tag}_bench.json 2>/dev/null | head -2
if [ $# -gt 0 ]; then
  specs="main="; last=""
  for v in "$@"; do specs="$specs $v=ab_libs/lib_$v.so"; last=$v; done
  python tools/ab_libs.py $specs -- 5:16384 4:8192 > gpurun_out/${tag}_ab_calendar.txt 2>&1; cat gpurun_out/${tag}_ab_calendar.txt
  LBFT_LIB_PATH=$PWD/ab_libs/lib_$last.so timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_wide.py -m gpu -x -q > gpurun_out/${tag}_pytest_gpu_$last.txt 2>&1
  tail -2 gpurun_out/${tag}_pytest_gpu_$last.txt
fi
