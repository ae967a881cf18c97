#!/bin/bash
# usage (one gpurun call at the start of a GPU session, ~4-5 min of box time):
#   gpurun --timeout 600 -- 'tools/gpu_session_start.sh r02a'
# Runs, in this order so that the cheap verdicts come first: the GPU test suite, smoke(), the default bench line, the sanitizer pass
# over the kernels added last (tools/sanitize_next_rows.sh), and the launch list of one bench step.  Everything lands in gpurun_out/.
TAG=${1:-gpu}
mkdir -p gpurun_out
timeout 180 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 240 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -12
timeout 300 tools/sanitize_next_rows.sh ${TAG}
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$?"
