#!/bin/bash
# usage: tools/gpu_quick.sh TAG [bench args...] — perf iteration: only the bench (no tests, no CPU legs), per-kernel table to stdout
TAG=$1; shift
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-parity-check "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json | head -${ROWS:-24}
