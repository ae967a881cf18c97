#!/bin/bash
# usage: tools/ncu_full_only.sh TAG N_SEQ REGEX [skip] [count] — ncu --set full (+ SASS source page) for the kernels matching REGEX only
TAG=$1; NSEQ=${2:-64}; SRC=$3; SKIP=${4:-30}; CNT=${5:-10}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$SRC" --launch-skip $SKIP -c $CNT -f -o /tmp/prof_$TAG $BENCH --n-seq $NSEQ > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full set rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>> gpurun_out/${TAG}_ncu_full.log
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_ncu_source.csv 2>> gpurun_out/${TAG}_ncu_full.log
python tools/ncu_table.py gpurun_out/${TAG}_ncu_raw.csv > gpurun_out/${TAG}_ncu_selected_kernels.md
cat gpurun_out/${TAG}_ncu_selected_kernels.md
ls -la gpurun_out/${TAG}_*
