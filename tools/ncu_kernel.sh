#!/bin/bash
# usage: tools/ncu_kernel.sh TAG KERNEL_REGEX [count] [n_seq] — ncu --set full on the launches matching the regex (steady state)
TAG=$1; RE=$2; CNT=${3:-2}; NSEQ=${4:-128}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$RE --launch-skip 8 -c $CNT -f -o /tmp/k_$TAG \
    python bench.py --steps 2 --warmup 3 --n-seq $NSEQ --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/k_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>> gpurun_out/${TAG}_ncu.log
ncu -i /tmp/k_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_source.csv 2>> gpurun_out/${TAG}_ncu.log
python tools/ncu_table.py gpurun_out/${TAG}_raw.csv
ls -la gpurun_out/${TAG}_*
