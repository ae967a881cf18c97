#!/bin/bash
# usage: tools/ncu_step.sh TAG [n_seq]  — ncu --set full over one whole pipeline step; exports the raw page as CSV
# (the .ncu-rep itself is too large for gpurun_out) plus the launch list with gpu__time_duration.
TAG=$1; NSEQ=${2:-128}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -s 130 -c 45 -f -o /tmp/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --n-seq $NSEQ --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>> gpurun_out/${TAG}_ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --n-seq $NSEQ --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ls -la gpurun_out/${TAG}_* 
