#!/bin/bash
# usage: tools/ncu_step_r02.sh TAG [n_seq_for_full_set] [regex]
# (1) launch list of two steady-state steps of the default bench workload (n_seq 256) with device time AND DRAM bytes per launch
#     (three cheap metrics: two passes per kernel) -> CSV, per-kernel table and the traffic JSON bench.py reads;
# (2) ncu --set full (+ source pages) for the kernels matching the regex only, at a smaller batch (every pass of a full-set
#     capture saves and restores the device memory the kernel can touch — at 256 sequences that is seconds per kernel).
TAG=$1; NSEQ=${2:-64}; SRC=${3:-"ccl_strip|mr_fill|mr_bbox|ip_image|ip_project|lo_curv_occl|vox_order_warp|vox_order_cta"}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 330 -c 270 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH --n-seq 256 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/ncu_launch_table.py gpurun_out/${TAG}_launches.csv 256 hdl64_1800 3 gpurun_out/${TAG}_ncu_traffic.json > gpurun_out/${TAG}_launch_table.md
head -60 gpurun_out/${TAG}_launch_table.md
timeout 420 ncu --set full --clock-control none --import-source on -k regex:"$SRC" --launch-skip 30 -c 10 -f -o /tmp/prof_$TAG $BENCH --n-seq $NSEQ > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full set rc=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_raw.csv 2>> gpurun_out/${TAG}_ncu_full.log
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_ncu_source.csv 2>> gpurun_out/${TAG}_ncu_full.log
python tools/ncu_table.py gpurun_out/${TAG}_ncu_raw.csv > gpurun_out/${TAG}_ncu_selected_kernels.md
cat gpurun_out/${TAG}_ncu_selected_kernels.md
ls -la gpurun_out/${TAG}_*
