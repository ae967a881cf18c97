#!/bin/bash
# usage: tools/ncu_launches_only.sh TAG — launch list (device time + DRAM bytes per launch) of ~two steady-state steps at n_seq 256
TAG=$1
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 330 -c 270 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --n-seq 256 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/ncu_launch_table.py gpurun_out/${TAG}_launches.csv 256 hdl64_1800 3 gpurun_out/${TAG}_ncu_traffic.json > gpurun_out/${TAG}_launch_table.md
cat gpurun_out/${TAG}_launch_table.md
