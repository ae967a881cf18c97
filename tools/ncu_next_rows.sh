#!/bin/bash
# usage: tools/ncu_next_rows.sh TAG — ncu --set full on the kernels of the SURVEY §8f rows N2 / N4 (tools/next_rows_bench.py)
TAG=$1
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:icp_iterate --launch-skip 3 -c 3 -f -o /tmp/k_${TAG}_icp \
    python tools/next_rows_bench.py --only icp --quick > gpurun_out/${TAG}_icp_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lo_adjust -c 2 -f -o /tmp/k_${TAG}_dist \
    python tools/next_rows_bench.py --only distortion --quick > gpurun_out/${TAG}_dist_ncu.log 2>&1
for K in icp dist; do
  ncu -i /tmp/k_${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_${K}_raw.csv 2>> gpurun_out/${TAG}_${K}_ncu.log
  python tools/ncu_table.py gpurun_out/${TAG}_${K}_raw.csv > gpurun_out/${TAG}_${K}_table.md
  cat gpurun_out/${TAG}_${K}_table.md
done
