#!/bin/bash
# usage: tools/gpu_cfg4.sh TAG — BASELINE config 4: 64x2048 sweeps, full IP -> LO -> LM on one GPU (bench line with parity check and CPU
# legs) + ncu --set full on the projection and curvature kernels at that shape; and the default (64x1800) full bench line.
TAG=$1
mkdir -p gpurun_out
timeout 400 python bench.py --preset hdl64_2048 > gpurun_out/${TAG}_bench_hdl64_2048.json 2> gpurun_out/${TAG}_bench_hdl64_2048.err; echo "cfg4 bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench_hdl64_2048.json | head -14
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ip_project|ip_image|lo_curv_occl" --launch-skip 9 -c 3 -f -o /tmp/prof_$TAG \
    python bench.py --preset hdl64_2048 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --n-seq 128 > gpurun_out/${TAG}_ncu_cfg4.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_cfg4_raw.csv 2>> gpurun_out/${TAG}_ncu_cfg4.log
python tools/ncu_table.py gpurun_out/${TAG}_ncu_cfg4_raw.csv > gpurun_out/${TAG}_ncu_cfg4_projection_curvature.md; cat gpurun_out/${TAG}_ncu_cfg4_projection_curvature.md
timeout 400 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "default bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench_default.json | head -6
