#!/bin/bash
# usage (one gpurun --gpus 8 call): tools/multi_gpu_session.sh TAG — H2D topology probe, weak scaling at 8 / 4 (interleaved and
# identity rank -> GPU maps) / 2 ranks, and BASELINE config 5 (8 x 64x2048 sequences in total, graph mode) at 8 / 4 / 2 / 1 ranks.
TAG=${1:-mg}
mkdir -p gpurun_out
run() {  # N out extra-env args...
  local N=$1 OUT=$2 ENVV=$3; shift 3
  env $ENVV timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/${TAG}_${OUT}.json 2> gpurun_out/${TAG}_${OUT}.err
  echo "$OUT rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_${OUT}.json | head -1
}
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 200 python tools/h2d_probe.py > gpurun_out/${TAG}_h2d_probe.json 2> gpurun_out/${TAG}_h2d_probe.err; echo "probe rc=$?"
run 8 weak8 X=1 --steps 8 --warmup 3 --no-parity-check
run 4 weak4_interleaved X=1 --steps 8 --warmup 3 --no-parity-check
run 4 weak4_identity ALEGO_BENCH_DEVICE_ORDER=identity --steps 8 --warmup 3 --no-parity-check
run 2 weak2_interleaved X=1 --steps 8 --warmup 3 --no-parity-check
run 8 cfg5_8 X=1 --preset hdl64_2048 --total-seq 8 --graphs --steps 20 --warmup 5
run 4 cfg5_4 X=1 --preset hdl64_2048 --total-seq 8 --graphs --steps 20 --warmup 5
run 2 cfg5_2 X=1 --preset hdl64_2048 --total-seq 8 --graphs --steps 20 --warmup 5
timeout 300 python bench.py --no-cpu-baseline --preset hdl64_2048 --total-seq 8 --graphs --steps 20 --warmup 5 > gpurun_out/${TAG}_cfg5_1.json 2> gpurun_out/${TAG}_cfg5_1.err; echo "cfg5_1 rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_cfg5_1.json | head -1
