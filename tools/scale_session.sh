#!/bin/bash
# usage (one gpurun --gpus 8 call): tools/scale_session.sh TAG — the driver's scaling run at HEAD: bench.py --steps 20 --warmup 5 at
# 8 / 4 / 2 / 1 ranks (rank -> GPU map interleaved over the box's two host I/O halves), no CPU legs.
TAG=${1:-sc}
mkdir -p gpurun_out
for N in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/${TAG}_weak$N.json 2> gpurun_out/${TAG}_weak$N.err
  echo "weak$N rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_weak$N.json | head -1
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/${TAG}_weak1.json 2> gpurun_out/${TAG}_weak1.err
echo "weak1 rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_weak1.json | head -1
