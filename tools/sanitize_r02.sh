#!/bin/bash
# usage: tools/sanitize_r02.sh TAG — compute-sanitizer (memcheck, synccheck, racecheck) over GPU tests that drive the kernels written in
# round 2: record ordering (vox_order_warp / _cta / _wide, the all-in-one voxel_single), strip CCL, the voxel-row map index and its
# k-NN, lm_solve in both CTA shapes.  Small cases only (the tools slow kernels down 10-100x).  Run on the GPU box.
TAG=${1:-san2}
mkdir -p gpurun_out
SEL='test_voxel_grid[17-0.4] or test_voxel_grid[600-1.5] or test_voxel_grid[5000-0.8] or test_voxel_grid_adversarial_record_order[3000-organ_pipe] or test_voxel_grid_adversarial_record_order[2500-few_voxels] or test_ip_and_features_bit_exact[0] or test_ip_and_features_bit_exact[1] or test_scan_to_map_parity[6000-30000-2-20-True] or test_full_pipeline_sequence[0-1] or test_async_submit_collect_matches_sync'
for TOOL in memcheck synccheck racecheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $TOOL --error-exitcode 9 \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_${TOOL}.log 2>&1
  echo "$TOOL rc=$?"
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard|Invalid|Barrier error" gpurun_out/${TAG}_${TOOL}.log | sort | uniq -c | sort -rn | head -8
done
