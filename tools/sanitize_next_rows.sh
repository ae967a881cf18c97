#!/bin/bash
# usage: tools/sanitize_next_rows.sh TAG — compute-sanitizer (memcheck, racecheck) over the GPU tests of the SURVEY §8f rows N2 / N4 and
# the empty-map bootstrap (tests/test_next_rows.py, the N2 / N4 golden tests).  The kernels these cover were added after the round-1
# sanitizer session (profiles/r01_compute_sanitizer.md).  Run on the GPU box; logs land in gpurun_out/.
TAG=${1:-san}
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 \
      python -m pytest tests/test_next_rows.py tests/test_next_rows_golden.py -m gpu -x -q \
      > gpurun_out/${TAG}_${TOOL}.log 2>&1
  echo "$TOOL rc=$?"
  grep -E "ERROR SUMMARY|passed|failed|Race|Invalid" gpurun_out/${TAG}_${TOOL}.log | tail -5
done
