#!/bin/bash
# compute-sanitizer passes over one iteration of the small-shape stress cases (scripts/stress.py): every tensor-core kernel
# of the path (bulk_dp, rnn_fwd, rnn_bwd, dw_mm16, decoder launches) plus the glue kernels.
# usage (GPU box): scripts/sanitize.sh > gpurun_out/sanitizer.txt 2>&1 ; the summary lines are copied to profiles/sanitizer_r02.txt
cd "$(dirname "$0")/.."
for tool in memcheck synccheck racecheck; do
  echo "=== compute-sanitizer --tool $tool python scripts/stress.py 1"
  timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 10 python scripts/stress.py 1 2>&1 | grep -v "^$" | tail -n 15
  echo "=== exit: $?"
done
