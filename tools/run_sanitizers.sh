#!/bin/bash
# compute-sanitizer over the reduced sweep (tools/sanitize_sweep.py); logs under gpurun_out/.
# Usage (on the GPU box): bash tools/run_sanitizers.sh [tools...]      default: memcheck racecheck initcheck synccheck
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck initcheck synccheck}
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
for t in $TOOLS; do
  extra=""
  [ "$t" = "memcheck" ] && extra="--leak-check no"
  [ "$t" = "racecheck" ] && extra="--racecheck-report all"
  timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool $t $extra --print-limit 40 \
      --log-file gpurun_out/sanitizer_$t.log python tools/sanitize_sweep.py > gpurun_out/sanitizer_$t.out 2>&1
  echo "$t exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$t.log | tail -1) $(tail -1 gpurun_out/sanitizer_$t.out)"
done
