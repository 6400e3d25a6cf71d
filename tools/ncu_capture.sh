#!/bin/bash
# One `ncu --set full` capture of the three hot kernels at full launch size (1184 candidates of the Log-MMSE
# 1024/128 min_tracking group), after the same command has run clean without ncu.  On the GPU box:
#   bash tools/ncu_capture.sh r02x       ->  gpurun_out/r02x_prof.ncu-rep (+ plain / ncu logs)
# Launch arithmetic: the regex also matches the clean-side align_kernel<true> of prepare_scoring (1 launch); the grid of
# the group splits by noise_mu into 3 sub-groups of 11 x 144 = 1584 candidates, each ONE launch per kernel (near-equal
# chunking), so the warm-up sweep is 9 launches and skipping 16 lands on an enhance / align / stoi triple of 1584 candidates.
tag=${1:-r02}
CMD="python tools/profile_group.py --alg omlsa --n-fft 1024 --hop 128 --method min_tracking --utts 11 --reps 1 --chunk 1184"
mkdir -p gpurun_out
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"enhance_kernel|stoi_stream|align_kernel" -s 16 -c 3 \
    -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?"
