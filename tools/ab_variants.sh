#!/bin/bash
# A/B of prebuilt library variants on one GPU box: tools/ab_variants.sh <utts> variants/a.so variants/b.so ...
# Each variant is copied over the in-tree product library (the box copy is scratch) and timed with
# tools/kernel_breakdown.py; summaries land in gpurun_out/ab_<name>.log.
utts=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
    name=$(basename "$v" .so)
    cp "$v" classical_speech_enhancement_b200/libcse_sm100a.so
    python tools/kernel_breakdown.py --utts "$utts" > "gpurun_out/ab_${name}.log" 2>&1
    echo "== $name: $(grep 'total step' gpurun_out/ab_${name}.log) $(tail -1 gpurun_out/ab_${name}.log)"
done
