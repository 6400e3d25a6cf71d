#!/bin/bash
# AddressSanitizer+UBSan and ThreadSanitizer over the kernel sources on the CPU thread emulation
# (tests/emu/sanitize_main.cpp explains why: compute-sanitizer is closed on the GPU pool).
# Usage: bash tools/run_cpu_sanitizers.sh [asan|tsan ...]   logs -> profiles/r02_cpu_<tool>.log
set -u
cd "$(dirname "$0")/.."
mkdir -p build
COMMON="-O1 -g -std=c++17 -DCSE_EMU -x c++ -pthread -ffp-contract=fast -Wno-unknown-pragmas -I tests/emu -I include"
for t in ${@:-asan tsan}; do
  case $t in
    asan) FLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer" ;;
    tsan) FLAGS="-fsanitize=thread" ;;
  esac
  g++ $COMMON $FLAGS tests/emu/sanitize_main.cpp -o build/sanitize_$t || exit 1
  env CSE_EMU_BLOCKS=2 TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" \
      ASAN_OPTIONS="detect_leaks=0" setarch "$(uname -m)" -R build/sanitize_$t > profiles/r02_cpu_$t.log 2>&1
  rc=$?
  ctl=$([ $t = asan ] && echo control-oob || echo control-race)
  env TSAN_OPTIONS="halt_on_error=0" ASAN_OPTIONS="detect_leaks=0:halt_on_error=0" setarch "$(uname -m)" -R build/sanitize_$t $ctl > profiles/r02_cpu_${t}_control.log 2>&1
  echo "$t negative control ($ctl): $(grep -c 'WARNING: ThreadSanitizer\|ERROR: AddressSanitizer' profiles/r02_cpu_${t}_control.log) report(s) (must be > 0)"
  echo "$t exit=$rc $(grep -c 'WARNING: ThreadSanitizer\|ERROR: AddressSanitizer\|runtime error' profiles/r02_cpu_$t.log) reports; $(tail -2 profiles/r02_cpu_$t.log | tr '\n' ' ')"
done
