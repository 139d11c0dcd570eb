#!/bin/bash
# compute-sanitizer over the small-size GPU parity tests -> profiles/r02_sanitizer_summary.txt
#   bash tools/run_sanitizer.sh          (on a GPU box; ~10 min)
# memcheck: out-of-bounds / misaligned global, shared and local accesses of every kernel the tests launch;
# racecheck: shared-memory hazards (the FFT exchange buffers, the TMA-aliased staging areas, the GEMM rings).
# The BASELINE-size tests (c2 / c3) are left out: under the sanitizer they take tens of minutes and launch the same
# kernels as the tuned-kernel tests selected below.
cd "$(dirname "$0")/.."
OUT=profiles/r02_sanitizer_summary.txt
SEL='not c2_2048 and not c3_4096 and not 2048 and not 4096 and not large'
TESTS="tests/test_gpu_parity.py tests/test_gpu_tuned_axis.py tests/test_gpu_batched_focus.py"
{
  echo "# compute-sanitizer $(compute-sanitizer --version | tail -1) on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
  for tool in memcheck racecheck; do
    echo "== compute-sanitizer --tool $tool   pytest $TESTS -m gpu -k \"$SEL\""
    PB_SANITIZER=1 timeout 3000 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
        python -m pytest $TESTS -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_$tool.log 2>&1
    echo "exit code $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_$tool.log | tail -6
    grep -E "========= (Invalid|Race|Error|Warning|Hazard)" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -20
  done
} > $OUT 2>&1
cat $OUT
