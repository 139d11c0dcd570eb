#!/bin/bash
# compute-sanitizer memcheck over the tests that exercise the second session's kernels at small sizes: the 128-bit column
# accesses and the reordered |.|^2 accumulate of the axis engine, the native wavelength loop, the 4-stage / stream-K MDFT
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_s2_summary.txt
SEL='not 2048 and not 4096 and not large and not c2_ and not c3_'
TESTS="tests/test_gpu_tuned_axis.py tests/test_gpu_polychromatic.py tests/test_gpu_czt_intensity.py tests/test_gpu_coronagraph.py"
{
  echo "# compute-sanitizer $(compute-sanitizer --version | tail -1) on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
  echo "== compute-sanitizer --tool memcheck   pytest $TESTS -m gpu -k \"$SEL\""
  PB_SANITIZER=1 timeout 170 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
      python -m pytest $TESTS -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/sanitizer_s2_memcheck.log 2>&1
  echo "exit code $?"
  grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanitizer_s2_memcheck.log | tail -6
  grep -E "========= (Invalid|Race|Error|Warning|Hazard)" gpurun_out/sanitizer_s2_memcheck.log | sort | uniq -c | head -20
} > $OUT 2>&1
cat $OUT
