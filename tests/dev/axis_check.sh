#!/bin/bash
# one gpurun call: parity of the axis-engine paths after a kernel change + timing of the C5 / C4 paths per variant
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/axis_check.log
: > $L
timeout 600 python -m pytest tests/test_gpu_tuned_axis.py tests/test_gpu_fused_screen.py tests/test_gpu_czt_intensity.py tests/test_gpu_zz_full_size_reference.py -x -q -m gpu >> $L 2>&1; echo "pytest rc=$?" >> $L
run() { echo "== $*" >> $L; ( export "$@"; timeout 200 python tools/bench_paths.py >> $L 2>&1 ); echo "rc=$?" >> $L; }
run PB_NONE=1
for x in ${EXTRA}; do run $x; done
tail -40 $L
