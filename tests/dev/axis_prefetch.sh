#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/axis_prefetch.log
: > $L
for v in 0 1 2 3; do
  echo "== PB_AXIS_PREFETCH=$v" >> $L
  ( export PB_AXIS_PREFETCH=$v; timeout 150 python tools/bench_paths.py 2>&1 | grep "fft2\|C5\|mtf" >> $L )
done
( export PB_AXIS_PREFETCH=3; timeout 300 python -m pytest tests/test_gpu_tuned_axis.py tests/test_gpu_fused_screen.py tests/test_gpu_zz_full_size_reference.py -x -q -m gpu >> $L 2>&1 ); echo "pytest(prefetch=3) rc=$?" >> $L
cat $L
