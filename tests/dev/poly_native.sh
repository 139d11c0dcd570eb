#!/bin/bash
# one gpurun call: the native wavelength loop (pb_polychromatic_czt) -- parity tests, then C4 timing against the Python loop
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/poly_native.log
: > $L
timeout 600 python -m pytest tests/test_gpu_polychromatic.py tests/test_gpu_zz_full_size_reference.py tests/test_gpu_czt_intensity.py tests/test_gpu_graphs.py -x -q -m gpu >> $L 2>&1; echo "pytest rc=$?" >> $L
for v in 1 0; do
  echo "== PB_POLY_NATIVE=$v" >> $L
  ( export PB_POLY_NATIVE=$v; timeout 200 python tools/bench_c4.py >> $L 2>&1 ); echo "rc=$?" >> $L
done
tail -25 $L
