#!/bin/bash
# one gpurun call: stream-K MDFT against the classic kernel (accuracy + C3 timing), then the tests that use the MDFT
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/mdft_sk.log
: > $L
for v in 1 0; do
  echo "== PB_MDFT_STREAMK=$v" >> $L
  ( export PB_MDFT_STREAMK=$v; timeout 120 python tests/dev/check_mdft_tc.py --big >> $L 2>&1 ); echo "rc=$?" >> $L
  ( export PB_MDFT_STREAMK=$v; timeout 120 python tools/bench_mdft.py >> $L 2>&1 ); echo "rc=$?" >> $L
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_coronagraph.py tests/test_gpu_graphs.py -x -q -m gpu >> $L 2>&1; echo "pytest rc=$?" >> $L
( timeout 120 python tools/bench_coronagraph.py >> $L 2>&1 ); echo "rc=$?" >> $L
tail -60 $L
