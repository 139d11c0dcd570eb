#!/bin/bash
# one gpurun call: the CTA-pair MDFT kernel against the single-CTA kernel (bit-identical outputs expected) + C3 timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/mdft_pair.log
: > $L
for v in 1 0; do
  echo "== PB_MDFT_PAIR=$v" >> $L
  ( export PB_MDFT_PAIR=$v; timeout 120 python tests/dev/check_mdft_tc.py --big >> $L 2>&1 ); echo "rc=$?" >> $L
done
if [ -n "$NCU" ]; then
  ( timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_pair -c 2 -f -o gpurun_out/mdft_pair python tools/bench_mdft.py >> $L 2>&1 ); echo "ncu rc=$?" >> $L
fi
tail -40 $L
