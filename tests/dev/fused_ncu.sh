#!/bin/bash
# one gpurun call: `ncu --set full` of one warmed-up launch of the fused focus kernel (16 fields) + a few timing variants
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/fused_ncu.log
: > $L
( export PB_FOCUS_V=3 REPS=2 $NCU_ENV; timeout 600 ncu --set full --clock-control none --import-source on -k regex:focus_fused -s 3 -c 1 -f -o gpurun_out/fused python tests/dev/check_fused.py >> $L 2>&1 ); echo "ncu rc=$?" >> $L
run() { echo "== $*" >> $L; ( export PB_FOCUS_V=3 "$@"; timeout 150 python tests/dev/check_fused.py >> $L 2>&1 ); echo "rc=$?" >> $L; }
for x in ${EXTRA}; do run $x; done
tail -30 $L
