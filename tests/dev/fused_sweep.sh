#!/bin/bash
# one gpurun call: correctness of the fused focus kernel, then timing of the split / ring variants, each in its own process
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/fused_sweep.log
: > $L
run() { echo "== $*" >> $L; ( export PB_FOCUS_V=3 "$@"; timeout 150 python tests/dev/check_fused.py $ARGS >> $L 2>&1 ); echo "rc=$?" >> $L; }
if [ -z "$SKIP_CHECK" ]; then ARGS=--check run PB_FOCUS_V=3; ARGS=; run PB_FOCUS_V=2; fi
ARGS=
for nc in ${NCOLS:-52 58}; do run PB_FUSED_NCOL=$nc; done
for x in ${EXTRA:-PB_FUSED_RING=2}; do run $x; done
tail -40 $L
