#!/bin/bash
# sustained (0.6 s, power-capped) timing: two-kernel pipeline vs the fused kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/fused_sustained.log
: > $L
run() { echo "== $*" >> $L; ( export REPS=600 "$@"; timeout 150 python tests/dev/check_fused.py >> $L 2>&1 ); echo "rc=$?" >> $L; }
run PB_FOCUS_V=2
run PB_FOCUS_V=3 PB_FUSED_NCOL=61 PB_FUSED_PREFETCH=1
run PB_FOCUS_V=3 PB_FUSED_NCOL=58 PB_FUSED_PREFETCH=1
run PB_FOCUS_V=3 PB_FUSED_NCOL=61 PB_FUSED_PREFETCH=1 B=32
run PB_FOCUS_V=2 B=32
cat $L
