#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/mdft_bk.log
: > $L
run() { echo "== $*" >> $L; ( export "$@"; timeout 120 python tests/dev/check_mdft_tc.py --big >> $L 2>&1 ); echo "rc=$?" >> $L; }
run PB_MDFT_STREAMK=1
run PB_MDFT_STREAMK=0
run PB_MDFT_PAIR=1
cat $L
