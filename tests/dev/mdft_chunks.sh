#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/mdft_chunks.log
: > $L
for v in 1 0; do for c in ${CHUNKS:-2 8 64}; do
  ( export PB_MDFT_PAIR=$v PB_MDFT_CHUNK=$c; timeout 120 python tools/bench_mdft.py >> $L 2>&1 ); echo "rc=$?" >> $L
done; done
cat $L
