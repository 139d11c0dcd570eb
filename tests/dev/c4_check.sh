#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/c4_check.log
: > $L
timeout 600 python -m pytest tests/test_gpu_czt_intensity.py tests/test_gpu_polychromatic.py tests/test_gpu_zz_full_size_reference.py tests/test_gpu_tuned_axis.py -x -q -m gpu >> $L 2>&1; echo "pytest rc=$?" >> $L
timeout 200 python tools/bench_paths.py >> $L 2>&1
( REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_launches.csv python tools/profile_axis.py >> $L 2>&1 )
grep "axis_reg\|phase_screen" gpurun_out/c4_launches.csv | awk -F'","' '{print $5, $9, $NF}' | cut -c1-160 | tail -8 >> $L
tail -22 $L
