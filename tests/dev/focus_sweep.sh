#!/bin/bash
# Focus-pipeline variants, one process each (the switches are read once per process).  Run on a GPU box:
#   bash tests/dev/focus_sweep.sh > gpurun_out/focus_sweep.log 2>&1
cd "$(dirname "$0")/../.."
run() { env "$@" timeout 300 python tests/dev/time_focus.py 2>&1 | grep -v "^$"; }
prof() { tag=$1; shift; env "$@" REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:focus_ --csv --log-file gpurun_out/launches_$tag.csv python tests/dev/time_focus.py > /dev/null 2>&1
  python - "$tag" <<'PY'
import csv, sys, collections
tag = sys.argv[1]
rows = list(csv.reader(l for l in open(f'gpurun_out/launches_{tag}.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
acc = collections.defaultdict(list)
for r in rows[1:]:
    acc[r[ki].split('(')[0][-40:]].append(float(r[vi].replace(',', '')))
for k, v in acc.items():
    print(f'  [{tag}] {k}: n={len(v)} mean {sum(v)/len(v)/1e3:.1f} us (min {min(v)/1e3:.1f})')
PY
}
echo "=== correctness"; PB_FOCUS_V=2 timeout 300 python tests/dev/check_focus.py --check 2>&1 | grep -E "^N=|batched =="
echo "=== timing"
run PB_FOCUS_V=1
run PB_FOCUS_V=2
run PB_FOCUS_V=2 PB_FOCUS_L2_HINTS=0
run PB_FOCUS_V=2 PB_FOCUS_BATCH=2
run PB_FOCUS_V=2 PB_FOCUS_BATCH=4
run PB_FOCUS_V=2 PB_FOCUS_BATCH=1
prof v1b8 PB_FOCUS_V=1
prof v2b8 PB_FOCUS_V=2
prof v2b2 PB_FOCUS_V=2 PB_FOCUS_BATCH=2
