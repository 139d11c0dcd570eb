#!/bin/bash
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
echo "=== correctness"; timeout 300 python tests/dev/check_focus.py --check 2>&1 | grep -E "^N=|batched =="
echo "=== timing"
run PB_ROW_V=2
run PB_ROW_V=4
run PB_ROW_V=4 PB_FOCUS_BATCH=4
run PB_ROW_V=4 PB_FOCUS_BATCH=2
run PB_ROW_V=4 PB_FOCUS_BATCH=16
prof r4b8 PB_ROW_V=4
