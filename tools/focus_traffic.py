"""profiles/r02_focus_traffic.json from an ncu --set full capture of the two focus kernels of one launch pair:
DRAM bytes read + written by the column and the row kernel, divided by the fields the pair processed.

    python tools/focus_traffic.py gpurun_out/r02_focus_final.ncu-rep 8 > profiles/r02_focus_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

rep, fields = sys.argv[1], int(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]


def bytes_of(d, key):
    v = float(d[key].replace(',', ''))
    u = units[hdr.index(key)].lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]


kern, total = {}, 0.0
for r in rows[2:]:
    d = dict(zip(hdr, r))
    rd, wr = bytes_of(d, 'dram__bytes_read.sum'), bytes_of(d, 'dram__bytes_write.sum')
    name = d['Kernel Name'].split('(')[0].replace('<unnamed>::', '')
    kern[name] = {'dram_read_bytes': rd, 'dram_write_bytes': wr, 'duration_us': float(d['gpu__time_duration.sum'].replace(',', ''))}
    total += rd + wr
print(json.dumps({'source': f'ncu --set full capture {rep.split("/")[-1]} of `bench.py --steps 1 --warmup 3 --calls-per-step 1 --no-extras` '
                            f'(one launch pair = {fields} fields): dram__bytes_read.sum + dram__bytes_write.sum of both kernels / {fields}',
                  'fields_per_launch_pair': fields, 'kernels': kern, 'dram_bytes_per_propagation': total / fields,
                  'algorithmic_bytes_per_propagation': 167772160}, indent=1))
