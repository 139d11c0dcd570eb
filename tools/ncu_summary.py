"""Text summary of an ncu report (run where ncu is installed; no GPU needed):

    python tools/ncu_summary.py gpurun_out/r02_focus.ncu-rep [--hot 12] > profiles/r02_ncu_focus_summary.txt

Per profiled launch: duration, DRAM bytes, occupancy limits, pipe / issue / l1tex / DRAM utilisation, shared-memory
wavefronts and bank conflicts, the stall reasons per issue, and the source lines that collect the most stall samples."""
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def page(rep, name, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    hot = int(sys.argv[sys.argv.index('--hot') + 1]) if '--hot' in sys.argv else 10
    rows = page(rep, 'raw')
    hdr, units = rows[0], rows[1]
    print(f'# {rep}: {len(rows) - 2} profiled launches (ncu --set full --clock-control none)')
    names = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        names.append(d['Kernel Name'])
        print(f"\n== {d['Kernel Name']}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] not in ('', 'n/a'):
                print(f'   {k:78s} {d[k]:>18s} {units[hdr.index(k)]}')
        st = []
        for h in hdr:
            if 'average_warps_issue_stalled' in h and 'per_issue_active' in h and d[h] not in ('', 'n/a'):
                st.append((float(d[h].replace(',', '')), h.split('stalled_')[1].split('_per_')[0]))
        st.sort(reverse=True)
        print('   stalls per issue: ' + ', '.join(f'{n} {v:.2f}' for v, n in st[:8]))
    # source page: blocks of ("Kernel Name", name) / header / one row per SASS instruction
    blocks, cur = [], None
    for r in page(rep, 'source'):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1] if len(r) > 1 else '', 'hdr': None, 'rows': []}
            blocks.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = r
        elif cur is not None and len(r) == len(cur['hdr']):
            cur['rows'].append(r)
    seen = set()
    for b in blocks:
        m = re.search(r'(\w+)\s*(<\(|\()', b['name'].replace('<unnamed>', ''))
        base = m.group(1) if m else '?'
        if base in seen or not b['rows']:
            continue
        seen.add(base)
        h = b['hdr']
        ix = {k: i for i, k in enumerate(h)}

        def num(r, k):
            try:
                return float(r[ix[k]])
            except (ValueError, KeyError):
                return 0.0
        tot = sum(num(r, '# Samples') for r in b['rows']) or 1.0
        print(f"\n-- {base}: {len(b['rows'])} SASS instructions, {int(tot)} stall samples; top lines")
        for r in sorted(b['rows'], key=lambda r: -num(r, '# Samples'))[:hot]:
            reasons = {k[6:]: num(r, k) for k in h if k.startswith('stall_') and '(' not in k}
            top = max(reasons, key=reasons.get)
            print(f"   {num(r, '# Samples') / tot * 100:5.2f}%  {top:12s} {r[ix['Source']].strip()[:100]}")


if __name__ == '__main__':
    main()
