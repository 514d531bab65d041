"""ncu --csv launch/metric list (tools/prof_all.py) -> markdown table, one row per kernel (averages over its launches)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path, errors='replace')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]
ix = {n: h.index(n) for n in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'Grid Size', 'Block Size')}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    d = launch.setdefault(r[ix['ID']], {'name': r[ix['Kernel Name']], 'grid': r[ix['Grid Size']], 'block': r[ix['Block Size']]})
    try:
        v = float(r[ix['Metric Value']].replace(',', ''))
    except ValueError:
        continue
    unit = r[ix['Metric Unit']]
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 1.0)
    d[r[ix['Metric Name']]] = v * scale


def short(n):
    n = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', n)
    n = re.sub(r'\(.*$', '', n)
    return n.strip()[:58]


# one section per plan (tools/prof_all.py plans configurations 1, 2/4, 3, 5 in this order; k_bin_keys opens a plan)
SECTIONS = ['configuration 1 (2-D 256^2 / 512^2, 1 coil, PROPELLER M = 122 880)',
            'configurations 2 and 4 (2-D 256^2 / 512^2, 32 coils, radial M = 205 824; CG and L1TVOLS iterations)',
            'configuration 3 (3-D 128^3 / 256^3, 1 coil, M = 2 000 000)',
            'configuration 5, one GPU of eight (3-D 128^3 / 256^3, 4 coils, M = 2 000 000; k-space CG iterations)']
secs = []
for d in launch.values():
    if short(d['name']) == 'k_bin_keys':
        secs.append(collections.OrderedDict())
    if secs:
        secs[-1].setdefault(short(d['name']) + ' grid ' + d['grid'], []).append(d)
M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
     'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
def table(agg):
    print('| kernel (launch grid) | launches | avg us | DRAM read MB | DRAM write MB | DRAM GB/s | L2 hit % | issue active % | warps active % | warp inst (M) | smem wavefronts (M) |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for name, ds in agg.items():
        av = lambda m: sum(d.get(m, 0.0) for d in ds) / len(ds)
        t, rd, wr = av(M[0]), av(M[1]), av(M[2])
        print('| `%s` | %d | %.1f | %.1f | %.1f | %.0f | %.0f | %.0f | %.0f | %.2f | %.2f |' % (
            name, len(ds), t, rd, wr, (rd + wr) / t * 1e3 if t > 0 else 0, av(M[3]), av(M[4]), av(M[5]), av(M[6]) / 1e6, av(M[7]) / 1e6))


for i, agg in enumerate(secs):
    print('\n### %s\n' % (SECTIONS[i] if i < len(SECTIONS) else 'plan %d' % i))
    table(agg)
