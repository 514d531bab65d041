"""Turn gpurun_out/{launches,prof}_<tag> into profiles/<tag>_*.md / traffic.json (run in the build container)."""
import csv, io, json, subprocess, sys, collections, os
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
# ---- launch list
rows = list(csv.reader(open(os.path.join(root, 'gpurun_out', 'launches_%s.csv' % tag))))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i
        break
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
seq = [(r[ik], float(r[iv])) for r in rows[start + 2:] if len(r) > iv]
# last 'pair' = the kernels of the final timed step: take the last occurrence window between two interp launches
names = [n for n, _ in seq]
idx = [i for i, n in enumerate(names) if 'k_fft256<1>' in n] or [i for i, n in enumerate(names) if 'k_interp_tiled' in n]
# a device-resident step is 11 launches apart (3 FFT passes, interp, memset, gather, counters, scatter, 3 inverse
# passes); the e2e / kernel-timing legs of bench.py launch other sequences: keep the last window of that length
pairs = [(a, b) for a, b in zip(idx[:-1], idx[1:]) if any('k_gridding' in n for n in names[a:b]) and any('k_interp' in n for n in names[a:b])]
if pairs:
    idx = [pairs[-1][0], pairs[-1][1]]
agg = collections.OrderedDict()
if len(idx) >= 2:
    # find a window that holds exactly one forward+adjoint step (from one interp to the next, shifted to start at the FFT passes)
    w = seq[idx[-2]:idx[-1]]
    tot = sum(t for _, t in w)
    for n, t in w:
        key = n.split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:70]
        agg[key] = agg.get(key, 0) + t
    out.append('## Launch list of one timed step (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py)\n')
    out.append('Cold-cache, serialised per-launch times: shares, not absolutes.\n')
    out.append('| kernel | ns | share |\n|---|---|---|')
    for k, t in agg.items():
        out.append('| `%s` | %.0f | %.1f %% |' % (k, t, 100 * t / tot))
    out.append('| total | %.0f | |\n' % tot)
# ---- full captures
def run(page, kern):
    cmd = ['ncu', '-i', os.path.join(root, 'gpurun_out', 'prof_%s.ncu-rep' % tag), '--page', page, '--csv', '--kernel-name', 'regex:' + kern]
    return list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
traffic = {}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for kern, short in (('k_interp_tiled', 'interp'), ('k_gridding_col', 'gridding'), ('k_gridding_tiled', 'gridding_tiled')):
    rows = run('raw', kern)
    if len(rows) < 3:
        continue
    hdr, units, r = rows[0], rows[1], rows[2]
    out.append('## `%s` (`ncu --set full --clock-control none --import-source on`, one launch, configuration 3)\n' % kern)
    out.append('| metric | value | unit |\n|---|---|---|')
    vals = {}
    for h in want:
        if h in hdr:
            vals[h] = r[hdr.index(h)]
            out.append('| %s | %s | %s |' % (h, r[hdr.index(h)], units[hdr.index(h)]))
    items = [(float(r[i]), h) for i, h in enumerate(hdr) if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    out.append('\nTop stall reasons (warps stalled per issue-active cycle): ' + ', '.join(
        '%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in sorted(items, reverse=True)[:7]) + '\n')
    def num(x):
        return float(x.replace(',', ''))
    def to_bytes(h):
        u = units[hdr.index(h)].lower()
        return num(vals[h]) * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]
    traffic[short] = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
    out.append('DRAM traffic per launch: %.1f MB (algorithmic bytes 582.2 MB).\n' % (traffic[short] / 1e6))
    src = run('source', kern)
    hdr2 = src[1]
    ia, isrc, isamp, iex, iw, iid = [hdr2.index(x) for x in ('Address', 'Source', '# Samples', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal')]
    seen, data = set(), []
    for x in src[2:]:                                   # first captured launch only (later launches repeat the header)
        if len(x) > iw and x[isamp].strip().isdigit() and x[ia] not in seen:
            seen.add(x[ia])
            data.append(x)
    tot = sum(int(x[isamp]) for x in data)
    out.append('Shared-memory wavefronts: %d (ideal %d); instructions %d; hottest instructions by stall samples:\n' % (
        sum(int(x[iw]) for x in data), sum(int(x[iid]) for x in data), sum(int(x[iex]) for x in data)))
    out.append('| samples % | executed | SASS |\n|---|---|---|')
    for x in sorted(data, key=lambda x: -int(x[isamp]))[:10]:
        out.append('| %.1f | %s | `%s` |' % (100 * int(x[isamp]) / tot, x[iex], x[isrc].strip()[:70]))
    out.append('')
# the gridding STAGE bench.py times = pre-pass (zero-fill + sorted-data gather) + scatter: add the pre-pass traffic
rows = run('raw', 'k_gather_sorted_col')
if len(rows) >= 3 and 'gridding' in traffic:
    hdr, units, r = rows[0], rows[1], rows[2]
    def _b(h):
        u = units[hdr.index(h)].lower()
        return float(r[hdr.index(h)].replace(',', '')) * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[u]
    pre = _b('dram__bytes_read.sum') + _b('dram__bytes_write.sum')
    traffic['gridding_scatter_kernel'] = traffic['gridding']
    traffic['gridding_prepass_kernel'] = pre
    traffic['gridding'] = traffic['gridding'] + pre
    out.append('Gridding stage as timed by bench.py (pre-pass + scatter): DRAM traffic %.1f MB (pre-pass %.1f MB).\n' % (traffic['gridding'] / 1e6, pre / 1e6))
os.makedirs(os.path.join(root, 'profiles'), exist_ok=True)
open(os.path.join(root, 'profiles', '%s_ncu_summary.md' % tag), 'w').write('# ncu summary %s (B200, configuration 3)\n\n' % tag + '\n'.join(out) + '\n')
json.dump(traffic, open(os.path.join(root, 'profiles', 'traffic.json'), 'w'))
print('\n'.join(out)[:3000])
