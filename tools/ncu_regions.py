"""Summarise an ncu report: headline metrics + per-instruction hot spots (needs -lineinfo / --import-source)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
def run(page):
    cmd = ['ncu', '-i', rep, '--page', page, '--csv']
    if kern: cmd += ['--kernel-name', 'regex:' + kern]
    return list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
rows = run('raw'); hdr = rows[0]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:60])
    for h in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.max','smsp__warps_eligible.avg.per_cycle_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','lts__t_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum']:
        if h in hdr: print('  %-70s %s' % (h, r[hdr.index(h)]))
    items = [(float(r[i]), h) for i, h in enumerate(hdr) if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    print('  stalls:', ', '.join('%s=%.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in sorted(items, reverse=True)[:8]))
rows = run('source'); hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
iw, iid = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
data = [r for r in rows[2:] if len(r) > iw and r[isamp].strip().isdigit()]
base = int(data[0][ia], 16)
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print('total samples', tot, 'instructions', totex, 'shared wavefronts', sum(int(r[iw]) for r in data), 'ideal', sum(int(r[iid]) for r in data))
# split into regions at BAR.SYNC / backward branches
bounds = [0]
for r in data:
    s = r[isrc]
    if 'BAR.SYNC' in s or 'SYNCS' in s or 'UBLKCP' in s: bounds.append(int(r[ia], 16) - base)
bounds.append(1 << 30)
bounds = sorted(set(bounds))
for lo, hi in zip(bounds[:-1], bounds[1:]):
    sel = [r for r in data if lo <= int(r[ia], 16) - base < hi]
    s = sum(int(r[isamp]) for r in sel); e = sum(int(r[iex]) for r in sel); w = sum(int(r[iw]) for r in sel)
    if s or e: print('  region %5x-%5x  samples %5.1f%%  inst %5.1f%% (%d)  wavefronts %d' % (lo, min(hi, 0xfffff), 100 * s / tot, 100 * e / totex, e, w))
print('top stall instructions:')
for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
    print('  %5d %9s %5x %s' % (int(r[isamp]), r[iex], int(r[ia], 16) - base, r[isrc].strip()[:64]))
