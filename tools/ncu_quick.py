"""ncu_quick.py <report> <kernel regex>: key metrics, stall reasons and hottest SASS lines of one captured kernel."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
def run(page):
    cmd = ['ncu', '-i', rep, '--page', page, '--csv', '--kernel-name', 'regex:' + kern]
    return list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
rows = run('raw')
hdr, units, r = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']
for h in want:
    if h in hdr:
        print('%-70s %s %s' % (h, r[hdr.index(h)], units[hdr.index(h)]))
items = [(float(r[i]), h) for i, h in enumerate(hdr) if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
print('stalls:', ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
                           for v, h in sorted(items, reverse=True)[:9]))
src = run('source')
hdr2 = src[1]
ia, isrc, isamp, iex = [hdr2.index(x) for x in ('Address', 'Source', '# Samples', 'Instructions Executed')]
data = [x for x in src[2:] if len(x) > iex and x[isamp].strip().isdigit()]
tot = sum(int(x[isamp]) for x in data)
print('instructions executed (warp):', sum(int(x[iex]) for x in data), 'samples', tot)
for x in sorted(data, key=lambda x: -int(x[isamp]))[:top]:
    print('%5.1f%% %10s  %s' % (100 * int(x[isamp]) / tot, x[iex], x[isrc].strip()[:90]))
