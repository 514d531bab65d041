"""ncu_stalls.py <report> <kernel regex> [top]: per-instruction stall-reason samples (source page)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern]
src = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
hdr = src[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
seen, data = set(), []
for x in src[2:]:
    if len(x) > iex and x[isamp].strip().isdigit() and x[0] not in seen:
        seen.add(x[0]); data.append(x)
tot = sum(int(x[isamp]) for x in data)
agg = {r: sum(int(x[hdr.index(r)] or 0) for x in data) for r in reasons}
print('total samples', tot, ' executed', sum(int(x[iex]) for x in data))
print('by reason:', ', '.join('%s %.1f%%' % (r[6:], 100 * v / tot) for r, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for x in sorted(data, key=lambda x: -int(x[isamp]))[:top]:
    rs = sorted(((int(x[hdr.index(r)] or 0), r[6:]) for r in reasons), reverse=True)[:2]
    print('%5.1f%% %9s  %-60s %s' % (100 * int(x[isamp]) / tot, x[iex], x[isrc].strip()[:60], ' '.join('%s:%d' % (n, v) for v, n in rs if v)))
