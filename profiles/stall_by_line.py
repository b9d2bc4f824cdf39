"""Per-source-line stall-sample summary of one kernel from an ncu report (needs -lineinfo and
--import-source on).  usage: python profiles/stall_by_line.py <report.ncu-rep> <kernel regex> [top]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                      '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = None
agg = collections.OrderedDict()
tot = 0
for r in rows:
    if r and r[0] in ('Address', '#') or (H is None and 'Source' in r):
        if H is not None:
            break                       # second kernel instance
        H = r
        continue
    if H is None or len(r) != len(H):
        continue
    try:
        n = int(r[H.index('# Samples')])
    except ValueError:
        continue
    src = r[H.index('Source')].strip()
    tot += n
    agg[src] = agg.get(src, 0) + n
print('total samples', tot, 'columns', H[:4])
for s, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print('%6d %5.1f%%  %s' % (n, 100.0 * n / max(tot, 1), s[:110]))
