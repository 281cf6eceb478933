#!/usr/bin/env python3
"""Per-CUDA-source-line instruction counts and stall samples from `ncu --page source --csv
--print-source cuda,sass` output (stdin or file).  usage: ncu_lines.py file.csv [top]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hi = [i for i, r in enumerate(rows) if '# Samples' in r][0]
hdr = rows[hi]
si, ei = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
cur, agg = None, {}
for r in rows[hi + 1:]:
    if len(r) <= ei:
        continue
    if r[0] != '':
        cur = (num(r[0]), r[1])
        agg.setdefault(cur, [0, 0, {}])
    elif cur is not None:
        a = agg[cur]
        a[0] += num(r[ei])
        a[1] += num(r[si])
        for i in stall:
            v = num(r[i])
            if v:
                a[2][hdr[i][6:]] = a[2].get(hdr[i][6:], 0) + v
print('total warp instructions', sum(a[0] for a in agg.values()))
nb = sum(a[1] - a[2].get('barrier', 0) for a in agg.values())
print('non-barrier samples', nb)
tot = {}
for a in agg.values():
    for k, v in a[2].items():
        tot[k] = tot.get(k, 0) + v
print('stalls', sorted(tot.items(), key=lambda kv: -kv[1])[:8])
for k in sorted(agg, key=lambda k: -(agg[k][1] - agg[k][2].get('barrier', 0)))[:top]:
    a = agg[k]
    st = {x: y for x, y in a[2].items() if x != 'barrier'}
    print(str(k[0]).rjust(5), str(a[0]).rjust(9), str(a[1] - a[2].get('barrier', 0)).rjust(6),
          k[1].strip()[:88].ljust(88), dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
