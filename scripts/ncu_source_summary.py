"""Summarise an `ncu --page source --csv` export: stall mix, opcode mix, hottest instructions."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        break
    data.append(r)
print(len(data), 'instructions')
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:12]:
    print('  %-24s %8d %5.1f%%' % (s, v, 100 * v / tot))
c = Counter()
for r in data:
    op = [o for o in r[idx['Source']].split() if not o.startswith('@')][0]
    c[op.split('.')[0]] += int(r[idx['Instructions Executed']])
ti = sum(c.values())
print('total warp instructions', ti)
for k, v in c.most_common(28):
    print('  %-10s %10d %5.1f%%' % (k, v, 100 * v / ti))
print('hottest instructions by samples:')
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = max(stalls, key=lambda s: int(r[idx[s]]))
    print('  %6s %-60s %s' % (r[idx['# Samples']], r[idx['Source']].strip()[:60], top))
