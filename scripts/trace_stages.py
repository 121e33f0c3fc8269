"""Stage-level view of the MMA issuer's timeline (gpurun_out/trace*.npy, scripts/probe_decoder.py trace): for one unit of the first
CTA pair, per weight stage: cycles waiting for the stage (W_FULL), cycles issuing its MMAs, gap to the next stage's wait."""
import sys
import numpy as np
tr = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.npy')
unit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w, c = tr[0, :, 0].astype(np.int64), tr[0, :, 1].astype(np.int64)
n = int(np.nonzero(w)[0].max()) + 1
d = np.diff(c[:n]); d[d < 0] += 1 << 32
t = np.concatenate([[0], np.cumsum(d)])
u = -1
ev = []
for i in range(1, n):
    code, op, idx = int(w[i] >> 24), int((w[i] >> 16) & 0xff), int(w[i] & 0xffff)
    if code == 1 and op == 0 and idx == 0:
        u += 1
    if u == unit:
        ev.append((int(t[i]), code, op, idx))
t0 = ev[0][0]
names = {1: 'group: A ready', 2: 'group: buffer free', 3: 'group: issued', 4: 'stage: wait W', 5: 'stage: W full', 6: 'stage: issued'}
prev = t0
tot = {}
for tm, code, op, idx in ev:
    print('%8d (+%5d) op %2d  %-20s %d' % (tm - t0, tm - prev, op, names.get(code, str(code)), idx))
    tot[code] = tot.get(code, 0) + tm - prev
    prev = tm
print('cycles spent BEFORE each event kind:', {names[k]: v for k, v in sorted(tot.items())})
