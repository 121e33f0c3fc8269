"""Print the decoder timeline captured by scripts/probe_decoder.py trace (gpurun_out/trace.npy): one unit of the first CTA pair."""
import sys
import numpy as np
tr = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.npy')
unit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [4, 5]
names = {1: 'mma: A ready', 2: 'mma: buffer free', 3: 'mma: group issued', 10: 'epi: partial full', 12: 'epi: promoted', 13: 'epi: fin start', 14: 'epi: fin done'}
ev = []
for r in range(3):
    w, c = tr[r, :, 0].astype(np.int64), tr[r, :, 1].astype(np.int64)
    n = int(np.nonzero(w)[0].max()) + 1 if w.any() else 0
    c = c[:n].copy(); w = w[:n]
    # unwrap the 32-bit clock and shift to the region's origin event (index 0, taken right after the cluster sync)
    d = np.diff(c); d[d < 0] += 1 << 32
    t = np.concatenate([[0], np.cumsum(d)])
    # count units: a new unit starts when op goes back to 0 (epilogue: code 10 op 0 first; mma: code 1 op 0 g 0)
    u = -1
    for i in range(1, n):
        code, op, idx = w[i] >> 24, (w[i] >> 16) & 0xff, w[i] & 0xffff
        if (r == 0 and code == 1 and op == 0 and idx == 0) or (r > 0 and code == 10 and op == 0 and (i < 2 or ((w[i - 1] >> 16) & 0xff) != 0)):
            u += 1
        ev.append((u, int(t[i]), r, int(code), int(op), int(idx)))
sel = sorted(e for e in ev if e[0] == unit and e[4] in ops)
t0 = sel[0][1]
for u, t, r, code, op, idx in sel:
    print('%8d  %s  %-20s op %2d  %s' % (t - t0, ['MMA   ', '  EPI0', '  EPI1'][r], names.get(code, code), op, ('g%d' % idx) if r == 0 else (("h%d" % idx) if code >= 13 else ('seq%d' % idx))))
