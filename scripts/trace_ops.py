"""Per-op summary of the decoder timeline (gpurun_out/trace.npy, scripts/probe_decoder.py trace): for each unit of the first CTA pair,
when the MMA warp began waiting for each op's first A operand -> duration of every op of the chain, in cycles of the instrumented build."""
import sys
import numpy as np
tr = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.npy')
w, c = tr[0, :, 0].astype(np.int64), tr[0, :, 1].astype(np.int64)
n = int(np.nonzero(w)[0].max()) + 1
d = np.diff(c[:n]); d[d < 0] += 1 << 32
t = np.concatenate([[0], np.cumsum(d)])
starts = []            # (time, op) of the first "A ready" wait of each op occurrence
prev_op = -1
for i in range(1, n):
    code, op = w[i] >> 24, (w[i] >> 16) & 0xff
    if code == 1 and op != prev_op:
        starts.append((int(t[i]), int(op)))
        prev_op = op
units, cur = [], []
for tm, op in starts:
    if cur and op <= cur[-1][1]:
        units.append(cur); cur = []
    cur.append((tm, op))
units.append(cur)
for u in units[1:6]:
    ops = [o for _, o in u]
    dur = [u[i + 1][0] - u[i][0] for i in range(len(u) - 1)]
    print("ops", ops, "durations", dur, "unit total (to the last op's start)", u[-1][0] - u[0][0])
if len(units) > 3:
    per = [units[i + 1][0][0] - units[i][0][0] for i in range(1, len(units) - 1)]
    print("tile period (start of op", units[1][0][1], "to the next tile's):", per[:8])
