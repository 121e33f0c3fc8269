"""Per-op summary of the decoder timeline (gpurun_out/trace.npy, scripts/probe_decoder.py trace): for each unit of the first CTA pair,
when the MMA warp began waiting for each op's first A operand -> duration of every op of the chain, in cycles of the instrumented build."""
import sys
import numpy as np
tr = np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/trace.npy')
# default: the MMA issuer's "A ready" events (region 0, code 1); a -DHM_TC_LIGHT build only records the leader CTA's first epilogue
# warp seeing the first partial accumulator of each op (region 1, code 10): `trace_ops.py file 1 10`
REGION, CODE = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1)
w, c = tr[REGION, :, 0].astype(np.int64), tr[REGION, :, 1].astype(np.int64)
n = int(np.nonzero(w)[0].max()) + 1
d = np.diff(c[:n]); d[d < 0] += 1 << 32
t = np.concatenate([[0], np.cumsum(d)])
starts = []            # (time, op) of the first "A ready" wait of each op occurrence
prev_op = -1
for i in range(0, n):
    code, op = w[i] >> 24, (w[i] >> 16) & 0xff
    if code == CODE and op != prev_op:
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
