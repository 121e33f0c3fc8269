"""Tiny dependency simulator of the decoder's MMA / epilogue hand-shake (design aid, not product code).
Groups of an op: (step, nh).  MMA group k needs: A step ready, TMEM buffer (k & 1) freed by promote(k - 2).
Epilogue: a fixed sequence of P(k) (needs group k executed) and F(nh, jj) (publishes A step 2*nh+jj of the NEXT op)."""
import sys
G, P, F, LAT = 2048, 700, 2000, 250
NOPS = 12

def run(order, epi_seq_fn, verbose=False):
    # order: list of (step, nh) per op; epi_seq_fn(op) -> list of actions ('P', k) / ('F', op_x, nh, jj)
    a_ready = {(0, s): 0 for s in range(4)}          # (op, step) -> time
    exec_end, promoted = {}, {}
    t_mma, t_epi = 0, 0
    acts = []
    for op in range(NOPS):
        acts += [(op, a) for a in epi_seq_fn(op)]
    ai = 0
    gl = [(op, k) for op in range(NOPS) for k in range(8)]
    gi = 0
    progress = True
    while gi < len(gl) or ai < len(acts):
        progress = False
        if gi < len(gl):
            op, k = gl[gi]
            step, nh = order[k]
            need = [a_ready.get((op, step))]
            if gi >= 2:
                need.append(promoted.get(gl[gi - 2]))
            if all(n is not None for n in need):
                start = max([t_mma] + need)
                exec_end[(op, k)] = start + G
                t_mma = start + G
                gi += 1
                progress = True
        if ai < len(acts):
            op, a = acts[ai]
            if a[0] == 'P':
                e = exec_end.get((op, a[1]))
                if e is not None:
                    t_epi = max(t_epi, e + LAT) + P
                    promoted[(op, a[1])] = t_epi + LAT
                    ai += 1
                    progress = True
            else:
                _, opx, nh, jj = a
                t_epi += F
                a_ready[(opx + 1, 2 * nh + jj)] = t_epi + LAT
                ai += 1
                progress = True
        if not progress:
            raise RuntimeError('deadlock at group %s action %s' % (gl[gi] if gi < len(gl) else None, acts[ai] if ai < len(acts) else None))
    return t_mma / (NOPS * 8 * G)

ORD_A = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (3, 0), (2, 1), (3, 1)]
ORD_B = [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (3, 0), (2, 1), (3, 1)]

def seq_old(op):      # before: P x6, F0 (2 quarters), P, P, F1 (2 quarters)
    return [('P', k) for k in range(6)] + [('F', op, 0, 0), ('F', op, 0, 1), ('P', 6), ('P', 7), ('F', op, 1, 0), ('F', op, 1, 1)]

def seq_s2(op):       # current: P0 [F11 prev] P1..P5 F00 P6 F01 P7 F10
    s = [('P', 0)] + ([('F', op - 1, 1, 1)] if op else []) + [('P', k) for k in range(1, 6)]
    return s + [('F', op, 0, 0), ('P', 6), ('F', op, 0, 1), ('P', 7), ('F', op, 1, 0)]

def seq_s3(op):       # proposed (with ORD_B): P0 [F10 prev] P1 [F11 prev] P2..P5 F00 P6 F01 P7
    s = [('P', 0)] + ([('F', op - 1, 1, 0)] if op else []) + [('P', 1)] + ([('F', op - 1, 1, 1)] if op else []) + [('P', k) for k in range(2, 6)]
    return s + [('F', op, 0, 0), ('P', 6), ('F', op, 0, 1), ('P', 7)]

for name, order, fn in (('old', ORD_A, seq_old), ('S2 (current)', ORD_A, seq_s2), ('S3', ORD_B, seq_s3)):
    for (p, f) in ((700, 2000), (700, 2400), (500, 1500), (350, 1200)):
        P, F = p, f
        print('%-14s P=%4d F=%4d  time / MMA floor = %.3f' % (name, p, f, run(order, fn)))

# ---- 16 epilogue warps: finalize works on whole output halves (publishes both k-steps of the half at its end)
def seq16_nodefer(op):
    return [('P', k) for k in range(6)] + [('F2', op, 0), ('P', 6), ('P', 7), ('F2', op, 1)]

def seq16_defer(op):
    return [('P', 0)] + ([('F2', op - 1, 1)] if op else []) + [('P', k) for k in range(1, 6)] + [('F2', op, 0), ('P', 6), ('P', 7)]

def run16(order, fn, f0, f1):
    global F
    def expand(op):
        out = []
        for a in fn(op):
            if a[0] == 'F2':
                out.append(('F', a[1], a[2], 0, f1 if a[2] else f0))
            else:
                out.append(a)
        return out
    # re-implement with per-action finalize time; a half publishes steps 2nh and 2nh+1 together
    a_ready = {(0, s): 0 for s in range(4)}
    exec_end, promoted = {}, {}
    t_mma = t_epi = 0
    acts = [(op, a) for op in range(NOPS) for a in expand(op)]
    gl = [(op, k) for op in range(NOPS) for k in range(8)]
    gi = ai = 0
    while gi < len(gl) or ai < len(acts):
        progress = False
        if gi < len(gl):
            op, k = gl[gi]
            step, nh = order[k]
            need = [a_ready.get((op, step))] + ([promoted.get(gl[gi - 2])] if gi >= 2 else [])
            if all(n is not None for n in need):
                start = max([t_mma] + need)
                exec_end[(op, k)] = t_mma = start + G
                gi += 1
                progress = True
        if ai < len(acts):
            op, a = acts[ai]
            if a[0] == 'P':
                e = exec_end.get((op, a[1]))
                if e is not None:
                    t_epi = max(t_epi, e + LAT) + P
                    promoted[(op, a[1])] = t_epi + LAT
                    ai += 1
                    progress = True
            else:
                _, opx, nh, _, dur = a
                t_epi += dur
                a_ready[(opx + 1, 2 * nh)] = a_ready[(opx + 1, 2 * nh + 1)] = t_epi + LAT
                ai += 1
                progress = True
        assert progress
    return t_mma / (NOPS * 8 * G)

print()
for p in (600,):
    P = p
    for f0, f1 in ((2400, 3800), (2400, 2400), (1850, 2900), (1850, 1850), (1500, 1500)):
        print('16 warps P=%d F0=%d F1=%d: no-defer %.3f   defer %.3f' % (p, f0, f1, run16(ORD_A, seq16_nodefer, f0, f1), run16(ORD_A, seq16_defer, f0, f1)))
