import sys, os; sys.path.insert(0,'.')
import ctypes as C, numpy as np, torch
from tests.gpu_helpers import pepper_decoder, random_rows
from tests.helpers import pepper_weights
dec=pepper_decoder(); L=dec._L
L.hm_debug_tc_wait_cycles.argtypes=[C.c_void_p, C.c_void_p]
_,_,codes=pepper_weights()
n=131072
g=np.random.default_rng(0)
rows=np.concatenate([codes[g.integers(0,919,n)], ((g.random((n,3))*2-1)*0.05).astype(np.float32)],1)
t=torch.from_numpy(rows).cuda()
out=(C.c_ulonglong*13)()
for jac in [True, False]:
    dec._eval_rows(t, with_jac=jac); torch.cuda.synchronize()
    L.hm_debug_tc_wait_cycles(dec.handle, out)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); dec._eval_rows(t, with_jac=jac); e1.record(); torch.cuda.synchronize()
    L.hm_debug_tc_wait_cycles(dec.handle, out)
    v=[int(x) for x in out]; tot=v[4]
    if tot == 0:
        print('jac',jac,'ms',e0.elapsed_time(e1),'(library built without -DHM_TC_COUNTERS)'); continue
    nlead = 148 if os.environ.get('HM_TC_PAIR')=='0' else 74
    for nm,o in (('leader',5),('peer',9)):
        if v[o+3]: print('   epilogue[%s]: total %.0f  wait_full %.1f%%  promote %.1f%%  finalize %.1f%%'%(nm, v[o+3]/nlead, 100*v[o]/v[o+3], 100*v[o+1]/v[o+3], 100*v[o+2]/v[o+3]))
    print('jac',jac,'ms',e0.elapsed_time(e1),'per-CTA avg cycles: total %.0f  wait_A %.1f%%  wait_part %.1f%%  wait_W %.1f%%  (producer wait_empty %.1f%%)'%(tot/nlead, 100*v[1]/tot,100*v[2]/tot,100*v[3]/tot,100*v[0]/tot))
