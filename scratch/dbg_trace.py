import sys, os; sys.path.insert(0,'.')
import ctypes as C, numpy as np, torch
from tests.gpu_helpers import pepper_decoder
from tests.helpers import pepper_weights
dec=pepper_decoder(); L=dec._L
L.hm_debug_tc_trace.argtypes=[C.c_void_p, C.c_int, C.c_void_p]
_,_,codes=pepper_weights()
n=131072
g=np.random.default_rng(0)
rows=np.concatenate([codes[g.integers(0,919,n)], ((g.random((n,3))*2-1)*0.05).astype(np.float32)],1)
t=torch.from_numpy(rows).cuda()
dec._eval_rows(t, with_jac=True); torch.cuda.synchronize()
L.hm_debug_tc_trace(dec.handle, 1, None)
dec._eval_rows(t, with_jac=True); torch.cuda.synchronize()
out=np.zeros(3*8192*2, np.uint32)
L.hm_debug_tc_trace(dec.handle, 0, out.ctypes.data)
np.save('gpurun_out/trace.npy', out.reshape(3,8192,2))
print('trace saved', [(int((out.reshape(3,8192,2)[r,:,0]!=0).sum())) for r in range(3)])
