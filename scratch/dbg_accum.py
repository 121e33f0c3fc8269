import sys; sys.path.insert(0,'.')
import ctypes as C, numpy as np
from hortimapping_b200 import _lib
from tests.gpu_helpers import pepper_decoder
dec=pepper_decoder(); L=_lib.lib()
L.hm_debug_tc_selftest.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p,C.c_void_p,C.c_int,C.c_int,C.c_int]
g=np.random.default_rng(0)
lanes=np.array([32*(r//16)+r%16 for r in range(64)])
for kind in ['normal','positive']:
    A=g.standard_normal((64,64)); B=g.standard_normal((128,64))
    if kind=='positive': A=np.abs(A); B=np.abs(B)
    A=A.astype(np.float16); B=B.astype(np.float16)
    exact1=A.astype(np.float64)@B.astype(np.float64).T
    for rep in [1,2,4,8,24,96]:
        out=np.zeros((128,256),np.float32)
        _lib.check(L.hm_debug_tc_selftest(dec.handle,A.view(np.uint16).ctypes.data,B.view(np.uint16).ctypes.data,out.ctypes.data,0,0,rep),'st')
        got=out[lanes,:128].astype(np.float64); ex=exact1*rep
        scale=(np.abs(A.astype(np.float64))@np.abs(B.astype(np.float64)).T)*rep
        err=(got-ex)
        print(kind,'rep',rep,'MMAs',rep*4,'max |err|/scale %.2e'%(np.abs(err)/scale).max(), 'mean signed err/|ex| %.2e'%np.mean(err/np.abs(ex)*np.sign(ex)), 'rms err/|ex| %.2e'%np.sqrt(np.mean((err/ex)**2)))
