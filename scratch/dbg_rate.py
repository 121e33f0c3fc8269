import sys; sys.path.insert(0,'.')
import ctypes as C
from tests.gpu_helpers import pepper_decoder
dec=pepper_decoder(); L=dec._L
L.hm_debug_tc_mma_rate.argtypes=[C.c_void_p,C.c_int,C.c_int,C.c_int,C.c_int,C.c_void_p]
out=(C.c_longlong*2)()
for M in (64,128):
    for N in (64,128,256):
        for nacc in (1,2):
            if N*nacc>512: continue
            L.hm_debug_tc_mma_rate(dec.handle,M,N,8,nacc,out)
            reps=2000
            L.hm_debug_tc_mma_rate(dec.handle,M,N,reps,nacc,out)
            n=reps*4
            print(f'M={M} N={N} acc_bufs={nacc}: issue {out[0]/n:.1f} cyc/MMA, complete {out[1]/n:.1f} cyc/MMA -> {M*N*16/(out[1]/n):.0f} MAC/clk')
