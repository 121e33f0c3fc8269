import sys; sys.path.insert(0,'.')
import ctypes as C
from tests.gpu_helpers import pepper_decoder
dec=pepper_decoder(); L=dec._L
L.hm_debug_tc_ingest.argtypes=[C.c_void_p,C.c_int,C.c_int,C.c_int,C.POINTER(C.c_double),C.POINTER(C.c_double)]
bpc=C.c_double(); ms=C.c_double()
for cl in (1,2,4,8):
    for bytes_ in (16384,32768):
        n=4000
        L.hm_debug_tc_ingest(dec.handle,cl,n,bytes_,C.byref(bpc),C.byref(ms))
        grid=(148//cl)*cl
        print(f'cluster {cl} stage {bytes_}: {bpc.value:.1f} B/clk/SM ingest, {ms.value:.3f} ms, SM-ingest total {grid*n*bytes_/ms.value/1e9:.2f} TB/s, L2 reads {grid*n*bytes_/cl/ms.value/1e9:.2f} TB/s')
