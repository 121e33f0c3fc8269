import sys; sys.path.insert(0,'.')
import ctypes as C, numpy as np
from tests.gpu_helpers import pepper_decoder
dec=pepper_decoder(); L=dec._L
L.hm_debug_tc_pair_probe.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p,C.c_void_p,C.c_void_p,C.c_int,C.c_int,C.c_int]
def run(m_rows,N,A,B,reps=1):
    a=np.ascontiguousarray(A,np.float16).view(np.uint16); b=np.ascontiguousarray(B,np.float16).view(np.uint16)
    out=np.zeros((2,128,256),np.float32); cyc=(C.c_longlong*2)()
    rc=L.hm_debug_tc_pair_probe(dec.handle,a.ctypes.data,b.ctypes.data,out.ctypes.data,cyc,m_rows,N,reps)
    assert rc==0, L.hm_last_error()
    return out,(cyc[0],cyc[1])
for m_rows in (64,128):
    N=128
    # rows: A = "identity" (A[r][k]=1 iff k==r%64), B[n][k]=k+1 for all n  -> D[r][n] = (r%64)+1 ; plus row-block id via second run
    A=np.zeros((2,m_rows,64)); 
    for c in range(2):
        for r in range(m_rows): A[c,r,r%64]=1
    B=np.tile(np.arange(1,65)[None,:],(N,1))
    out,_=run(m_rows,N,A,B)
    # columns: A[r][0]=1, B[n][0]=n+1
    A2=np.zeros((2,m_rows,64)); A2[:,:,0]=1
    B2=np.zeros((N,64)); B2[:,0]=np.arange(1,N+1)
    out2,_=run(m_rows,N,A2,B2)
    # row block (which half of m_rows, which CTA): A[c][r][0] = 1 + (r//64) + 2*c, B[n][0]=1
    A3=np.zeros((2,m_rows,64)); 
    for c in range(2):
        for r in range(m_rows): A3[c,r,0]=1+(r//64)+2*c
    B3=np.zeros((N,64)); B3[:,0]=1
    out3,_=run(m_rows,N,A3,B3)
    print(f'=== cta_group::2, M={2*m_rows} ({m_rows} rows per CTA), N={N}')
    for c in range(2):
        o,o2,o3=out[c],out2[c],out3[c]
        used=~np.isnan(o)
        lanes=np.nonzero(used.any(1))[0]; cols=np.nonzero(used.any(0))[0]
        print(f' CTA{c}: lanes used {lanes.min() if len(lanes) else None}..{lanes.max() if len(lanes) else None} ({len(lanes)}), cols used {cols.min() if len(cols) else None}..{cols.max() if len(cols) else None} ({len(cols)})')
        for lane in (0,1,15,16,31,32,47,48,63,64,96,127):
            cc=np.nonzero(used[lane])[0]
            if len(cc)==0: print(f'   lane {lane:3d}: unused'); continue
            print(f'   lane {lane:3d}: row%64={int(o[lane,cc[0]])-1:3d} block={int(o3[lane,cc[0]])} cols[{cc.min()}..{cc.max()}] -> n = {int(o2[lane,cc[0]])-1}..{int(o2[lane,cc[-1]])-1}')
# rate
for m_rows in (64,128):
    for N in (128,256):
        A=np.ones((2,m_rows,64))*0.01; B=np.ones((N,64))*0.01
        run(m_rows,N,A,B,8)
        _,cyc=run(m_rows,N,A,B,2000)
        n=8000
        print(f'rate cta_group::2 M={2*m_rows} N={N}: issue {cyc[0]/n:.1f} cyc/MMA, complete {cyc[1]/n:.1f} cyc/MMA -> {2*m_rows*N*16/(cyc[1]/n)/2:.0f} MAC/clk/SM')
