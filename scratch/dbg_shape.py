import sys, copy; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import hm_oracle as O
from tests.helpers import *
from tests.test_gpu_optimizer import make_opt, zero_eps, last_system, rel
c=load_npz('fruit_wild')
for engine in ['simt','tc']:
    for k in [1,2,3,5,6,7,8,12,30]:
        cfg=zero_eps(cfg_of(c),k)
        opt,dec=make_opt(cfg,engine)
        lat=torch.from_numpy(c['init_latent'].copy()).cuda().reshape(1,32); T=torch.from_numpy(c['init_T_ow'].copy()).cuda().reshape(1,4,4)
        _,_,it,st=opt.shape_opt_deepsdf_batch(lat,T,[c['points_w']])
        H,b,dx=last_system(dec,1,32)
        l64=c['init_latent'].astype(np.float64); tr=O.OptTrace()
        O.shape_opt_deepsdf(oracle_decoder(np.float64),cfg,l64,c['init_T_ow'].astype(np.float64),c['points_w'],trace=tr)
        print(engine,k,'iters',it.item(),'status',st.item(),'lat rel',rel(lat.cpu().numpy()[0],l64),'H rel',rel(H[0],tr.H[-1]),'b rel',rel(b[0],tr.b[-1]), 'dx rel', rel(dx[0],tr.dx[-1]))
    dec.set_engine('tc')
