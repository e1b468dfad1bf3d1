import sys, json
sys.path.insert(0,'cellranger-atac_b200'); sys.path.insert(0,'.')
import torch, numpy as np
from cratac import _ffi, synth
import bench
frs=int(sys.argv[1]) if len(sys.argv)>1 else 250000000
ctx=_ffi.Context(synth.GRCH38, device=0)
dev=torch.device('cuda',0)
text,nb,_n=bench.make_workload(torch, ctx, synth.GRCH38, frs, 10000, 100000, 2, dev)
ctx.decode_device(text.data_ptr(), nb)
v,w=ctx.count_cut_sites()
print('bins',len(v),'vmax',int(v.max()),'sum',int(w.sum()))
np.savez('gpurun_out/hist_%d.npz'%frs, v=v, w=w)
for mode in (2,1):
    ctx.set_option('em_fast',mode)
    import time; t=time.time(); r=ctx.fit_threshold(v,w,0.2); dt=time.time()-t
    print(mode, r, 'ms',dt*1e3, ctx.em_debug())
