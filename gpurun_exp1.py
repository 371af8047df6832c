import sys, time, torch, numpy as np
sys.path.insert(0,'.')
import bench
from flypylib_b200 import fplmodels, fplnetwork, fplobjdetect, _lib
net = fplnetwork.FplNetwork(fplmodels.vgg_like2); net.train_single.set_weights(bench.seeded_weights('vgg_like2')); net.set_precision('bf16'); net._set_infer()
ctx=_lib.context(0)
dev=torch.device('cuda',0)
size=512
vol=bench.synth_volume_device(size,1,dev); pred=torch.empty((size,)*3,dtype=torch.float32,device=dev)
for i in range(8):
    e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(); t0=time.perf_counter()
    e[0].record()
    net.infer_device(vol, normalize=(128.,33.), out=pred)
    t1=time.perf_counter()
    e[1].record()
    o=fplobjdetect.voxel2obj_device(pred,27,5,(0,0,0),15,0)
    e[2].record()
    torch.cuda.synchronize(); t2=time.perf_counter()
    print('step',i,'host infer-call %.1f ms, total wall %.1f ms; device infer %.1f ms, v2o %.1f ms'%((t1-t0)*1e3,(t2-t0)*1e3,e[0].elapsed_time(e[1]),e[1].elapsed_time(e[2])))
