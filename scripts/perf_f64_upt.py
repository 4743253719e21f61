import sys; sys.path.insert(0,"/root/repo/scripts"); sys.path.insert(0,"/root/repo")
import torch
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev=torch.device("cuda:0")
for kind,spec,scale in (("potvin",MuscleSpec.potvin(120),67.0),("standard",MuscleSpec.standard(),1.0)):
  for M in (16384, 32768):
    for upt in (2,4):
        K=1000
        b=MuscleBatch(spec,M,dev,"f64",fresh_outputs=False)
        exc=(scale*torch.rand(K,M,device=dev)).double(); out=torch.empty(K,M,device=dev,dtype=torch.float64)
        best,mean=time_cuda(lambda: b.step_many(exc,0.02,out=out,units_per_thread=upt),iters=3,warm=1)
        print(kind,M,"U=",upt,b.fused_geometry(K,upt), round(M*120*K/best/1e9,1),"G unit-steps/s", flush=True)
