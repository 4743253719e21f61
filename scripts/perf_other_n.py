import sys; sys.path.insert(0,"/root/repo/scripts"); sys.path.insert(0,"/root/repo")
import torch
from quick_perf import time_cuda
from pymuscle_b200 import MuscleBatch, MuscleSpec
dev=torch.device("cuda:0")
for spec,M,scale in ((MuscleSpec.standard(90.0),24576,1.0),(MuscleSpec.potvin(64),131072,67.0),(MuscleSpec.potvin(480),16384,67.0),(MuscleSpec.potvin(16),524288,67.0)):
    K=500
    b=MuscleBatch(spec,M,dev,"f32",fresh_outputs=False)
    e=(scale*torch.rand(K,M,device=dev)); out=torch.empty(K,M,device=dev)
    best,_=time_cuda(lambda: b.step_many(e,0.02,out=out),iters=3,warm=1)
    print(f"{spec.name} n={spec.n} x {M}: {M*spec.n*K/best/1e9:.0f} G unit-steps/s {b.fused_geometry(K)}", flush=True)
