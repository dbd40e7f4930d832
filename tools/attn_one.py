import os, sys, torch
sys.path.insert(0, '/root/repo')
from icka_b200 import ops
DEV='cuda:0'
B,Sq,Skv,nh=512,256,196,12
H=nh*64
g=torch.Generator(DEV).manual_seed(1)
q=torch.randn(B*Sq,H,device=DEV,generator=g).bfloat16()
kv=torch.randn(B*Skv,2*H,device=DEV,generator=g).bfloat16()
mask=torch.zeros(B,Skv,device=DEV)
for _ in range(4):
    ops.cross_attn_core(q,kv[:,:H],kv[:,H:],mask,B,Sq,Skv,nh,64)
torch.cuda.synchronize()
