import torch, sys
sys.path.insert(0,'/root/repo')
from settlers_of_catan_rl_b200 import gae, normalise_advantages
T,N=200,131072
r=torch.rand(T,N,device='cuda'); v=torch.rand(T+1,N,device='cuda')*300; m=(torch.rand(T+1,N,device='cuda')>0.01).float()
ret,adv=torch.empty_like(r),torch.empty_like(r)
for _ in range(5): gae(r,v,m,0.999,0.95,ret,adv)
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): gae(r,v,m,0.999,0.95,ret,adv)
b.record(); torch.cuda.synchronize()
ms=a.elapsed_time(b)/20
print('gae ms %.4f GB/s %.0f'%(ms, T*N*20/ms/1e6))
