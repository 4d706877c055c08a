import torch, sys
sys.path.insert(0, '/root/repo')
from settlers_of_catan_rl_b200 import VecCatanEnv
v = VecCatanEnv(300, seed=5)
v.reset()
a = v.sample_random()
for _ in range(int(sys.argv[1])):
    v.step_sample(a)
torch.cuda.synchronize()
print("done", v.lr_stats()[:3], int(v.info[:,0].sum()))
