"""Direct calls vs replay of a captured CUDA graph of one step (six launches on two streams): ms per 65 536-env step."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from settlers_of_catan_rl_b200 import VecCatanEnv  # noqa: E402

env = VecCatanEnv(65536, device="cuda:0", seed=0)
env.reset()
acts = env.sample_random()
for _ in range(200):
    env.step_sample(acts)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    env.step_sample(acts)


def timed(fn, n=1000):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("direct %.4f ms/step" % timed(lambda: env.step_sample(acts)))
print("graph  %.4f ms/step" % timed(g.replay))
print("direct %.4f ms/step" % timed(lambda: env.step_sample(acts)))
