"""time the env step of a build variant: CATAN_B200_LIB=<so> python profiles/variant_bench.py [envs] [skip] [ticks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from settlers_of_catan_rl_b200 import VecCatanEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 500
env = VecCatanEnv(n, seed=0); env.reset(); a = env.sample_random()
for _ in range(skip): env.step_sample(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(ticks): env.step_sample(a)
e1.record(); torch.cuda.synchronize()
env.set_timing(True)
for _ in range(200): env.step_sample(a)
_, tr, en, rows = env.read_timing(detail=True)
print("%s: %.4f ms/step  transition %.4f  encode %.4f (rows %.4f, masks %.4f)  errs %d" % (os.environ.get("CATAN_B200_LIB", "default"), e0.elapsed_time(e1) / ticks, tr, en, rows, en - rows, int(env.err_flags().any())))
print("   streams (ms from the end of the transition):", {k: round(v, 4) for k, v in env.stream_timing.items()})
h = env.lr_histograms()
for name, row in zip(("cycles per search", "walk steps per search", "cycles of a step's longest search"), h):
    print("   log2 histogram,", name + ":", {int(b): int(c) for b, c in enumerate(row) if c})
