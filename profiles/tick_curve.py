"""ms per step against the tick (65 536 games from a fresh reset): where the workload becomes stationary.
Usage: python profiles/tick_curve.py [envs] [ticks] [block]  ->  one line per block of ticks"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
block = int(sys.argv[3]) if len(sys.argv) > 3 else 250
env = VecCatanEnv(n, seed=0)
env.reset()
a = env.sample_random()
for _ in range(3):
    env.step_sample(a)
out = []
done = torch.zeros((), dtype=torch.int64, device=env.device)
for t0 in range(3, ticks, block):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(block):
        env.step_sample(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / block
    out.append({"tick": t0, "ms_per_step": round(ms, 4), "M_steps_per_s": round(n / ms / 1e3, 1)})
    print(json.dumps(out[-1]), flush=True)
