"""Steady-state profiling driver: play `--skip` ticks unprofiled, then `--ticks` ticks inside a cudaProfilerStart/Stop
range.  Run under ncu with `--profile-from-start off` so that only those ticks are captured, e.g.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python profiles/profile_step.py --skip 1500 --ticks 20
  ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/full \
      python profiles/profile_step.py --skip 1500 --ticks 1
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from settlers_of_catan_rl_b200 import VecCatanEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--skip", type=int, default=1500)
ap.add_argument("--ticks", type=int, default=20)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
env = VecCatanEnv(a.envs, device="cuda:0", seed=a.seed)
env.reset()
acts = env.sample_random()
for _ in range(a.skip):
    env.step_sample(acts)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.ticks):
    env.step_sample(acts)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(200):
    env.step_sample(acts)
ev1.record()
torch.cuda.synchronize()
st = env.lr_stats()
print("longest-road updates per tick: %.1f, needing a block-wide search: %.1f" % (st[0] / (a.skip + a.ticks + 200), st[1] / (a.skip + a.ticks + 200)))
print("slow jobs: full-search %d of %d; cycles/job avg %.0f max %d; walk steps/job avg %.1f max %d; tasks/job avg %.1f" % (st[2], st[1], st[4] / max(1, st[1]), st[5], st[6] / max(1, st[1]), st[7], st[3] / max(1, st[1])))
print("ms/tick after the profiled range (unprofiled launches): %.4f" % (ev0.elapsed_time(ev1) / 200))
