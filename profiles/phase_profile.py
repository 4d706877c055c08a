"""Per-phase cycle breakdown of the env kernel at steady state, using the instrumented build
(-DCATAN_PROFILE_PHASES) of the same sources.  Run on the GPU box:

    python profiles/phase_profile.py [n_envs] [warm_ticks] [ticks]  > profiles/phase_profile_rNN.txt
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from settlers_of_catan_rl_b200.build import build_profiling_extension, PROF_SO  # noqa: E402

if not os.path.exists(PROF_SO):
    build_profiling_extension()
os.environ["CATAN_B200_LIB"] = PROF_SO
import torch  # noqa: E402
from settlers_of_catan_rl_b200 import VecCatanEnv, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 200
names = ["load", "scalar", "dice", "est", "longest_road", "finish", "masks", "obs", "sample", "store"]
v = VecCatanEnv(n, seed=0)
v.reset()
a = v.sample_random()
for _ in range(warm):
    v.step_sample(a)
lib = _lib.load()
buf = (C.c_ulonglong * (len(names) * 4))()
lib.catan_prof_read(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(ticks):
    v.step_sample(a)
e1.record()
torch.cuda.synchronize()
lib.catan_prof_read(buf, 0)
ms = e0.elapsed_time(e1) / ticks
print("n_envs %d  ticks %d after %d warm  (instrumented) %.3f ms/tick" % (n, ticks, warm, ms))
tot = sum(buf[i * 4] for i in range(len(names)))
# phase-major kernel: one mark per warp per batch and phase (the time includes waiting at the block barrier)
print("%-14s %14s %12s %12s %8s" % ("phase", "marks/tick", "avg cyc", "max cyc", "share"))
for i, nm in enumerate(names):
    s, mx, cnt = buf[i * 4], buf[i * 4 + 1], buf[i * 4 + 2]
    print("%-14s %14.0f %12.0f %12d %7.1f%%" % (nm, cnt / ticks, s / max(cnt, 1), mx, 100.0 * s / tot))
print("warp-cycles per tick per warp (sum of phases): %.0f" % (tot / ticks / (148 * 32)))
