"""Per-phase cycle markers of the two step kernels (instrumented twin build, -DCATAN_PROFILE_PHASES): average cycles from block
start to each marker at steady state.  usage: python profiles/phase_profile.py [envs] [skip] [ticks]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from settlers_of_catan_rl_b200 import build
so = build.PROF_SO if os.path.exists(build.PROF_SO) else build.build_profiling_extension()
os.environ["CATAN_B200_LIB"] = so
import numpy as np, torch
from settlers_of_catan_rl_b200 import VecCatanEnv, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 200
env = VecCatanEnv(n, seed=0); env.reset(); a = env.sample_random()
for _ in range(skip): env.step_sample(a)
torch.cuda.synchronize()
lib = _lib.load(); out = np.zeros(64, np.uint64)
lib.catan_debug_read_phases.argtypes = [C.c_void_p, C.c_int]
lib.catan_debug_read_phases(C.c_void_p(out.ctypes.data), 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(ticks): env.step_sample(a)
e1.record(); torch.cuda.synchronize()
lib.catan_debug_read_phases(C.c_void_p(out.ctypes.data), 0)
names = {0: "transition: chunk staged", 1: "transition: rule warp after the scalar rules", 2: "transition: rule warp after longest road",
         3: "transition: follow-up warps done", 4: "transition: chunk written back", 8: "encode: chunk + topology staged",
         9: "encode: warp 0 after done / reward / reset", 10: "encode: warp 0 after masks_pre", 11: "encode: board scans done (3rd barrier)",
         12: "encode: warp 0 done (masks + sampler)", 13: "encode: tile warps done (avg of 4)", 14: "encode: player / list warps done (avg of 5)"}
print("%d envs, ticks %d..%d, %.4f ms/tick (instrumented)" % (n, skip, skip + ticks, e0.elapsed_time(e1) / ticks))
for k, nm in names.items():
    s, c = int(out[2 * k]), int(out[2 * k + 1])
    if c: print("%-55s avg %8.0f cycles  (%6.2f us at 1.965 GHz)  marks/tick %d" % (nm, s / c, s / c / 1965.0, c // ticks))
