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
         12: "encode: mask warps done (avg of 4)", 15: "encode: LAST warp of the block done", 13: "encode: row warps after their tile part (avg of 4)", 14: "encode: row warps done (avg of 4)"}
print("%d envs, ticks %d..%d, %.4f ms/tick (instrumented)" % (n, skip, skip + ticks, e0.elapsed_time(e1) / ticks))
for k, nm in names.items():
    s, c = int(out[2 * k]), int(out[2 * k + 1])
    if c: print("%-55s avg %8.0f cycles  (%6.2f us at 1.965 GHz)  marks/tick %d" % (nm, s / c, s / c / 1965.0, c // ticks))
print("encode: longest block %.1f us; blocks with a reset per tick %.1f; block duration histogram (<15, <30, <46, <61, <102, more us): %s" % (
    int(out[40]) / 1965.0, int(out[41]) / ticks, [int(x) // ticks for x in out[42:48]]))
if int(out[33]):
    print("resets (one game per block, own stream): avg %.1f us, longest %.1f us, per tick %.1f; histogram (<10, <20, <41, <81, <163, more us): %s" % (
        int(out[32]) / int(out[33]) / 1965.0, int(out[48]) / 1965.0, int(out[33]) / ticks, [round(int(x) / ticks, 2) for x in out[50:56]]))
if int(out[33]):
    n = int(out[33])
    print("reset parts, avg us per reset: clear + pre-draw %.1f | terrain + numbers shuffle %.1f | 6/8 rejection loop %.1f (%.2f retries) | rest %.1f" % (
        int(out[56]) / n / 1965.0, int(out[57]) / n / 1965.0, int(out[58]) / n / 1965.0, int(out[60]) / n, int(out[59]) / n / 1965.0))
