import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L
from tests.host_emu.emu_lib import EmuEnv
from tests.common import state_diff
n = 4096
env = VecCatanEnv(n, seed=12, first_env_id=77); env.reset(); a = env.sample_random()
for _ in range(700): env.step_sample(a)
before = env.export_state()
ctrl = torch.from_numpy((np.arange(n) % 5).astype(np.uint8)).cuda()
env.err_flags(clear=True)
env.randomise_uncertainty(ctrl, max_attempts=300)
after = env.export_state()
c = ctrl.cpu().numpy()
nbad = 0
for e in range(1, 200):
    if c[e] == 0: continue
    emu = EmuEnv(seed=12, env_id=77 + e, auto_reset=0)
    emu.import_state(before[e])
    rt = np.array_equal(emu.state(), before[e])
    na = emu.randomise_uncertainty(int(c[e]), 300)
    same = np.array_equal(emu.state(), after[e])
    # the same game alone in a fresh handle
    one = VecCatanEnv(1, seed=12, first_env_id=77 + e, auto_reset=0); one.reset(); one.import_state(before[e:e + 1])
    one.randomise_uncertainty(int(c[e]), 300)
    solo = one.export_state()[0]
    one.close()
    if not same or not rt:
        nbad += 1
        if nbad <= 6:
            print("game", e, "ctrl", c[e], "roundtrip", rt, "emu attempts", na, "emu==gpu", same, "solo==emu", np.array_equal(solo, emu.state()), "solo==gpu", np.array_equal(solo, after[e]),
                  "ctr before/emu/gpu/solo", before[e][-2], emu.state()[-2], after[e][-2], solo[-2])
            print("   ", state_diff(emu.state(), after[e])[:3])
print("bad", nbad)
