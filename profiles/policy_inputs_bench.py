"""Stand-alone timing of catan_policy_inputs (CUDA events, 20 launches after 5 warm-up) on random packed rows."""
import json
import sys

import torch

sys.path.insert(0, ".")
from settlers_of_catan_rl_b200 import layout as L  # noqa: E402
from settlers_of_catan_rl_b200.policy_io import PolicyInputs  # noqa: E402

out = {}
sizes = tuple(int(a) for a in sys.argv[1:]) or (16384, 65536, 131072)
for n in sizes:
    obs = torch.randint(0, 12, (n, L.OBS_STRIDE), dtype=torch.uint8, device="cuda")
    masks = torch.randint(0, 2, (n, L.MASK_STRIDE), dtype=torch.uint8, device="cuda")
    for dt, name, nbytes in ((torch.float32, "f32", 2256 + 1792 * 4 + 1000 + 325 * 4), (torch.bfloat16, "bf16", 2256 + 1792 * 2 + 1000 + 325 * 2)):
        pin = PolicyInputs(n, "cuda:0", dt)
        for _ in range(5):
            pin(obs, masks)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pin(obs, masks)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out["%s_N%d" % (name, n)] = {"ms": round(ms, 4), "GB/s": round(n * nbytes / ms / 1e6, 1), "bytes_per_row": nbytes}
        del pin
print(json.dumps(out))
