"""End-to-end rate of the double-buffered host loop (bench.e2e_double_buffered) for 1-4 env groups in flight."""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402

out = {}
for g in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (1, 2, 3, 4):
    for fused in (False, True):
        rate, errs = bench.e2e_double_buffered(65536, 300, 1000, torch.device("cuda:0"), 0, n_groups=g, fused_sampler=fused)
        out["groups_%d_%s" % (g, "fused_sampler" if fused else "split_sampler")] = {"env_steps_per_s": round(rate), "rejected": errs}
print(json.dumps(out))
