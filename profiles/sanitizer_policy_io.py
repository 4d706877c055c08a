"""compute-sanitizer driver for the rollout-side kernels added late in round 1: catan_route_by_policy, catan_policy_inputs
(all rows / indexed / value-only, fp32 and bf16, ragged batch sizes), catan_masked_categorical, catan_step_sample_host_async."""
import sys

import torch

sys.path.insert(0, ".")
from settlers_of_catan_rl_b200 import VecCatanEnv, PolicyInputs  # noqa: E402
from settlers_of_catan_rl_b200.policy_io import masked_categorical  # noqa: E402

n = 301
v = VecCatanEnv(n, seed=5)
v.reset()
a = v.sample_random()
pmap = torch.stack([torch.randperm(4, device="cuda") for _ in range(n)]).to(torch.uint8).contiguous()
pins = {dt: PolicyInputs(n, "cuda:0", dt) for dt in (torch.float32, torch.bfloat16)}
h_act = torch.empty((n, 20), dtype=torch.int32).pin_memory()
h_rew, h_info = torch.empty((n, 4)).pin_memory(), torch.empty((n, 16), dtype=torch.uint8).pin_memory()
for tick in range(40):
    v.step_sample(a)
    counts, lists = v.route_by_policy(pmap, 4)
    for dt, pin in pins.items():
        pin(v.obs, v.masks)
        pin(v.obs)
        for k, c in enumerate(counts.tolist()):
            if c:
                pin(v.obs, v.masks, index=lists[k, :c])
        for b in (1, 2, 33):
            pin(v.obs[:b].contiguous(), v.masks[:b].contiguous())
torch.cuda.synchronize()
h_act.copy_(a)
for tick in range(10):
    v.step_sample_host_async(h_act.numpy(), h_rew.numpy(), h_info.numpy())
    torch.cuda.synchronize()
for D in (2, 13, 73, 200):
    for B in (1, 7, 300):
        logits = torch.randn(B, D, device="cuda")
        mask = (torch.rand(B, D, device="cuda") < 0.5).float()
        mask[:, 0] = 1
        masked_categorical(logits, mask)
        masked_categorical(logits, None, deterministic=True)
        masked_categorical(logits, mask, actions=torch.zeros(B, 1, dtype=torch.int64, device="cuda"))
torch.cuda.synchronize()
print("done", int(v.err_flags().astype(bool).sum()))
