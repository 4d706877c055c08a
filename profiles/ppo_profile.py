"""Where the time of a PPO tick / minibatch step goes: torch profiler (CUDA time per kernel) over a few un-graphed rollout ticks
and a few forward+backward micro-batches.  usage: python profiles/ppo_profile.py [envs] [fp32|bf16] [micro]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from settlers_of_catan_rl_b200 import SelfPlayTrainer, PPOConfig, CatanPolicy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
micro = int(sys.argv[3]) if len(sys.argv) > 3 else 51200
torch.manual_seed(0)
cfg = PPOConfig(num_steps=8, dtype=dtype, graph=False, micro_batch=micro, num_mini_batch=2)
tr = SelfPlayTrainer(n, CatanPolicy(), cfg, seed=0)
tr.store.begin(fresh=True)
for _ in range(40):
    tr._tick()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        tr._tick()
    torch.cuda.synchronize()
print("==== 5 rollout ticks, %d envs, %s" % (n, dtype))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    tr._tick()
e1.record(); torch.cuda.synchronize()
print("ms per un-graphed tick:", e0.elapsed_time(e1) / 20)
tr.collect()
tr.warmup_update()
returns, adv = tr.compute_advantages()
tr.policy.train()
perm = torch.randperm(tr.T * tr.N, device="cuda").to(torch.int32)
def step(k):
    idx = perm[k * tr.micro:(k + 1) * tr.micro].contiguous()
    m = tr.store.gather(idx, tr.values, returns, adv, out=tr._mb_out(idx.numel()))
    obs, masks = tr.mb_inputs(m["obs"], m["masks"])
    with tr._autocast():
        values, logp, entropy = tr.policy.evaluate_actions(obs, masks, m["actions"])
    (values.mean() + logp.mean() + entropy).backward()
step(0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(1)
    torch.cuda.synchronize()
print("==== one forward+backward micro-batch of %d rows" % tr.micro)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
