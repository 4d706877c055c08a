"""One launch (after a warm-up) of every kernel of the path OUTSIDE the env step, at benchmark sizes, inside a
cudaProfilerStart/Stop range:

  ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/aux python profiles/aux_kernels.py

rollout_store (16 384 envs, T = 200), minibatch_gather (51 200 rows), route (65 536 envs, 4 policies), policy_inputs fp32 / bf16
(65 536 rows), masked_categorical (65 536 x 54, sample and evaluate), the stand-alone sampler, GAE and advantage statistics /
normalisation (T = 200, N = 131 072), the reset encode (MODE_RESET, 65 536 games), randomise_uncertainty (65 536 games),
the tile attention and small LayerNorm of the policy network (16 384 samples)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from settlers_of_catan_rl_b200 import VecCatanEnv, RolloutStorage, gae, normalise_advantages, layout as L
from settlers_of_catan_rl_b200.policy_io import PolicyInputs, masked_categorical
from settlers_of_catan_rl_b200 import policy_ops

dev = "cuda:0"
env = VecCatanEnv(65536, device=dev, seed=0); env.reset(); acts = env.sample_random()
for _ in range(600): env.step_sample(acts)
small = VecCatanEnv(16384, device=dev, seed=1); small.reset()
st = RolloutStorage(small, 200)
st.begin(True)
a16 = small.sample_random(); logp = torch.zeros(16384, device=dev)
def tick():
    small.step(a16); st.record(a16, logp); small.sample_random(a16)
for _ in range(20): tick()
vals = torch.rand(201, 16384, device=dev); rets = torch.rand(200, 16384, device=dev); adv = torch.rand(200, 16384, device=dev)
perm = torch.randperm(200 * 16384, device=dev)[:51200].to(torch.int32).contiguous()
mb = st.gather(perm, vals, rets, adv)
pmap = torch.randint(0, 4, (65536, 4), device=dev, dtype=torch.uint8)
pin32, pin16 = PolicyInputs(65536, dev, torch.float32), PolicyInputs(65536, dev, torch.bfloat16)
logits = torch.randn(65536, 54, device=dev); cmask = (torch.rand(65536, 54, device=dev) > 0.5).float(); cmask[:, 0] = 1
T, N = 200, 131072
r = torch.rand(T, N, device=dev); v = torch.rand(T + 1, N, device=dev) * 300; m = (torch.rand(T + 1, N, device=dev) > 0.01).float()
ret, ad = torch.empty_like(r), torch.empty_like(r)
qkv = torch.randn(16384, 19, 192, device=dev); x16 = torch.randn(16384 * 19, 64, device=dev); w = torch.ones(64, device=dev); b = torch.zeros(64, device=dev)
ctrl = torch.ones(65536, dtype=torch.uint8, device=dev)

def everything():
    tick()
    st.gather(perm, vals, rets, adv, out=mb)
    env.route_by_policy(pmap, 4)
    pin32(env.obs, env.masks); pin16(env.obs, env.masks)
    got = masked_categorical(logits, cmask)
    masked_categorical(logits, cmask, actions=got[0])
    env.sample_random(acts)
    gae(r, v, m, 0.999, 0.95, ret, ad); normalise_advantages(ad)
    policy_ops.tile_attention(qkv); policy_ops.layer_norm_small(x16, w, b, 1e-5)
everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
everything()
env.randomise_uncertainty(ctrl, max_attempts=50)
env.reset()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
