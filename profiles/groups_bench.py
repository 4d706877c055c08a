"""device-resident step rate with the games split over G handles on G streams (each replays its own step graph):
python profiles/groups_bench.py [envs] [skip] [ticks] [G ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from settlers_of_catan_rl_b200 import VecCatanEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 500
for G in [int(x) for x in sys.argv[4:]] or [1, 2, 3, 4]:
    sizes = [n // G + (1 if g < n % G else 0) for g in range(G)]
    envs, acts, streams = [], [], []
    for g in range(G):
        e = VecCatanEnv(sizes[g], seed=0, first_env_id=sum(sizes[:g])); e.set_graphs(True); e.reset()
        a = e.sample_random()
        for _ in range(skip): e.step_sample(a)
        envs.append(e); acts.append(a); streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()
    def run(k):
        for _ in range(k):
            for e, a, s in zip(envs, acts, streams):
                with torch.cuda.stream(s):
                    e.step_sample(a)
    run(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for s in streams: s.wait_event(e0)
    run(ticks)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ticks
    print("G=%d: %.4f ms per step of all %d envs = %.1f M env steps/s  errs %d" % (G, ms, n, n / ms / 1e3, sum(int(e.err_flags().any()) for e in envs)))
    for e in envs: e.close()
