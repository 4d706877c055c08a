"""Step time against the number of games: the encode kernel runs 4 blocks of one 32-game chunk per SM (592 slots), so 65 536
games = 2 048 chunks = 3.46 waves cost four block latencies; 56 832 games are exactly 3 waves, 75 776 exactly 4."""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from settlers_of_catan_rl_b200 import VecCatanEnv  # noqa: E402

out = {}
for n in (37888, 56832, 65536, 75776, 94720):
    env = VecCatanEnv(n, seed=0)
    env.reset()
    acts = env.sample_random()
    for _ in range(1200):
        env.step_sample(acts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.set_timing(True)
    e0.record()
    for _ in range(600):
        env.step_sample(acts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 600
    _, t_ms, enc_ms = env.read_timing()
    out[str(n)] = {"chunks": n // 32, "encode_waves": round(n / 32 / 592, 2), "ms_per_step": round(ms, 4), "transition_ms": round(t_ms, 4),
                   "encode_ms": round(enc_ms, 4), "M_env_steps_per_s": round(n / ms / 1e3, 1)}
    env.close()
print(json.dumps(out))
