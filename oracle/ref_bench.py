"""BENCH INFRASTRUCTURE — times the UNMODIFIED reference on the host cores.

Used only by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg) and by ``tests/``.  Never on the
product path.  The reference tree is the copy ``oracle/build_ref.py`` made (``oracle/_ref/``): it travels to the
GPU box; ``bench.py`` never reads ``/root/reference``.

The loop is the one BASELINE.md §3 prescribes, per worker process (P = os.cpu_count() of them by default):
``EnvWrapper()`` with default kwargs (env/wrapper.py:12-13), seeded ``np.random.seed(1000 + rank);
random.seed(1000 + rank)`` after construction, then per tick and env

    masks = env.get_action_masks()                # env/wrapper.py:168
    obs, reward, done, _ = env.step(sample_random_legal(masks, obs))     # env/wrapper.py:36-50
    if done: obs = env.reset()

where a "tick" advances every env of the worker once, so that the tick window is comparable with the GPU arm's
(all games are pre-rolled ``preroll`` ticks from their first reset before anything is timed).
"""
from __future__ import annotations

import os
import random as _py_random
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


# --------------------------------------------------------------------------------------------------
# BASELINE.md §3 sampler on the reference's own mask list (12 float arrays, env/wrapper.py:172-185)
# --------------------------------------------------------------------------------------------------
def _pick(rng, m) -> int:
    idx = np.flatnonzero(m)
    return int(idx[rng.randrange(len(idx))]) if len(idx) else 0


def sample_random_legal(rng, masks, obs):
    """type uniform over the legal types; every dependent head uniform over its legal entries (type-conditional rows per
    RL/models/build_agent_model.py:113-127); ProposeTrade = uniform target seat, give one held resource, ask one uniform
    resource.  Returns the list of 12 ``EnvWrapper.step`` takes (RL/models/policy.py:192-199 format)."""
    a = [0] * 7 + [[0, 0, 0, 0], [0, 0, 0, 0], 0, 0, 0]
    t = _pick(rng, masks[0])
    a[0] = t
    if t == 0:
        a[1] = _pick(rng, masks[1][0])
    elif t == 2:
        a[1] = _pick(rng, masks[1][1])
    elif t == 1:
        a[2] = _pick(rng, masks[2])
    elif t == 8:
        a[3] = _pick(rng, masks[3])
    elif t == 4:
        card = _pick(rng, masks[4])
        a[4] = card
        if card == 4:                       # Monopoly -> row 2, YearOfPlenty -> row 3 (build_agent_model.py:123-124)
            a[9] = _pick(rng, masks[9][2])
        elif card == 2:
            a[9] = _pick(rng, masks[9][3])
            a[10] = _pick(rng, masks[10])
    elif t == 5:
        a[9] = _pick(rng, masks[9][0])
        a[10] = _pick(rng, masks[10])
    elif t == 6:
        a[6] = _pick(rng, masks[6][0])
        a[7][0] = 1 + _pick(rng, np.asarray(obs["current_resources"])[1:6] > 0)
        a[8][0] = 1 + rng.randrange(5)
    elif t == 7:
        a[5] = _pick(rng, masks[5])
    elif t == 11:
        a[6] = _pick(rng, masks[6][1])
    elif t == 12:
        a[11] = _pick(rng, masks[11])
    return a


# --------------------------------------------------------------------------------------------------
# worker
# --------------------------------------------------------------------------------------------------
def _worker(rank, n_envs, preroll, warmup, steps, seconds, ref_root, barrier, out_q):
    try:
        from oracle import ref_harness as H
        H.REFERENCE_ROOT = ref_root
        EnvWrapper = H.import_reference()["EnvWrapper"]
        envs = [EnvWrapper() for _ in range(n_envs)]
        np.random.seed(1000 + rank)
        _py_random.seed(1000 + rank)
        rng = _py_random.Random(7000 + rank)
        obs = [e.reset() for e in envs]
        games = 0

        def tick():
            nonlocal games
            for i, e in enumerate(envs):
                masks = e.get_action_masks()
                o, _, done, _ = e.step(sample_random_legal(rng, masks, obs[i]))
                if done:
                    o = e.reset()
                    games += 1
                obs[i] = o

        for _ in range(preroll):
            tick()
        barrier.wait()
        for _ in range(warmup):
            tick()
        barrier.wait()
        t0 = time.perf_counter()
        done_ticks = 0
        if steps is not None:
            for _ in range(steps):
                tick()
            done_ticks = steps
        else:
            while time.perf_counter() - t0 < seconds:
                tick()
                done_ticks += 1
        dt = time.perf_counter() - t0
        out_q.put((rank, done_ticks * n_envs, dt, games, None))
    except Exception as exc:                                    # reported, never hidden
        try:
            barrier.abort()
        except Exception:
            pass
        out_q.put((rank, 0, 0.0, 0, "%s: %s" % (type(exc).__name__, exc)))


def reference_available() -> bool:
    from oracle import build_ref
    return build_ref.root(copy_only=True) is not None


def time_reference_env(n_procs=None, envs_per_proc=8, preroll=1500, warmup=0, steps=None, seconds=15.0):
    """Σ env steps/s of the reference's EnvWrapper loop over ``n_procs`` worker processes.  Either ``steps`` ticks are timed
    (a tick = one step of every env of every worker) or, with ``steps=None``, as many ticks as fit ``seconds``.
    Returns a dict: value (steps/s), cores, env_steps, seconds (max over the workers), per_core, sample."""
    import multiprocessing as mp
    from oracle import build_ref
    ref_root = build_ref.root(copy_only=True)
    if ref_root is None:
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    n_procs = int(n_procs or os.cpu_count() or 1)
    ctx = mp.get_context("spawn")                               # the parent may hold a CUDA context
    barrier = ctx.Barrier(n_procs)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, envs_per_proc, preroll, warmup, steps, seconds, ref_root, barrier, q), daemon=True)
             for r in range(n_procs)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    errs = [r[4] for r in res if r[4]]
    if errs:
        raise RuntimeError("reference worker failed: " + errs[0])
    total = sum(r[1] for r in res)
    dt = max(r[2] for r in res)
    rate = sum(r[1] / r[2] for r in res)                        # Σ of the workers' own rates (they run concurrently)
    return {"value": rate, "cores": n_procs, "env_steps": total, "seconds": dt, "per_core": rate / n_procs,
            "games_finished": sum(r[3] for r in res),
            "sample": "%d procs x %d envs, %d pre-roll ticks, %d env steps in %.1f s; unmodified reference EnvWrapper loop "
                      "(get_action_masks -> random-legal sample -> step -> reset on done)" % (n_procs, envs_per_proc, preroll, total, dt)}


# --------------------------------------------------------------------------------------------------
# the reference's GAE + advantage normalisation lines on CPU tensors (RL/ppo/process_batch.py:134-142)
# --------------------------------------------------------------------------------------------------
def reference_gae_cpu(rewards, values, masks, gamma=0.999, gae_lambda=0.95):
    """the nine reference lines, verbatim in meaning, on torch CPU tensors shaped [T(+1), N, 1]"""
    import torch
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    gae = 0
    for step in reversed(range(T)):
        delta = rewards[step] + gamma * values[step + 1] * masks[step + 1] - values[step]
        gae = delta + gamma * gae_lambda * masks[step + 1] * gae
        returns[step] = gae + values[step]
    advantages = returns - values[:-1]
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-5)
    return returns, advantages


def time_reference_gae(T=200, N=131072, reps=2):
    import torch
    g = torch.Generator().manual_seed(0)
    r = torch.rand(T, N, 1, generator=g)
    v = torch.rand(T + 1, N, 1, generator=g) * 300.0
    m = (torch.rand(T + 1, N, 1, generator=g) > 0.01).float()
    reference_gae_cpu(r[:8], v[:9], m[:9])
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        reference_gae_cpu(r, v, m)
        best = min(best, time.perf_counter() - t0)
    return {"ms": best * 1e3, "T": T, "N": N, "threads": torch.get_num_threads(),
            "what": "RL/ppo/process_batch.py:134-142 on CPU fp32 tensors (torch, %d threads)" % torch.get_num_threads()}


if __name__ == "__main__":
    import json
    print(json.dumps(time_reference_env(envs_per_proc=2, preroll=50, seconds=3.0)))
