#!/usr/bin/env python
"""bench.py — headline benchmark of the Catan env-step hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference's algorithm on host cores

Workload (config.workload): 65 536 parallel 4-player games per GPU, random-legal policy, one "step" = one
lock-step tick of every game (apply action -> auto-reset -> legal-action masks -> packed observation -> next
random-legal action).  metric = env steps per second, whole job.

STEADY STATE, whatever --steps / --warmup are: every leg of every arm first advances its games `--preroll` ticks
(default 1500) from their first reset, untimed, so that the timed window is ticks [preroll + warmup, preroll + warmup +
steps) of games in every phase and not the opening of freshly reset games; `config.tick_window`
records it.  Prints ONE JSON line on rank 0.  See DESIGN.md §6 for how every field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ENVS_PER_GPU = 65536
METRIC = "env_steps_per_sec"
UNIT = "env steps/s"
# algorithmic bytes per env step with the int8 obs/mask layout =
# obs 1867 + masks 325 + action 80 + state 640 read + 640 write
ALGO_BYTES_PER_ENV_STEP = 3552
# the same per kernel: transition_kernel reads the action row and reads + writes the state; encode_kernel reads the state
# and writes the obs row, the mask row and the next action row
TRANSITION_BYTES_PER_ENV_STEP = 80 + 640 + 640
ENCODE_BYTES_PER_ENV_STEP = 640 + 1867 + 325 + 80
ROWS_BYTES_PER_ENV_STEP = 640 + 1867            # the observation-rows launch: record in, obs row out
MASKS_BYTES_PER_ENV_STEP = 640 + 325 + 80         # the masks + sampler launch: record in, mask row + next action out
WORKLOAD = "65536 parallel envs per GPU, random-legal policy, fused step+auto-reset+masks+obs+sample (BASELINE configs[1])"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", choices=["env", "ppo16k", "ppo131k", "ppo8x"], default="env",
                    help="env: BASELINE configs[1] (the headline, default); ppo16k / ppo131k / ppo8x: BASELINE configs[2] / [3] / [4], "
                         "whole PPO self-play cycles (rollouts + 10 epochs x 64 minibatches), --steps = cycles timed (default 1)")
    ap.add_argument("--ppo-envs", type=int, default=0, help="PPO configs: envs per GPU (0: the BASELINE config's)")
    ap.add_argument("--ppo-micro-batch", type=int, default=51200)
    ap.add_argument("--checkpoint", default="", help="PPO configs: a reference state_dict to start from (default: seeded random init)")
    ap.add_argument("--envs", type=int, default=N_ENVS_PER_GPU, help="games per GPU (default: the BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--e2e-groups", type=int, default=6, help="handles (env groups) kept in flight by the double-buffered e2e loop")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--preroll", type=int, default=1500, help="untimed ticks every leg plays from the first reset before anything is timed")
    ap.add_argument("--ref-envs-per-proc", type=int, default=0, help="reference arm: envs per worker process (0: sized from --steps)")
    ap.add_argument("--no-python-reference", action="store_true", help="cpu_baseline / reference arm: the C port only")
    ap.add_argument("--no-graphs", action="store_true", help="issue every launch of a step directly instead of replaying its CUDA graph")
    return ap.parse_args()


def bench_config(args):
    """`config` of the JSON line: identical keys and values in both arms"""
    w = max(3, args.warmup)
    return {"workload": WORKLOAD, "envs_per_gpu": args.envs, "seed": args.seed,
            "tick_window": [args.preroll + w, args.preroll + w + args.steps],
            "l2": "no explicit flush: one step streams >= 230 MB (records+obs+masks+actions) > 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref, Python, all host cores) and, beside it, the pinned C port of its algorithm
# ------------------------------------------------------------------------------------------------
def cpu_port_rate(seconds: float, preroll: int, n_envs: int = 4096, chunk: int = 25, seed: int = 0):
    """the C oracle's sample->step->masks->obs loop on all host threads for ~`seconds`, from tick `preroll` on"""
    from oracle import oracle_lib as O
    v = O.OracleVec(n_envs, seed=seed, first_env_id=0)
    v.run(0)
    v.run(preroll)                                 # untimed: the same steady state as every other leg
    done_steps, t0 = 0, time.perf_counter()
    while True:
        v.run(chunk)
        done_steps += n_envs * chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return {"value": done_steps / dt, "unit": UNIT, "cores": v.threads_used, "kind": "port",
            "sample": "%d envs x %d ticks from tick %d, %.1f s, C port of game.py+wrapper.py (oracle/catan_oracle.c) on %d host threads" % (
                n_envs, done_steps // n_envs, preroll, dt, v.threads_used)}


def python_reference_rate(args, steps=None, seconds=15.0, envs_per_proc=8, warmup=0):
    """the unmodified reference's EnvWrapper loop (oracle/ref_bench.py) on os.cpu_count() processes"""
    from oracle import ref_bench
    r = ref_bench.time_reference_env(envs_per_proc=envs_per_proc, preroll=args.preroll, warmup=warmup, steps=steps, seconds=seconds)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"],
            "per_core": r["per_core"], "seconds": r["seconds"], "env_steps": r["env_steps"]}


def python_reference_available(args) -> bool:
    if args.no_python_reference:
        return False
    try:
        from oracle import ref_bench
        return ref_bench.reference_available()
    except Exception:
        return False


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = max(3, args.warmup)
    port = None
    if python_reference_available(args):
        # a "step" = one tick of every env of every worker; envs per worker sized so that the timed region is ~10 s of
        # host work when --steps is small, and the untimed pre-roll (1500 ticks of Python) stays near a minute
        e = args.ref_envs_per_proc or int(max(8, min(64, 12000 // max(1, args.steps))))
        base = python_reference_rate(args, steps=args.steps, envs_per_proc=e, warmup=w)
        port = cpu_port_rate(min(args.cpu_seconds, 6.0), args.preroll, seed=args.seed)
        dt = base["seconds"]
    else:
        from oracle import oracle_lib as O
        rate = cpu_port_rate(2.0, 0, n_envs=2048, chunk=10, seed=args.seed)["value"]
        total_ticks = max(1, args.steps + w + args.preroll)
        n_envs = int(max(64, min(args.envs, rate * 60.0 / total_ticks)))
        n_envs = max(64, (n_envs // 64) * 64)
        v = O.OracleVec(n_envs, seed=args.seed, first_env_id=0)
        v.run(0)
        v.run(args.preroll)
        for _ in range(w):
            v.run(1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            v.run(1)
        dt = time.perf_counter() - t0
        base = {"value": n_envs * args.steps / dt, "unit": UNIT, "cores": v.threads_used, "kind": "port",
                "sample": "%d envs x %d ticks per timed run (one tick per step), C port on all %d host threads" % (n_envs, args.steps, v.threads_used)}
    value = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": w, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args),
        "cpu_baseline": base,
        "cpu_baseline_port": port,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "kind=reference: the UNMODIFIED Python reference (oracle/_ref, copied from the reference tree by oracle/build_ref.py) on "
                "os.cpu_count() processes, a bounded sample of the workload (the per-step sample is in cpu_baseline.sample); "
                "cpu_baseline_port: the pinned C restatement of the same algorithm on all host threads, for scale",
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# this build
# ------------------------------------------------------------------------------------------------
def e2e_double_buffered(n, steps, warm_ticks, dev, seed, rank=0, world=1, n_groups=2, barrier=None, fused_sampler=True, graphs=True, packed=True):
    """The games split over n_groups handles kept in flight from the host (HostEnvGroups: own stream and pinned buffers each).  One
    library call per round (catan_step_sample_host_groups) waits for each group's previous result, reads its done flags on the host
    and issues its next step from its pinned actions."""
    import torch
    from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L
    from settlers_of_catan_rl_b200.vec_env import HostEnvGroups
    sizes = [n // n_groups + (1 if gi < n % n_groups else 0) for gi in range(n_groups)]     # sum = n
    envs = []
    for gi in range(n_groups):
        e = VecCatanEnv(sizes[gi], device=dev, seed=seed, first_env_id=rank * n + sum(sizes[:gi]))
        e.set_graphs(graphs)
        e.reset()
        a = e.sample_random()
        for _ in range(warm_ticks):
            e.step_sample(a)
        envs.append(e)
    torch.cuda.synchronize()
    groups = HostEnvGroups(envs, packed_actions=packed)
    groups.prime()
    groups.synchronize()
    groups.pump(3)
    if barrier is not None:
        barrier()
    t0 = time.perf_counter()
    for it in range(steps):
        groups.pump(1)                                         # (Python gets every round's results: one ctypes call per tick)
    groups.synchronize()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    errs_e2e = int(sum(int(e.err_flags().any()) for e in envs))
    for e in envs:
        e.close()
    return n * world * steps / float(te.item()), errs_e2e


def _lib_record_bytes() -> int:
    from settlers_of_catan_rl_b200 import _lib
    return int(_lib.load().catan_record_bytes())


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L, gae, normalise_advantages

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.envs
    rec_bytes = _lib_record_bytes()
    # games are numbered globally: rank r owns [r*n, (r+1)*n) — no data-path collective
    env = VecCatanEnv(n, device=dev, seed=args.seed, first_env_id=rank * n)
    env.set_graphs(not args.no_graphs)             # one cudaGraphLaunch per step instead of ~20 driver calls (the same launches)
    env.reset()
    acts = env.sample_random()
    warm = max(3, args.warmup)
    # untimed pre-roll to the steady state: the games are in every phase by tick `preroll`.  The clock
    # sampler (nvidia-smi every 100 ms) is started for the last CLOCK_LEAD ticks of it, runs through the timed region and a
    # tail of CLOCK_TAIL further ticks of the same workload, so that it sees >= 1 s of this load even when --steps is small.
    CLOCK_LEAD, CLOCK_TAIL = min(1000, args.preroll), 2000
    for _ in range(args.preroll - CLOCK_LEAD):
        env.step_sample(acts)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(CLOCK_LEAD + warm):
        env.step_sample(acts)
    launches0 = env.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        env.step_sample(acts)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.kernel_launches - launches0
    for _ in range(CLOCK_TAIL):
        env.step_sample(acts)
    torch.cuda.synchronize()
    clock_info = clocks.stop() if rank == 0 else None
    if clock_info is not None:
        clock_info["window"] = "last %d pre-roll ticks + warm-up + the timed region + %d further ticks of the same workload" % (CLOCK_LEAD, CLOCK_TAIL)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # per-kernel durations for the roofline: CUDA events recorded by the library around the two kernels of a step on this
    # stream (catan_set_timing), over `kernel_steps` further steps of the same games -- outside the timed region above
    kernel_steps = 200
    env.set_timing(True)
    for _ in range(kernel_steps):
        env.step_sample(acts)
    timed_n, transition_ms, encode_ms, rows_ms = env.read_timing(detail=True)
    env.set_timing(False)
    games_done = int(env.info[:, L.INFO_DONE].sum().item())  # touch the result
    errs = int(env.err_flags().any())

    # ---- e2e: the same tick through the host-buffer call of the C ABI (catan_step_host), every step:
    #   inputs : the step's actions come from PINNED HOST memory (H2D inside the call);
    #   result : the step's reward + done/info rows are read back to pinned host memory (D2H inside the call) and the
    #            call synchronises.  Observations and masks stay in HBM, which is the point of the design: they are the
    #            GPU-resident policy's input (north_star).  The variant that also ships every observation and mask row
    #            to the host (the EnvWrapper-shaped adapter's call; PCIe-bound) is reported as e2e_full_obs_to_host.
    e2e_steps = max(1, args.e2e_steps)
    h_act = torch.empty((n, L.ACTION_WORDS), dtype=torch.int32).pin_memory()
    h_obs = torch.empty((n, L.OBS_STRIDE), dtype=torch.uint8).pin_memory()
    h_masks = torch.empty((n, L.MASK_STRIDE), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    h_info = torch.empty((n, L.INFO_STRIDE), dtype=torch.uint8).pin_memory()
    d_act = torch.empty((n, L.ACTION_WORDS), dtype=torch.int32, device=dev)

    def e2e_tick(full):
        env.sample_random(d_act)                       # stand-in for the policy; its actions are brought to the host ...
        h_act.copy_(d_act, non_blocking=False)
        # ... and handed to the host-buffer call
        if full:
            env.step_host(h_act.numpy(), h_obs.numpy(), h_masks.numpy(), h_rew.numpy(), h_info.numpy())
        else:
            env.step_host(h_act.numpy(), None, None, h_rew.numpy(), h_info.numpy())

    def time_e2e(full, steps):
        for _ in range(3):
            e2e_tick(full)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_tick(full)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return n * world * steps / float(te.item())

    e2e_value = time_e2e(False, e2e_steps * 10)
    e2e_full = time_e2e(True, e2e_steps)

    # ---- e2e, double-buffered: the same per-step traffic (actions H2D from pinned memory, reward + done/info rows and the
    # sampler's actions D2H to pinned memory, every step, for every game), with the games split over TWO handles on two
    # streams (catan_step_host_async): while the host waits for one half's result, the other half's kernels and copies run.
    # This is how a host-side policy loop drives the library (the reference's sub-process manager also keeps several env
    # groups in flight, RL/ppo/vec_gather_experience.py); each half's step still depends on that half's previous result.
    e2e_pipe_error = None
    try:
        e2e_pipe, e2e_pipe_errs = e2e_double_buffered(n, e2e_steps * 10, args.preroll + warm, dev, args.seed, rank, world,
                                                      max(1, args.e2e_groups), barrier, graphs=not args.no_graphs)
    except Exception as exc:  # reported, never hidden: the synchronous loop above then stands as e2e
        e2e_pipe, e2e_pipe_errs, e2e_pipe_error = 0.0, -1, "%s: %s" % (type(exc).__name__, exc)
    h2d = n * L.ACTION_WORDS * 4
    d2h = n * (16 + L.INFO_STRIDE) + n * L.ACTION_WORDS * 4
    d2h_full = d2h + n * (L.OBS_STRIDE + L.MASK_STRIDE)

    # ---- PPO-side kernels at BASELINE config 4 size (T=200, N=131072), rank 0, reported as aux numbers
    aux = {}
    if rank == 0:
        T, N = 200, 131072
        r = torch.rand(T, N, device=dev)
        val = torch.rand(T + 1, N, device=dev) * 300.0
        m = (torch.rand(T + 1, N, device=dev) > 0.01).float()
        ret, adv = torch.empty_like(r), torch.empty_like(r)
        for _ in range(3):
            gae(r, val, m, 0.999, 0.95, ret, adv)
            normalise_advantages(adv)
        torch.cuda.synchronize()
        a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 10
        gae_ms = norm_ms = 0.0
        for _ in range(reps):
            a0.record()
            gae(r, val, m, 0.999, 0.95, ret, adv)
            a1.record()
            normalise_advantages(adv)
            a2.record()
            torch.cuda.synchronize()
            gae_ms += a0.elapsed_time(a1)
            norm_ms += a1.elapsed_time(a2)
        gae_ms /= reps
        norm_ms /= reps
        elems = T * N
        aux = {"gae_T200_N131072": {"ms": gae_ms, "GB/s": elems * 20 / gae_ms / 1e6, "bytes_per_elem": 20},
               "adv_norm_T200_N131072": {"ms": norm_ms, "GB/s": elems * 12 / norm_ms / 1e6, "bytes_per_elem": 12}}
        del r, val, m, ret, adv
        # minibatch gather (process_batch.py:169-200) on a rollout buffer that is larger than L2: T=16, N=16384 (0.6 GB)
        from settlers_of_catan_rl_b200 import RolloutStorage
        Tg, Ng = 16, 16384
        genv = VecCatanEnv(Ng, device=dev, seed=args.seed)
        st = RolloutStorage(genv, Tg)
        gv = torch.rand(Tg + 1, Ng, device=dev)
        gr, ga = torch.rand(Tg, Ng, device=dev), torch.rand(Tg, Ng, device=dev)
        perm = torch.randperm(Tg * Ng, device=dev).to(torch.int32)
        out = st.gather(perm, gv, gr, ga)
        for _ in range(3):
            st.gather(perm, gv, gr, ga, out=out)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            st.gather(perm, gv, gr, ga, out=out)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1) / 10
        aux["minibatch_gather_T16_N16384"] = {"ms": gather_ms, "GB/s": Tg * Ng * 4712 / gather_ms / 1e6, "bytes_per_row": 4712,
                                              "rows": Tg * Ng}
        del st, genv, out
        # policy inputs (policy.py:168-190 batched): the env's own 65 536 packed rows -> fp32 / bf16 policy tensors, one launch
        from settlers_of_catan_rl_b200.policy_io import PolicyInputs
        for pdt, pname, pbytes in ((torch.float32, "f32", 2256 + 1792 * 4 + 1000 + 325 * 4), (torch.bfloat16, "bf16", 2256 + 1792 * 2 + 1000 + 325 * 2)):
            pin = PolicyInputs(n, dev, pdt)
            for _ in range(3):
                pin(env.obs, env.masks)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(10):
                pin(env.obs, env.masks)
            g1.record()
            torch.cuda.synchronize()
            pms = g0.elapsed_time(g1) / 10
            aux["policy_inputs_%s_N%d" % (pname, n)] = {"ms": pms, "GB/s": n * pbytes / pms / 1e6, "bytes_per_row": pbytes}
            del pin

    # ---- cpu_baseline (rank 0, N=1 only): the UNMODIFIED reference on the box's host cores, the pinned C port beside it, and
    # the reference's GAE + advantage-normalisation lines on CPU tensors beside aux.gae_* / aux.adv_norm_*
    cpu_baseline = cpu_baseline_port = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline_port = cpu_port_rate(min(args.cpu_seconds, 8.0), args.preroll, seed=args.seed)
        if python_reference_available(args):
            try:
                cpu_baseline = python_reference_rate(args, steps=None, seconds=args.cpu_seconds, envs_per_proc=8)
            except Exception as exc:  # reported, never hidden: the port then stands as the baseline
                cpu_baseline_port["python_reference_error"] = "%s: %s" % (type(exc).__name__, exc)
        if cpu_baseline is None:
            cpu_baseline, cpu_baseline_port = cpu_baseline_port, None
        try:
            from oracle import ref_bench
            g = ref_bench.time_reference_gae(200, 131072, reps=1)
            aux["gae_plus_adv_norm_cpu_reference_T200_N131072"] = g
            if "gae_T200_N131072" in aux:
                aux["gae_plus_adv_norm_T200_N131072_speedup_vs_cpu_reference"] = g["ms"] / (aux["gae_T200_N131072"]["ms"] + aux["adv_norm_T200_N131072"]["ms"])
        except Exception as exc:
            aux["gae_cpu_reference_error"] = "%s: %s" % (type(exc).__name__, exc)

    e2e_sync = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps * 10,
                "call": "VecCatanEnv.step_host -> catan_step_host, one handle: pinned host actions in, reward+done/info rows out to "
                        "pinned host, synchronous; obs/masks stay in HBM for the GPU policy (d2h also counts the sampler's actions)"}
    e2e_db = {"value": e2e_pipe, "unit": UNIT, "h2d_bytes_per_step": n * L.ACTION_WORDS, "d2h_bytes_per_step": n * (16 + L.INFO_STRIDE) + n * L.ACTION_WORDS,
              "steps": e2e_steps * 10, "action_rows": "uint8 [N, 20] on the host side (catan_step_sample_host_async_u8: one byte per word, expanded on the device)",
              "rejected_actions": e2e_pipe_errs, "error": e2e_pipe_error, "groups": max(1, args.e2e_groups),
              "call": "HostEnvGroups.pump -> catan_step_sample_host_groups -> catan_step_sample_host_async_u8 per handle (the random-legal policy fused into the step as in `value`), pipelined: the games split over %d handles on their own "
                      "streams, every step of every game still takes its actions from pinned host memory and returns reward+done/info "
                      "rows (and the sampler's next actions) to pinned host memory; the host waits for one group while the others run" % args.e2e_groups}
    e2e_best, e2e_other = (e2e_db, e2e_sync) if e2e_pipe >= e2e_value else (e2e_sync, e2e_db)
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        per_launch_ms = ms_max / args.steps
        # dominant kernel = the observation-rows launch of the encode: its algorithmic bytes per env step over its own duration
        achieved = ROWS_BYTES_PER_ENV_STEP * n / (rows_ms * 1e-3) / 1e9
        masks_ms = encode_ms - rows_ms
        step_achieved = ALGO_BYTES_PER_ENV_STEP * n / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("encode_rows_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        value = n * world * args.steps / (ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": bench_config(args),
            "detail": {"games_finished_in_last_step": games_done, "rejected_actions": errs, "preroll_ticks": args.preroll,
                       "streamed_MB_per_step": n * (2 * rec_bytes + L.OBS_STRIDE + L.MASK_STRIDE + 160 + 32) / 1e6},
            "clocks": clock_info,
            "e2e": e2e_best,
            "e2e_other": e2e_other,
            "e2e_full_obs_to_host": {"value": e2e_full, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_full,
                                     "steps": e2e_steps, "call": "same call with obs+masks rows also copied to pinned host (PCIe-bound)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/traffic.json (one ncu --set full capture of this kernel, dram__bytes_read.sum + dram__bytes_write.sum)",
                         "obs_write_only": {"bytes_per_env_step": 1867, "achieved": 1867 * n / (per_launch_ms * 1e-3) / 1e9,
                                            "frac": 1867 * n / (per_launch_ms * 1e-3) / 1e9 / peak,
                                            "note": "north_star's obs-write roofline: the 1 867 obs bytes of every env step over the WHOLE step time"},
                         "kernel": "encode_kernel<MODE_STEP, SAMPLE, ROLE_ROWS> (the observation rows: 640 B of record in, 1 867 B out per env step)",
                         "algorithmic_bytes_per_launch": ROWS_BYTES_PER_ENV_STEP * n, "peak_source": peak_src,
                         "kernel_ms": rows_ms, "timed_launches": timed_n,
                         "how": "CUDA events recorded by the library around each kernel on the launching stream (catan_set_timing: the rows launch "
                                "then runs in front of the masks launch on the caller's stream; in production it runs beside it on a library stream)",
                         "encode_masks_kernel": {"ms": masks_ms, "algorithmic_bytes_per_launch": MASKS_BYTES_PER_ENV_STEP * n,
                                                 "achieved": MASKS_BYTES_PER_ENV_STEP * n / (masks_ms * 1e-3) / 1e9},
                         "encode_both_launches": {"ms": encode_ms, "algorithmic_bytes_per_launch": ENCODE_BYTES_PER_ENV_STEP * n,
                                                  "achieved": ENCODE_BYTES_PER_ENV_STEP * n / (encode_ms * 1e-3) / 1e9,
                                                  "frac": ENCODE_BYTES_PER_ENV_STEP * n / (encode_ms * 1e-3) / 1e9 / peak},
                         "transition_kernel": {"ms": transition_ms, "algorithmic_bytes_per_launch": TRANSITION_BYTES_PER_ENV_STEP * n,
                                               "achieved": TRANSITION_BYTES_PER_ENV_STEP * n / (transition_ms * 1e-3) / 1e9},
                         "whole_step": {"ms": per_launch_ms, "algorithmic_bytes": ALGO_BYTES_PER_ENV_STEP * n,
                                        "achieved": step_achieved, "frac": step_achieved / peak,
                                        "note": "6 launches on 3 streams, replayed as one CUDA graph: transition, observation rows, masks + sampler | longest-road "
                                                "search, encode of the searched games | encode (done / reset / new game) of the games that ended"}},
            "cpu_baseline": cpu_baseline,
            "cpu_baseline_port": cpu_baseline_port,
            "aux": aux,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# PPO self-play configs (BASELINE configs[2..4]): env steps/s AND PPO updates/s over whole run_update cycles
# ------------------------------------------------------------------------------------------------
PPO_CONFIGS = {
    "ppo16k": dict(envs=16384, dtype="float32", workload="PPO self-play, repo default policy net (1 928 995 parameters, fp32), 16384 envs, 1xB200 (BASELINE configs[2])"),
    "ppo131k": dict(envs=131072, dtype="bfloat16", workload="PPO self-play, 131072 envs, bf16 policy (autocast), 1xB200, GAE + advantage-norm kernels (BASELINE configs[3])"),
    "ppo8x": dict(envs=65536, dtype="bfloat16", workload="PPO self-play, 65536 envs per GPU (524288 over 8), bf16 policy, NCCL all-reduce of the flat gradient bucket (BASELINE configs[4])"),
}


def run_ppo_arm(args):
    import torch
    import torch.distributed as dist
    from settlers_of_catan_rl_b200 import SelfPlayTrainer, PPOConfig, CatanPolicy

    pc = PPO_CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = args.ppo_envs or pc["envs"]
    dtype = getattr(torch, pc["dtype"])
    torch.manual_seed(args.seed)
    policy = CatanPolicy()
    if args.checkpoint:
        policy.load_reference_state_dict(torch.load(args.checkpoint, map_location="cpu", weights_only=False))
    cfg = PPOConfig(dtype=dtype, micro_batch=args.ppo_micro_batch)
    tr = SelfPlayTrainer(n, policy, cfg, device=dev, seed=args.seed, first_env_id=rank * n, group=True if world > 1 else None)
    cycles = args.steps if args.steps != 2000 else 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up, untimed: graph capture + one whole rollout (games leave the opening), and two optimiser steps whose effect is undone
    t0 = time.perf_counter()
    tr.collect()
    tr.warmup_update()
    warm_s = time.perf_counter() - t0
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    outs = [tr.run_update() for _ in range(cycles)]
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    steps = torch.tensor([float(sum(o["env_steps"] for o in outs))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps, op=dist.ReduceOp.SUM)
    clock_info = clocks.stop() if rank == 0 else None
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        o = outs[-1]
        opt_steps = sum(x["optimiser_steps"] for x in outs)
        line = {
            "metric": METRIC, "value": float(steps.item()) / sec, "unit": UNIT, "n_gpus": world, "steps": cycles, "warmup": 1,
            "ms_per_step": 1e3 * sec / cycles, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if dtype == torch.bfloat16 else "fp32", "data": "synthetic (seeded random-init policy weights; games generated by the engine)",
            "config": {"workload": pc["workload"], "envs_per_gpu": n, "seed": args.seed, "num_steps": cfg.num_steps, "ppo_epoch": cfg.ppo_epoch,
                       "num_mini_batch": cfg.num_mini_batch, "micro_batch": tr.micro, "step": "one run_update cycle (robust_train.py:95-156): rollouts until every env "
                       "holds 200 decisions of its recorded seat, then 10 epochs x (value pass over 201 x N obs + GAE + advantage norm + 64 minibatch steps)"},
            "ppo": {"updates_per_sec": cycles / sec, "optimiser_steps_per_sec": opt_steps / sec, "recorded_decisions_per_sec": cycles * tr.T * n * world / sec,
                    "collect_ms": o["collect_ms"], "update_ms": o["update_ms"], "ticks_per_rollout": o["ticks"], "env_steps_per_rollout_per_gpu": o["env_steps"],
                    "rollout_env_steps_per_sec": o["env_steps"] * world / (o["collect_ms"] * 1e-3), "value_loss": o["value_loss"], "action_loss": o["action_loss"],
                    "entropy": o["entropy"], "games_finished_in_rollout": o["games_finished"], "warmup_s": warm_s,
                    "grad_allreduce": None if world == 1 else "flat fp32 bucket of %d bytes, one NCCL all-reduce (AVG) per optimiser step; advantage statistics: 24 B per epoch" % (tr.flat_grad.numel() * 4)},
            "clocks": clock_info,
            "gpu_launches": None,
            "e2e": None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.config != "env" and args.impl == "b200":
        return run_ppo_arm(args)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
