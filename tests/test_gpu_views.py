"""VecCatanEnv's policy-facing views: obs_views / obs_float / mask_views slice the packed rows into the reference's
keys and shapes (env/wrapper.py:60-83, :172-185; RL/models/policy.py:168-190)."""
import numpy as np
import pytest
import torch

from tests.common import load_golden
from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


def test_views_match_reference_keys_shapes_and_values():
    from settlers_of_catan_rl_b200 import VecCatanEnv
    g = load_golden("default_s1")
    n = 3
    v = VecCatanEnv(n, seed=int(g["seed"]), first_env_id=int(g["env_id"]), auto_reset=0)
    v.reset()
    a = torch.zeros((n, L.ACTION_WORDS), dtype=torch.int32, device=v.device)
    mask = torch.tensor([1, 0, 0], dtype=torch.uint8, device=v.device)        # only env 0 replays the golden game
    for t in range(400):
        a[0] = torch.from_numpy(g["actions"][t]).to(v.device)
        v.step(a, step_mask=mask)
    want = g["obs"][400]
    views = v.obs_views()
    fl = v.obs_float()
    assert views["tile_representations"].shape == (n, 19, 60) and views["current_player_main"].shape == (n, 152)
    assert views["next_next_next_player_main"].shape == (n, 159) and views["current_player_hidden_dev"].shape == (n, 25)
    assert fl["tile_representations"].dtype == torch.float32 and fl["current_player_played_dev"].dtype == torch.int64
    ratio = dict(L.OBS_RATIO_COLUMNS)
    for key, off, shape in L.OBS_NUMERIC:
        k = int(np.prod(shape))
        assert np.array_equal(views[key][0].reshape(-1).cpu().numpy(), want[off:off + k]), key
        ref = want[off:off + k].astype(np.float64)
        for i in range(k):
            ref[i] /= ratio.get(off + i, 1.0)
        assert np.array_equal(fl[key][0].reshape(-1).double().cpu().numpy(), ref), key
    for key, li in L.OBS_LISTS:
        s = L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD
        assert np.array_equal(views[key][0].cpu().numpy(), want[s:s + L.OBS_DEV_PAD]), key
    assert int(views["player_id"][0]) == int(want[L.OBS_META])
    heads = v.mask_views()
    assert [tuple(h.shape[1:]) for h in heads] == [tuple(s) for _, s in L.MASK_HEADS]
    flat = torch.cat([h[0].reshape(-1) for h in heads]).cpu().numpy()
    assert np.array_equal(flat, g["masks"][400][:L.MASK_ENTRIES])
    # the frozen envs still show their reset situation
    st = v.export_state()
    assert st[1, L.STATE_DTYPE.fields["turn"][1] // 2] == 0


def test_step_timing_hooks():
    """catan_set_timing / catan_read_timing: device time of the two kernels on the caller's stream, and that timing does
    not change what a step computes"""
    import numpy as np
    from settlers_of_catan_rl_b200 import VecCatanEnv
    a = VecCatanEnv(2048, seed=2)
    b = VecCatanEnv(2048, seed=2)
    for v in (a, b):
        v.reset()
    acts_a, acts_b = a.sample_random(), b.sample_random()
    a.set_timing(True)
    for _ in range(100):                      # more than the ring of 32 timed steps
        a.step_sample(acts_a)
        b.step_sample(acts_b)
    n, t_ms, e_ms = a.read_timing()
    a.set_timing(False)
    assert n == 100 and 0.0 < t_ms < 5.0 and 0.0 < e_ms < 5.0
    assert torch.equal(a.obs, b.obs) and torch.equal(a.masks, b.masks)
    assert np.array_equal(a.export_state(), b.export_state())
    assert a.read_timing()[0] == 0


def test_search_diagnostics_add_up():
    """catan_read_lr_stats / catan_read_lr_histograms / the per-stream timing of catan_read_timing: every search is counted once
    in each of the two per-search histograms, a step contributes at most one entry to the longest-search histogram, and the
    rows launch is part of the encode time"""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    env = VecCatanEnv(4096, seed=4)
    env.reset()
    acts = env.sample_random()
    steps = 600
    for _ in range(steps):
        env.step_sample(acts)
    st, h = env.lr_stats(), env.lr_histograms()
    assert st[0] > 0 and 0 < st[1] <= st[0]                     # updates, of which searched by a block
    assert int(h[0].sum()) == int(st[1]) and int(h[1].sum()) == int(st[1])
    assert 0 < int(h[2].sum()) <= steps
    env.set_timing(True)
    for _ in range(40):
        env.step_sample(acts)
    n, t_ms, e_ms, rows_ms = env.read_timing(detail=True)
    env.set_timing(False)
    assert n == 40 and 0.0 < rows_ms < e_ms
    assert set(env.stream_timing) == {"search_wait", "search", "search_rest", "reset_stream", "after_transition"}
    assert env.stream_timing["after_transition"] >= e_ms - 1e-6


def test_a_step_can_be_captured_in_a_cuda_graph():
    """the six launches of a step (two streams, forked and joined with events) are capturable: replaying the graph advances
    the games exactly like direct calls"""
    import numpy as np
    from settlers_of_catan_rl_b200 import VecCatanEnv
    a = VecCatanEnv(1024, seed=9)
    b = VecCatanEnv(1024, seed=9)
    for v in (a, b):
        v.reset()
    acts_a, acts_b = a.sample_random(), b.sample_random()
    for _ in range(3):
        a.step_sample(acts_a)
        b.step_sample(acts_b)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        a.step_sample(acts_a)
    for _ in range(300):
        g.replay()
        b.step_sample(acts_b)
    torch.cuda.synchronize()
    assert torch.equal(a.obs, b.obs) and torch.equal(a.masks, b.masks) and torch.equal(acts_a, acts_b)
    assert np.array_equal(a.export_state(), b.export_state())
    assert int(a.lr_stats()[0]) == int(b.lr_stats()[0]) > 0


def test_async_host_step_equals_the_synchronous_one():
    """catan_step_host_async on a side stream + a stream synchronise == catan_step_host (same games, same host buffers)"""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    n = 700
    a, b = VecCatanEnv(n, seed=31), VecCatanEnv(n, seed=31)
    a.reset(); b.reset()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    bufs = [{"act": pin((n, L.ACTION_WORDS), torch.int32), "obs": pin((n, L.OBS_STRIDE), torch.uint8), "masks": pin((n, L.MASK_STRIDE), torch.uint8),
             "rew": pin((n, 4), torch.float32), "info": pin((n, L.INFO_STRIDE), torch.uint8)} for _ in range(2)]
    side = torch.cuda.Stream()
    for tick in range(120):
        acts = a.sample_random()
        torch.cuda.synchronize()
        for k in range(2):
            bufs[k]["act"].copy_(acts)
        a.step_host(*(bufs[0][k].numpy() for k in ("act", "obs", "masks", "rew", "info")))
        with torch.cuda.stream(side):
            b.step_host_async(*(bufs[1][k].numpy() for k in ("act", "obs", "masks", "rew", "info")))
        side.synchronize()
        for k in ("obs", "masks", "rew", "info"):
            assert torch.equal(bufs[0][k], bufs[1][k]), (tick, k)
    assert np.array_equal(a.export_state(), b.export_state())


def test_async_host_step_with_the_fused_sampler():
    """catan_step_sample_host_async == catan_step_host followed by catan_sample_random (actions, reward, info, state)"""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    n = 500
    a, b = VecCatanEnv(n, seed=33), VecCatanEnv(n, seed=33)
    a.reset(); b.reset()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    ha, hb = pin((n, L.ACTION_WORDS), torch.int32), pin((n, L.ACTION_WORDS), torch.int32)
    ra, rb, ia, ib = pin((n, 4), torch.float32), pin((n, 4), torch.float32), pin((n, L.INFO_STRIDE), torch.uint8), pin((n, L.INFO_STRIDE), torch.uint8)
    first, first_b = a.sample_random(), b.sample_random()      # both envs draw decision 0 of the sampler stream
    torch.cuda.synchronize()
    assert torch.equal(first, first_b)
    ha.copy_(first); hb.copy_(first_b)
    side = torch.cuda.Stream()
    for tick in range(150):
        a.step_host(ha.numpy(), None, None, ra.numpy(), ia.numpy())
        nxt = a.sample_random()
        torch.cuda.synchronize()
        ha.copy_(nxt)
        with torch.cuda.stream(side):
            b.step_sample_host_async(hb.numpy(), rb.numpy(), ib.numpy())
        side.synchronize()
        assert torch.equal(ha, hb) and torch.equal(ra, rb) and torch.equal(ia, ib), tick
    assert np.array_equal(a.export_state(), b.export_state())


def test_cuda_randomise_uncertainty_matches_the_game_core_and_conserves():
    """catan_randomise_uncertainty (game.py:1207-1282) on the device == the same game logic compiled for the host (which
    tests/test_randomise_uncertainty.py pins against the reference), game by game incl. the draw counters; and at scale the
    re-dealt games keep 19 cards per resource and the controlling player's hand and hidden cards."""
    import numpy as np
    from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L
    from tests.host_emu.emu_lib import EmuEnv
    n = 4096
    env = VecCatanEnv(n, seed=12, first_env_id=77)
    env.reset()
    a = env.sample_random()
    for _ in range(700):
        env.step_sample(a)
    before = env.export_state()
    ctrl = torch.from_numpy((np.arange(n) % 5).astype(np.uint8)).cuda()          # 0 = untouched
    env.err_flags(clear=True)
    env.randomise_uncertainty(ctrl, max_attempts=300)
    after = env.export_state()
    flags = env.err_flags()
    sb, sa = before.view(L.STATE_DTYPE)[:, 0], after.view(L.STATE_DTYPE)[:, 0]
    c = ctrl.cpu().numpy()
    ok = (flags >> 9) & 1 == 0
    assert ok.mean() > 0.5
    untouched = c == 0
    assert np.array_equal(before[untouched], after[untouched])
    dealt = ok & ~untouched & (sb["initial_phase"] == 0)
    assert np.array_equal((sa["res"].sum(axis=1) + sa["bank"])[dealt], np.full((int(dealt.sum()), 5), 19))
    # (hand SIZES are kept only when the controlling player's minimum beliefs do not overestimate a hand -- the reference has no
    # such guarantee either, its closing assert is the per-resource sum, game.py:1276-1282)
    # (measured on this seed: 75 % of the hands keep their size; the exact statement is the game-by-game comparison below)
    assert (sa["res"].sum(axis=2)[dealt] == sb["res"].sum(axis=2)[dealt]).mean() > 0.5
    idx = np.nonzero(dealt)[0]
    assert np.array_equal(sa["res"][idx, c[idx] - 1], sb["res"][idx, c[idx] - 1])
    assert np.array_equal(sa["hidden"][idx, c[idx] - 1], sb["hidden"][idx, c[idx] - 1])
    assert np.array_equal(sa["n_hidden"], sb["n_hidden"]) and np.array_equal(sa["deck_n"], sb["deck_n"])
    # game by game against the host build of the same logic
    for e in list(idx[:40]) + list(np.nonzero(~ok & ~untouched)[0][:5]):
        emu = EmuEnv(seed=12, env_id=77 + int(e), auto_reset=0)
        emu.import_state(before[e])
        emu.randomise_uncertainty(int(c[e]), 300)
        from tests.common import state_diff
        assert np.array_equal(emu.state(), after[e]), (e, int(c[e]), state_diff(emu.state(), after[e])[:8])
    # the bound rows were refreshed
    o = env.obs.cpu().numpy()
    e = int(idx[0])
    emu = EmuEnv(seed=12, env_id=77 + e, auto_reset=0)
    emu.import_state(after[e])
    assert np.array_equal(emu.obs(), o[e])


def test_library_graph_replay_is_the_same_step():
    """catan_set_graphs: the captured-and-replayed step (device-resident and pinned-host variants) == the directly issued one"""
    import numpy as np
    from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L
    n = 3000
    a_env, b_env = VecCatanEnv(n, seed=5), VecCatanEnv(n, seed=5)
    b_env.set_graphs(True)
    for e in (a_env, b_env):
        e.reset()
    acts_a, acts_b = a_env.sample_random(), b_env.sample_random()
    for t in range(300):
        a_env.step_sample(acts_a)
        b_env.step_sample(acts_b)
    assert torch.equal(a_env.obs, b_env.obs) and torch.equal(a_env.masks, b_env.masks) and torch.equal(acts_a, acts_b)
    assert np.array_equal(a_env.export_state(), b_env.export_state())
    # masked step with explicit actions, and a config change (drops the captured graphs)
    mask = (torch.arange(n, device="cuda") % 3 != 0).to(torch.uint8)
    b_env.set_reward_annealing_factor(0.5); a_env.set_reward_annealing_factor(0.5)
    for t in range(20):
        a_env.step(acts_a, step_mask=mask); a_env.sample_random(acts_a)
        b_env.step(acts_b, step_mask=mask); b_env.sample_random(acts_b)
    assert torch.equal(a_env.obs, b_env.obs) and torch.equal(a_env.reward, b_env.reward)
    # pinned host buffers: H2D actions, step + sampler, D2H actions / reward / info in one replayed graph
    h_act = [torch.empty((n, L.ACTION_WORDS), dtype=torch.int32).pin_memory() for _ in range(2)]
    h_rew = [torch.empty((n, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_info = [torch.empty((n, L.INFO_STRIDE), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for k, (e, acts) in enumerate(((a_env, acts_a), (b_env, acts_b))):
        h_act[k].copy_(acts)
    for t in range(30):
        for k, e in enumerate((a_env, b_env)):
            e.step_sample_host_async(h_act[k].numpy(), h_rew[k].numpy(), h_info[k].numpy())
        torch.cuda.synchronize()
        assert torch.equal(h_act[0], h_act[1]) and torch.equal(h_rew[0], h_rew[1]) and torch.equal(h_info[0], h_info[1]), t
    assert int(a_env.err_flags().any()) == 0 and int(b_env.err_flags().any()) == 0


@pytest.mark.parametrize("packed", [False, True])
def test_host_env_groups_pump_is_the_per_handle_loop(packed):
    """catan_step_sample_host_groups (one library call per round over several handles in flight) == the same handles stepped one
    by one with catan_step_sample_host_async: actions, reward, info rows, final state; the done flags it read add up.  packed: the
    groups move their action rows as one byte per word (catan_step_sample_host_async_u8), the per-handle loop as int32"""
    import numpy as np
    from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L
    from settlers_of_catan_rl_b200.vec_env import HostEnvGroups
    sizes, seed = [700, 650, 697], 9
    def make():
        envs = [VecCatanEnv(m, seed=seed, first_env_id=sum(sizes[:k])) for k, m in enumerate(sizes)]
        for e in envs:
            e.set_graphs(True)
            e.reset()
        return envs
    a_envs, b_envs = make(), make()
    groups = HostEnvGroups(a_envs, packed_actions=packed)
    groups.prime()
    h = [(torch.empty((m, L.ACTION_WORDS), dtype=torch.int32).pin_memory(), torch.empty((m, 4), dtype=torch.float32).pin_memory(),
          torch.zeros((m, L.INFO_STRIDE), dtype=torch.uint8).pin_memory()) for m in sizes]
    for e, (ha, _, _) in zip(b_envs, h):
        ha.copy_(e.sample_random())
    torch.cuda.synchronize()
    done_b = 0
    for r in range(12):
        groups.pump(50)
        for _ in range(50):
            for e, (ha, hr, hi) in zip(b_envs, h):
                done_b += int(hi[:, L.INFO_DONE].sum())                 # (the previous step's rows, as the pump reads them)
                e.step_sample_host_async(ha.numpy(), hr.numpy(), hi.numpy())
                torch.cuda.synchronize()
        groups.synchronize()
        for k in range(len(sizes)):
            want = h[k][0]
            if packed:
                assert int(want.max()) < 255
                want = torch.where(want < 0, torch.full_like(want, 255), want).to(torch.uint8)
            assert torch.equal(groups.actions[k], want) and torch.equal(groups.reward[k], h[k][1]) and torch.equal(groups.info[k], h[k][2]), (r, k)
    assert groups.done_seen.value == done_b and done_b > 0
    for ea, eb in zip(a_envs, b_envs):
        assert np.array_equal(ea.export_state(), eb.export_state())
        assert not ea.err_flags().any()
