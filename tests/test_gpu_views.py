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
