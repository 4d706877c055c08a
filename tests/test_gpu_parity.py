"""GPU parity tests proper: the CUDA path, called through the C ABI, against
(a) the golden fixtures generated from the real reference, step by step (bit-exact state / obs / masks /
    reward / done), and
(b) the C oracle on thousands of games driven by the pinned sampler (size-independent properties too)."""
import numpy as np
import pytest
import torch

from tests.common import golden_cases, load_golden, replay_golden, state_diff, row_diff
from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


class _GpuAdapter:
    """EnvWrapper-shaped view of a 1-env VecCatanEnv (auto_reset off), every call through the C ABI."""

    def __init__(self, g, n_pad=0):
        from settlers_of_catan_rl_b200 import VecCatanEnv
        self.v = VecCatanEnv(1, seed=int(g["seed"]), first_env_id=int(g["env_id"]), auto_reset=0, **g["cfg"])
        self.actions = torch.zeros((1, L.ACTION_WORDS), dtype=torch.int32, device=self.v.device)

    def reset(self):
        self.v.reset()

    def step(self, a):
        self.actions.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).view(1, -1))
        self.v.step(self.actions)
        info = self.v.info[0].cpu().numpy()
        return int(info[L.INFO_ERR]), self.v.reward[0].cpu().numpy(), info

    def state(self):
        return self.v.export_state()[0]

    def obs(self):
        return self.v.obs[0].cpu().numpy()

    def masks(self):
        return self.v.masks[0].cpu().numpy()


@pytest.mark.parametrize("case", golden_cases())
def test_cuda_replays_reference_golden(case):
    g = load_golden(case)
    assert replay_golden(_GpuAdapter(g), g) == len(g["actions"])


def test_cuda_host_buffer_path_replays_golden():
    """catan_step_host / catan_reset_host: the call an EnvWrapper-shaped adapter makes (host buffers in and out)."""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    g = load_golden("default_s0")
    v = VecCatanEnv(1, seed=int(g["seed"]), first_env_id=int(g["env_id"]), auto_reset=0)
    obs = np.zeros((1, L.OBS_STRIDE), np.uint8)
    masks = np.zeros((1, L.MASK_STRIDE), np.uint8)
    reward = np.zeros((1, 4), np.float32)
    info = np.zeros((1, L.INFO_STRIDE), np.uint8)
    v.reset_host(obs, masks, info)
    assert np.array_equal(obs[0], g["obs"][0]) and np.array_equal(masks[0], g["masks"][0])
    for t in range(len(g["actions"])):
        v.step_host(np.ascontiguousarray(g["actions"][t:t + 1]), obs, masks, reward, info)
        assert np.array_equal(obs[0], g["obs"][t + 1]), t
        assert np.array_equal(masks[0], g["masks"][t + 1]), t
        assert np.array_equal(reward[0], g["reward"][t]) and info[0, L.INFO_DONE] == g["done"][t]


def test_cuda_auto_reset_matches_reference_reset():
    """with auto_reset the step that ends game 1 returns the obs/masks of the reference's next reset()"""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    g = load_golden("default_s5_two_games")
    v = VecCatanEnv(1, seed=int(g["seed"]), first_env_id=int(g["env_id"]), auto_reset=1)
    v.reset()
    a = torch.zeros((1, L.ACTION_WORDS), dtype=torch.int32, device=v.device)
    resets = {int(t): i for i, t in enumerate(g["reset_at"])}
    for t in range(len(g["actions"])):
        a.copy_(torch.from_numpy(g["actions"][t]).view(1, -1))
        v.step(a)
        info = v.info[0].cpu().numpy()
        assert info[L.INFO_DONE] == g["done"][t]
        assert np.array_equal(v.reward[0].cpu().numpy(), g["reward"][t])
        if (t + 1) in resets:
            i = resets[t + 1]
            assert info[L.INFO_RESET] == 1
            assert np.array_equal(v.obs[0].cpu().numpy(), g["reset_obs"][i])
            assert np.array_equal(v.masks[0].cpu().numpy(), g["reset_masks"][i])
            assert not state_diff(v.export_state()[0], g["reset_state"][i])
        elif not g["done"][t]:
            assert np.array_equal(v.obs[0].cpu().numpy(), g["obs"][t + 1]), t


@pytest.mark.parametrize("n_envs,steps,chunk,cfg", [
    (4096, 1500, 100, {}),
    (1000, 600, 50, dict(dense_reward=1, reward_annealing_factor=0.5, max_proposed_trades_per_turn=-1)),
    (333, 400, 57, dict(max_actions_per_turn=4)),
    # soak: the BASELINE size (65 536 envs) through 2 200 ticks, and 4 096 envs through 10 000 ticks (about eight games per env: long
    # games, u16 turn counters, full card lists, every reset path), with unlimited trade proposals in the long one
    (65536, 2200, 200, {}),
    (4096, 10000, 500, dict(max_proposed_trades_per_turn=-1)),
])
def test_cuda_matches_oracle_at_scale(n_envs, steps, chunk, cfg):
    """same (seed, env ids), same pinned sampler: kernel (fused step+sample) vs the C oracle on host threads"""
    from oracle import oracle_lib as O
    from settlers_of_catan_rl_b200 import VecCatanEnv
    seed, first = 7, 10_000
    ov = O.OracleVec(n_envs, seed=seed, first_env_id=first, **cfg)
    ov.run(0)
    v = VecCatanEnv(n_envs, seed=seed, first_env_id=first, **cfg)
    v.reset()
    acts = v.sample_random()
    done_total = torch.zeros(n_envs, dtype=torch.int64, device=v.device)
    reward_total = torch.zeros((n_envs, 4), dtype=torch.float64, device=v.device)
    for t0 in range(0, steps, chunk):
        k = min(chunk, steps - t0)
        for _ in range(k):
            v.step_sample(acts)
            done_total += v.info[:, L.INFO_DONE].long()
            reward_total += v.reward.double()
        ov.run(k)
        st = v.export_state()
        bad = np.nonzero((st[:, :-2] != ov.states[:, :-2]).any(axis=1))[0]
        assert bad.size == 0, (t0, bad[:5], state_diff(ov.states[bad[0]], st[bad[0]])[:5])
        assert np.array_equal(st[:, -2:], ov.states[:, -2:]), "game RNG draw counters diverged"
        assert np.array_equal(v.obs.cpu().numpy(), ov.obs), t0
        assert np.array_equal(v.masks.cpu().numpy(), ov.masks), t0
        assert np.array_equal(v.info.cpu().numpy(), ov.info), t0
    assert np.array_equal(done_total.cpu().numpy(), ov.games_done)
    np.testing.assert_allclose(reward_total.cpu().numpy(), ov.reward_sum.astype(np.float64), rtol=0, atol=1e-2)
    assert not v.err_flags().any()
    if steps >= 1000:
        assert ov.games_done.sum() > 0


def test_cuda_imported_games_take_the_full_longest_road_search():
    """catan_import_state cannot know how cur_longest_path relates to the imported board, so the first road of every
    player re-enumerates all paths (game.py:843-862 as written) instead of the incremental rule; both must agree."""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    n = 2048
    a = VecCatanEnv(n, seed=11, first_env_id=500)
    a.reset()
    acts = a.sample_random()
    for _ in range(700):
        a.step_sample(acts)
    b = VecCatanEnv(n, seed=11, first_env_id=500)
    b.reset()
    b.import_state(a.export_state())
    assert torch.equal(a.obs, b.obs) and torch.equal(a.masks, b.masks)
    full0 = int(b.lr_stats()[2])
    for t in range(400):
        cur = acts.clone()                             # the actions `a` is about to apply
        a.step_sample(acts)
        b.step(cur)
        if t % 50 == 49:
            assert np.array_equal(a.export_state(), b.export_state()), t
            assert torch.equal(a.obs, b.obs) and torch.equal(a.masks, b.masks) and torch.equal(a.reward, b.reward), t
    sa, sb = a.lr_stats(), b.lr_stats()
    assert sb[2] - full0 > 100, "the imported games never took the full enumeration"
    assert sa[0] > 0 and sa[1] <= sa[0] and sa[2] <= sa[1]
    assert not a.err_flags().any() and not b.err_flags().any()


def test_cuda_split_sample_then_step_equals_fused():
    from settlers_of_catan_rl_b200 import VecCatanEnv
    a = VecCatanEnv(512, seed=3, first_env_id=99)
    b = VecCatanEnv(512, seed=3, first_env_id=99)
    a.reset()
    b.reset()
    acts_a = a.sample_random()
    for _ in range(300):
        a.step_sample(acts_a)
        acts_b = b.sample_random()
        b.step(acts_b)
    assert torch.equal(a.obs, b.obs) and torch.equal(a.masks, b.masks) and torch.equal(a.info, b.info)
    assert np.array_equal(a.export_state(), b.export_state())


def test_cuda_results_do_not_depend_on_sharding():
    """games are keyed by global env id: 2 shards of 300 == one engine of 600"""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    whole = VecCatanEnv(600, seed=11, first_env_id=0)
    parts = [VecCatanEnv(300, seed=11, first_env_id=0), VecCatanEnv(300, seed=11, first_env_id=300)]
    for v in [whole] + parts:
        v.reset()
        v._acts = v.sample_random()
    for _ in range(400):
        for v in [whole] + parts:
            v.step_sample(v._acts)
    assert torch.equal(whole.obs, torch.cat([p.obs for p in parts]))
    assert np.array_equal(whole.export_state(), np.concatenate([p.export_state() for p in parts]))


def test_cuda_full_size_invariants():
    """BASELINE config 2 size (65 536 envs): size-independent properties after 300 random-legal ticks —
    resource conservation (game/utils.py:18-26 with res_tot=19), VP identity, masks non-empty, no rejected action,
    export->import->export idempotence."""
    from settlers_of_catan_rl_b200 import VecCatanEnv
    n = 65536
    v = VecCatanEnv(n, seed=2024)
    v.reset()
    acts = v.sample_random()
    for _ in range(300):
        v.step_sample(acts)
    assert not v.err_flags().any()
    stv = v.export_state()
    st = stv.view(L.STATE_DTYPE).reshape(-1)
    assert np.all(st["bank"] + st["res"].sum(axis=1) == 19)
    settlements = np.stack([((st["corner_type"] == 1) & (st["corner_owner"] == p + 1)).sum(axis=1) for p in range(4)], 1)
    cities = np.stack([((st["corner_type"] == 2) & (st["corner_owner"] == p + 1)).sum(axis=1) for p in range(4)], 1)
    played_vp = np.stack([((st["played"][:, p, :] == 1) & (np.arange(25)[None, :] < st["n_played"][:, p, None])).sum(axis=1)
                          for p in range(4)], 1)
    lr = np.stack([(st["lr_holder"] == p + 1) for p in range(4)], 1)
    la = np.stack([(st["la_holder"] == p + 1) for p in range(4)], 1)
    assert np.array_equal(st["vp"], settlements + 2 * cities + played_vp + 2 * lr + 2 * la)
    assert (v.masks[:, :13].sum(dim=1) > 0).all()
    obs_before, masks_before = v.obs.clone(), v.masks.clone()
    v.import_state(stv[:1024], first=0)
    assert np.array_equal(v.export_state(0, 1024), stv[:1024])
    assert torch.equal(v.obs, obs_before) and torch.equal(v.masks, masks_before)


def test_cuda_rejects_illegal_actions_without_touching_state():
    from settlers_of_catan_rl_b200 import VecCatanEnv
    v = VecCatanEnv(8, seed=1)
    v.reset()
    before = v.export_state()
    a = torch.zeros((8, L.ACTION_WORDS), dtype=torch.int32, device=v.device)
    a[:, 0] = torch.tensor([9, 10, 13, 12, 3, 7, 11, 8], dtype=torch.int32)
    v.step(a)
    assert (v.info[:, L.INFO_ERR] != 0).all()
    assert np.array_equal(before, v.export_state())
    assert (v.err_flags(clear=True) != 0).all() and not v.err_flags().any()
