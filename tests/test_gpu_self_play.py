"""Per-seat policy routing end to end (game_manager.py:21-31, :82-93): four policies shuffled over the seats of every env,
each acting — in one batched call per tick — for exactly the envs whose decision belongs to it, through the indexed
catan_policy_inputs launch; only policy 0's seat is recorded by the rollout collector."""
import pytest
import torch

from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


def toy_policy(variant: int):
    """a deterministic legal policy written against the policy-facing tensors only: the first (even variant) or last (odd)
    legal option of every head that the chosen action type reads; log-prob tag = -(variant + 1)"""
    def pick(m):
        m = m.float()
        ar = torch.arange(m.shape[1], device=m.device, dtype=torch.float32)
        return ((m * (m.shape[1] - ar)) if variant % 2 == 0 else (m * (ar + 1))).argmax(dim=1)

    def fn(obs, masks):
        t = pick(masks[0])
        z = torch.zeros_like(t)
        card = torch.where(t == 4, pick(masks[4]), z)
        yop, mono = (t == 4) & (card == 2), (t == 4) & (card == 4)
        corner = torch.where(t == 0, pick(masks[1][0]), torch.where(t == 2, pick(masks[1][1]), z))
        player = torch.where(t == 6, pick(masks[6][0]), torch.where(t == 11, pick(masks[6][1]), z))
        res_a = torch.where(t == 5, pick(masks[9][0]), torch.where(mono, pick(masks[9][2]), torch.where(yop, pick(masks[9][3]), z)))
        res_b = torch.where((t == 5) | yop, pick(masks[10]), z)
        give = torch.where(t == 6, 1 + pick(obs["current_resources"][:, 1:6] > 0), z)
        recv = torch.where(t == 6, torch.full_like(t, 1 + variant), z)
        col = lambda x: x.view(-1, 1)
        heads = [col(t), col(corner), col(torch.where(t == 1, pick(masks[2]), z)), col(torch.where(t == 8, pick(masks[3]), z)),
                 col(card), col(torch.where(t == 7, pick(masks[5]), z)), col(player), [col(give), col(z), col(z), col(z)],
                 [col(recv), col(z), col(z), col(z)], col(res_a), col(res_b), col(torch.where(t == 12, pick(masks[11]), z))]
        return heads, torch.full((t.shape[0], 1), -float(variant + 1), device=t.device)
    return fn


def test_every_env_is_played_by_the_policy_its_map_names():
    from settlers_of_catan_rl_b200 import VecCatanEnv, SeatPolicies, PolicyInputs
    from settlers_of_catan_rl_b200.policy_io import actions_to_rows
    n = 2500
    env = VecCatanEnv(n, seed=21)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    seats = SeatPolicies(env, [toy_policy(k) for k in range(4)], generator=g)
    assert bool((seats.policy_map.long().gather(1, (seats.active_pid.long() - 1).view(-1, 1)) == 0).all())
    full = PolicyInputs(n, env.device)
    seen = torch.zeros(4, dtype=torch.int64)
    for tick in range(500):
        actor = env.info[:, L.INFO_ACTOR].long()
        owner = seats.policy_map.long().gather(1, (actor - 1).view(-1, 1)).view(-1)
        # what every policy would do for every env, from the un-indexed batch (before act() reuses nothing of it)
        obs, masks = full(env.obs, env.masks)
        want = torch.stack([actions_to_rows(toy_policy(k)(obs, masks)[0]) for k in range(4)])          # [4, n, 20]
        actions, logp = seats.act()
        assert sum(seats.last_counts) == n
        assert torch.equal(logp, -(owner + 1).float()), tick
        assert torch.equal(actions, want[owner, torch.arange(n, device=env.device)]), tick
        seen += torch.tensor(seats.last_counts)
        env.step(actions)
    assert int(env.err_flags().astype(bool).sum()) == 0          # every routed action was legal for the env it went to
    assert bool((seen > 0).all())


def test_collector_records_only_policy_zero():
    from settlers_of_catan_rl_b200 import VecCatanEnv, SeatPolicies, RolloutStorage
    n, T = 600, 12
    env = VecCatanEnv(n, seed=22)
    env.reset()
    seats = SeatPolicies(env, [toy_policy(k) for k in range(4)], generator=torch.Generator(device="cuda").manual_seed(6))
    st = RolloutStorage(env, T, active_pid=seats.active_pid)
    st.begin(True)
    seats.active = st.collecting
    ticks = st.collect(seats.act)
    assert st.finished() and ticks >= T
    assert bool((st.logp == -1.0).all())                          # game_manager.py:26, :94-110: policy 0's decisions only
    assert bool((st.cursors[:, 0] == T + 1).all())
