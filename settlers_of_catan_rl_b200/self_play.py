"""Per-seat policies for vectorised self-play.

The reference's ``GamesAndPoliciesManager`` (RL/ppo/game_manager.py:10-31) holds four policies — the learner and three
earlier snapshots (RL/ppo/update_opponent_policies.py:13-26) — shuffles them over the four seats of every env and, env by
env, lets the policy of the player whose decision it is act (:82-93); only policy 0's seat is recorded (:26, :94-133).
``SeatPolicies`` keeps the same maps for N lock-step envs on the device: per tick one routing launch
(``catan_route_by_policy``) gives every policy the list of envs it acts for, one ``catan_policy_inputs`` launch per
policy reads those envs' packed rows straight into its input tensors, each policy runs ONE batched forward, and the
sampled heads are scattered back into the int32 action rows the env steps on.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import layout as L
from .policy_io import PolicyInputs, actions_to_rows

#: ``policy(obs_dict, action_masks) -> (heads, log_probs)``: the 12 sampled heads and the joint log-probs ``[B, 1]``
PolicyFn = Callable[[dict, list], Tuple[Sequence, torch.Tensor]]


def reference_policy_fn(policy, deterministic: bool = False) -> PolicyFn:
    """adapter for the reference's ``SettlersAgentPolicy`` (RL/models/policy.py:70-92): no LSTM state, no terminal masks"""
    def fn(obs, masks):
        with torch.no_grad():
            _, heads, logp, _ = policy.act(obs, None, None, masks, deterministic=deterministic)
        return heads, logp
    return fn


class SeatPolicies:
    def __init__(self, env, policies: Sequence[PolicyFn], generator: Optional[torch.Generator] = None,
                 dtype: torch.dtype = torch.float32, order: Optional[torch.Tensor] = None):
        """``order`` (optional, [N, 4], a permutation of 0..3 per env): ``order[n, j]`` = PlayerId - 1 of the seat policy j
        plays in env n; default = a random shuffle per env (game_manager.py:24-25)."""
        assert len(policies) == 4, "four seats, four policies (game_manager.py:15)"
        self.env, self.policies = env, list(policies)
        N, dev = env.n_envs, env.device
        if order is None:
            order = torch.rand((N, 4), device=dev, generator=generator).argsort(dim=1)
        order = order.to(device=dev, dtype=torch.int64)
        assert order.shape == (N, 4) and bool((order.sort(dim=1).values == torch.arange(4, device=dev)).all())
        pm = torch.empty((N, 4), dtype=torch.int64, device=dev)
        pm.scatter_(1, order, torch.arange(4, device=dev).expand(N, 4))        # policy_map[order[j]] = policies[j]  (:28-30)
        self.policy_map = pm.to(torch.uint8).contiguous()
        self.active_pid = (order[:, 0] + 1).to(torch.uint8).contiguous()       # active player controlled by policy[0]  (:26)
        self.inputs = PolicyInputs(N, dev, dtype)
        self.actions = torch.zeros((N, L.ACTION_WORDS), dtype=torch.int32, device=dev)
        self.logp = torch.zeros(N, dtype=torch.float32, device=dev)
        self.active: Optional[torch.Tensor] = None     # uint8 [N]: envs still collecting (RolloutStorage.collecting); None = all
        self.last_counts: List[int] = [0, 0, 0, 0]

    def act(self, env=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """one decision for every (active) env by the policy that owns it: (actions int32 [N, 20], log-probs fp32 [N])"""
        env = self.env if env is None else env
        counts, lists = env.route_by_policy(self.policy_map, 4, self.active)
        self.last_counts = counts.tolist()             # the batch sizes are needed on the host: one small D2H per tick
        for k, c in enumerate(self.last_counts):
            if c == 0:
                continue
            idx = lists[k, :c]
            obs, masks = self.inputs(env.obs, env.masks, index=idx)
            heads, logp = self.policies[k](obs, masks)
            rows = idx.long()
            self.actions.index_copy_(0, rows, actions_to_rows(heads))
            self.logp.index_copy_(0, rows, logp.reshape(-1).float())
        return self.actions, self.logp

    __call__ = act


def league_probabilities(num_policies: int, linear_num: int = 800, linear_prob: float = 0.5):
    """Sampling weights over the stored earlier policies, oldest first (RL/ppo/update_opponent_policies.py:29-43
    ``get_prob_dist``): half of the mass uniform, half rising linearly over the ``linear_num`` most recent snapshots."""
    import numpy as np
    p = np.full(num_policies, (1.0 - linear_prob) / num_policies, dtype=np.float64)
    num_aux = min(linear_num, num_policies)
    grad = (2.0 * linear_prob) / (num_aux + 1) / num_aux
    p[num_policies - num_aux:] += np.arange(num_aux, dtype=np.float64) * grad
    return p / p.sum()


def sample_opponents(earlier_policies: Sequence, rng=None, count: int = 3) -> list:
    """three earlier snapshots for seats 1-3 of a worker (update_opponent_policies.py:13-26): drawn with replacement from
    ``league_probabilities``; seat 0 stays the learner.  ``rng``: a ``numpy.random.Generator`` / ``RandomState`` (default: the
    global numpy state, as in the reference)."""
    import numpy as np
    p = league_probabilities(len(earlier_policies))
    idx = (np.random if rng is None else rng).choice(len(earlier_policies), count, p=p)
    return [earlier_policies[int(i)] for i in idx]
