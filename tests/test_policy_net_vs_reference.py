"""``CatanPolicy`` (the batched restatement of the reference's policy / value network) against the reference's own
``SettlersAgentPolicy`` (RL/models/*) with the SHIPPED checkpoint ``RL/results/default_after_update_3825.pt``:
same parameter names and count, and for the same observations, masks and actions the same values, joint log-probs and
entropies (1e-5, fp32) — on states of real games incl. every action type, trades with 1-4 cards and empty hands."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness as H
from oracle import oracle_lib as O
from oracle.policy_ref import rows_to_policy_inputs
from settlers_of_catan_rl_b200 import layout as L
from settlers_of_catan_rl_b200.policy_io import actions_to_rows, rows_to_actions
from settlers_of_catan_rl_b200.policy_net import CatanPolicy

#: fp32 tolerance: the two networks are the same function evaluated in a different order (one GEMM for q / k / v and for the
#: twelve heads' trunk columns, fused attention), so sums are re-associated; observed differences are <= 2e-5 on values of O(1)
TOL = 1e-4

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference tree not present")


def _reference_policy():
    H.import_reference()
    from RL.models.build_agent_model import build_agent_model  # type: ignore
    ref = build_agent_model(device="cpu")
    sd = torch.load(os.path.join(H.REFERENCE_ROOT, "RL", "results", "default_after_update_3825.pt"), map_location="cpu", weights_only=False)
    ref.load_state_dict(sd)
    ref.eval()
    return ref, sd


def _game_rows(n_envs=64, ticks=(5, 40, 200, 600, 1100), seed=3):
    ov = O.OracleVec(n_envs, seed=seed, first_env_id=50)
    ov.run(0)
    obs, masks, t0 = [], [], 0
    for t in ticks:
        ov.run(t - t0)
        t0 = t
        obs.append(ov.obs.copy()); masks.append(ov.masks.copy())
    return torch.from_numpy(np.concatenate(obs)), torch.from_numpy(np.concatenate(masks))


def test_same_parameters_as_the_reference_network():
    ref, sd = _reference_policy()
    mine = CatanPolicy()
    mine.load_reference_state_dict(sd)
    n_ref = sum(p.numel() for p in ref.parameters())
    n_mine = sum(p.numel() for p in mine.parameters())
    assert n_ref == n_mine == 1928995
    want = {k for k in sd if not k.endswith("dummy_param")}
    assert set(mine.state_dict().keys()) == want
    ref.load_state_dict(mine.reference_state_dict(), strict=False)     # and back


def test_values_logprobs_entropies_match_the_reference_network():
    ref, sd = _reference_policy()
    mine = CatanPolicy()
    mine.load_reference_state_dict(sd)
    mine.eval()
    obs_rows, mask_rows = _game_rows()
    B = obs_rows.shape[0]
    with torch.no_grad():
        obs, masks = rows_to_policy_inputs(obs_rows, mask_rows)
        # (1) actions sampled by the REFERENCE network, evaluated by both
        torch.manual_seed(0)
        v_ref, acts, lp_ref, _, ent_ref = ref.act(obs, None, None, masks, return_entropy=True)
        rows = actions_to_rows(acts)
        v_ref2, lp_ref2, ent_ref2, _ = ref.evaluate_actions(obs, None, None, rows_to_actions(rows), masks)
        obs2, masks2 = rows_to_policy_inputs(obs_rows, mask_rows)
        v, lp, ent = mine.evaluate_actions(obs2, masks2, rows)
        torch.testing.assert_close(v, v_ref, rtol=TOL, atol=TOL)
        torch.testing.assert_close(lp, lp_ref, rtol=TOL, atol=TOL)
        torch.testing.assert_close(lp, lp_ref2, rtol=TOL, atol=TOL)
        torch.testing.assert_close(ent, ent_ref2, rtol=TOL, atol=TOL)
        torch.testing.assert_close(mine.get_value(obs2), ref.get_value(obs, None, None), rtol=TOL, atol=TOL)
        assert len(set(rows[:, L.A_TYPE].tolist())) >= 6
        # (2) actions sampled by CatanPolicy.act: legal for the env's masks, and the reference assigns them the same log-prob
        g = torch.Generator().manual_seed(5)
        v3, rows3, lp3 = mine.act(obs2, masks2, generator=g)
        _, lp_r3, _, _ = ref.evaluate_actions(obs, None, None, rows_to_actions(rows3), masks)
        assert torch.isfinite(lp3).all()
        torch.testing.assert_close(lp3, lp_r3, rtol=TOL, atol=TOL)
        torch.testing.assert_close(v3, v_ref, rtol=TOL, atol=TOL)
        typ = rows3[:, L.A_TYPE].long()
        assert bool((mask_rows[torch.arange(B), typ] == 1).all())
        # (3) deterministic mode == the reference's mode()
        _, acts_d, lp_d, _ = ref.act(obs, None, None, masks, deterministic=True)
        _, rows_d, lp_dm = mine.act(obs2, masks2, deterministic=True)
        same_type = rows_d[:, 0] == actions_to_rows(acts_d)[:, 0]
        assert same_type.float().mean() > 0.99          # (ties / 1e-7 differences may flip an argmax)
        torch.testing.assert_close(lp_dm[same_type], lp_d[same_type], rtol=1e-4, atol=1e-4)


def test_every_action_type_and_long_trades_are_evaluated_alike():
    """synthetic actions and masks: every type, every card, 1-4 card trade lists"""
    ref, sd = _reference_policy()
    mine = CatanPolicy()
    mine.load_reference_state_dict(sd)
    mine.eval()
    obs_rows, mask_rows = _game_rows(n_envs=48, ticks=(300, 900), seed=9)
    B = obs_rows.shape[0]
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        obs, masks = rows_to_policy_inputs(obs_rows, mask_rows)
        _, base_rows, _ = mine.act(obs, masks, generator=g)
        rows = base_rows.clone().long()
        for b in range(B):
            t = b % 13
            rows[b, L.A_TYPE] = t
            if t == 6:   # propose: give cards the player holds, 1-4 of them, then stop
                hand = obs_rows[b, L.OBS_CURRENT_RES + 1:L.OBS_CURRENT_RES + 6].long()
                give = [r + 1 for r in range(5) for _ in range(int(hand[r]))][: 1 + b % 4]
                if not give:
                    rows[b, L.A_TYPE] = 12
                    continue
                give += [0] * (4 - len(give))
                rows[b, L.A_GIVE:L.A_GIVE + 4] = torch.tensor(give)
                rows[b, L.A_RECV:L.A_RECV + 4] = torch.tensor([1 + (b % 5), (b % 3), 0, 0])
            if t == 4:
                rows[b, L.A_CARD] = b % 5
        # synthetic masks: random bits, with the entry of every chosen sub-action legal in every type row of its head, so that
        # whichever type-conditional row the networks select the log-prob is finite and the row selection matters
        mask_rows = (torch.rand(mask_rows.shape, generator=g) < 0.6).to(torch.uint8)
        cols = {0: L.A_TYPE, 1: L.A_CORNER, 2: L.A_EDGE, 3: L.A_TILE, 4: L.A_CARD, 5: L.A_ACCEPT, 6: L.A_PLAYER, 9: L.A_RES_A,
                10: L.A_RES_B, 11: L.A_DISCARD}
        for h, col in cols.items():
            off, shape = L.MASK_HEADS[h]
            types, dim = (shape[0], shape[1]) if len(shape) == 2 else (1, shape[0])
            for k in range(types):
                mask_rows[torch.arange(B), off + k * dim + rows[:, col]] = 1
        obs_a, masks_a = rows_to_policy_inputs(obs_rows, mask_rows)
        v_r, lp_r, ent_r, _ = ref.evaluate_actions(obs_a, None, None, rows_to_actions(rows), masks_a)
        obs_b, masks_b = rows_to_policy_inputs(obs_rows, mask_rows)
        v, lp, ent = mine.evaluate_actions(obs_b, masks_b, rows)
        assert torch.isfinite(lp_r).all() and len(set(rows[:, L.A_TYPE].tolist())) == 13
        torch.testing.assert_close(lp, lp_r, rtol=TOL, atol=TOL)
        torch.testing.assert_close(ent, ent_r, rtol=TOL, atol=TOL)
        torch.testing.assert_close(v, v_r, rtol=TOL, atol=TOL)
