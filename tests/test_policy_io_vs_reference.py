"""The policy hand-off layout against its real consumer: the reference's policy network (RL/models, random init) fed
(a) with the reference's own per-env conversions of the reference env's obs dict / mask list (policy.py:168-190) and
(b) with one batch decoded from the packed rows (oracle/policy_ref.py, the torch statement the CUDA kernel is checked
against on the GPU) must give the same values, log-probs and entropies.  Runs where /root/reference exists."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_harness as H
from oracle.policy_ref import rows_to_policy_inputs
from settlers_of_catan_rl_b200 import layout as L
from settlers_of_catan_rl_b200.policy_io import actions_to_rows, rows_to_actions   # pure tensor plumbing: runs on CPU tensors too

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference tree not present")


def test_reference_policy_reads_the_packed_rows_like_its_own_obs():
    R = H.import_reference()
    from RL.models.build_agent_model import build_agent_model  # type: ignore
    torch.manual_seed(0)
    policy = build_agent_model(device="cpu")
    policy.eval()
    game_rng, samp = H.PhiloxStream(3, 0, 0), H.PhiloxStream(3, 0, 1)
    env = R["EnvWrapper"]()
    obs_rows, mask_rows, act_rows, values, logps = [], [], [], [], []
    with H.patched_rng(game_rng), torch.no_grad():
        obs = env.reset()
        for t in range(700):
            masks = env.get_action_masks()
            o_row, m_row = H.obs_to_packed(obs), H.masks_to_packed(masks)
            if t % 7 == 0:   # the reference's own path: one env, its own conversions, its own sampling
                v, acts, lp, _ = policy.act(policy.obs_to_torch(copy.deepcopy(obs)), None, None,
                                            policy.act_masks_to_torch(copy.deepcopy(masks)))
                # the sampled heads go back to the env exactly as torch_act_to_np would hand them over (policy.py:192-199)
                back = H.action_to_reference(actions_to_rows(acts)[0].numpy())
                ref_np = policy.torch_act_to_np(copy.deepcopy(acts))
                assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(back, ref_np))
                obs_rows.append(o_row); mask_rows.append(m_row)
                act_rows.append(actions_to_rows(acts)); values.append(v.view(-1)); logps.append(lp.view(-1))
            a = H.sample_action(m_row, o_row, samp.block(t))
            obs, _, done, _ = env.step(H.action_to_reference(a))
            if done:
                obs = env.reset()
        assert len(obs_rows) == 100
        bobs, bmasks = rows_to_policy_inputs(torch.from_numpy(np.stack(obs_rows)), torch.from_numpy(np.stack(mask_rows)))
        actions = rows_to_actions(torch.cat(act_rows))
        v2, lp2, _, _ = policy.evaluate_actions(bobs, None, None, actions, bmasks)
        torch.testing.assert_close(v2.view(-1), torch.cat(values), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(lp2.view(-1), torch.cat(logps), rtol=1e-5, atol=1e-5)
        # a second pass one row at a time: the batch layout ([types, B, dim] masks, padded card lists) adds nothing
        for i in (0, 37, 99):
            o1, m1 = rows_to_policy_inputs(torch.from_numpy(obs_rows[i][None]), torch.from_numpy(mask_rows[i][None]))
            v1, lp1, _, _ = policy.evaluate_actions(o1, None, None, rows_to_actions(act_rows[i]), m1)
            torch.testing.assert_close(v1.view(-1), values[i], rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(lp1.view(-1), logps[i], rtol=1e-5, atol=1e-5)


def test_reference_policy_plays_the_reference_env_through_the_hand_off():
    """the whole loop on the reference's own objects: env obs / masks -> packed rows -> decoded batch -> policy.act -> action
    rows -> env.step, with validate_actions=True (an illegal action raises, wrapper.py:38-41)"""
    R = H.import_reference()
    from RL.models.build_agent_model import build_agent_model  # type: ignore
    torch.manual_seed(1)
    policy = build_agent_model(device="cpu")
    policy.eval()
    env = R["EnvWrapper"]()
    with H.patched_rng(H.PhiloxStream(4, 0, 0)), torch.no_grad():
        obs = env.reset()
        types_seen = set()
        for t in range(400):
            o_row, m_row = H.obs_to_packed(obs), H.masks_to_packed(env.get_action_masks())
            bobs, bmasks = rows_to_policy_inputs(torch.from_numpy(o_row[None]), torch.from_numpy(m_row[None]))
            _, heads, _, _ = policy.act(bobs, None, None, bmasks)
            row = actions_to_rows(heads)[0].numpy()
            types_seen.add(int(row[L.A_TYPE]))
            obs, _, done, _ = env.step(H.action_to_reference(row))
            if done:
                obs = env.reset()
    assert len(types_seen) >= 6
