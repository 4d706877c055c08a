"""League sampling weights (RL/ppo/update_opponent_policies.py:29-43) against the reference's own function where the
reference tree exists, and against their closed form everywhere."""
import sys
import types

import numpy as np
import pytest

from oracle import ref_harness as H
from settlers_of_catan_rl_b200.self_play import league_probabilities, sample_opponents


def test_weights_closed_form():
    for n in (1, 2, 7, 800, 801, 2500):
        p = league_probabilities(n)
        assert p.shape == (n,) and abs(p.sum() - 1.0) < 1e-12 and bool((np.diff(p) >= -1e-18).all())
        k = min(800, n)
        assert np.allclose(p[:n - k], p[0]) and (n == k or p[n - k] == p[0])      # older than the window: the uniform share only
    assert np.allclose(league_probabilities(1), [1.0])
    rng = np.random.default_rng(0)
    picks = sample_opponents(list(range(50)), rng)
    assert len(picks) == 3 and all(0 <= q < 50 for q in picks)
    counts = np.bincount([sample_opponents(list(range(4)), rng, 1)[0] for _ in range(20000)], minlength=4) / 20000
    assert np.allclose(counts, league_probabilities(4), atol=0.015)


@pytest.mark.skipif(not H.reference_available(), reason="reference tree not present")
def test_weights_equal_the_reference_function():
    H.import_reference()
    for name in ("matplotlib", "matplotlib.pyplot"):                       # update_opponent_policies.py:2 imports it, unused
        sys.modules.setdefault(name, types.ModuleType(name))
    from RL.ppo.update_opponent_policies import get_prob_dist  # type: ignore
    for n in (1, 2, 3, 10, 799, 800, 801, 1600):
        assert np.allclose(league_probabilities(n), get_prob_dist(n), rtol=1e-12, atol=0), n
    assert np.allclose(league_probabilities(30, linear_num=10, linear_prob=0.3), get_prob_dist(30, linear_num=10, linear_prob=0.3), rtol=1e-12)
