"""Pin the oracle against the REAL reference, live (only where /root/reference exists — this container)."""
import numpy as np
import pytest

from oracle import ref_harness as H
from oracle import oracle_lib as O
from tests.common import replay_golden
from tests.test_oracle_golden import _OracleAdapter

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("seed,env_id,kw", [(101, 3, {}), (202, 4, dict(dense_reward=True))])
def test_fresh_reference_game_matches_oracle(seed, env_id, kw):
    g = H.record_game(seed, env_id, max_steps=700, env_kwargs=kw)
    g["cfg"] = dict(dense_reward=int(bool(kw.get("dense_reward", False))))
    assert replay_golden(_OracleAdapter(g), g) == len(g["actions"])


def test_topology_header_is_current():
    from oracle import gen_topology
    assert open(gen_topology.OUT).read() == gen_topology.render()


def test_python_philox_matches_c():
    s = H.PhiloxStream(0xDEADBEEF12345678, 77, 1)
    for blk in (0, 1, 12345):
        assert s.block(blk) == O.philox(blk, 1, 77, 0, 0x12345678, 0xDEADBEEF)
