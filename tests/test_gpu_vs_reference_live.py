"""The CUDA path against the LIVE reference on the GPU box: the unmodified reference (the copy ``oracle/build_ref.py`` makes
under the git-ignored ``oracle/_ref/``, which travels with the working tree) plays fresh games under the shared Philox
stream, and the kernels, called through the C ABI, must reproduce every state / obs / mask / reward / done of them.
Unlike the committed fixtures these games are generated at test time, with seeds no fixture holds."""
import numpy as np
import pytest

from oracle import ref_harness as H
from tests.common import replay_golden

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not H.reference_available(), reason="no reference tree (oracle/_ref not built)")]


@pytest.mark.parametrize("seed,env_id,kw,cfg", [
    (9001, 17, {}, {}),
    (9002, 123456, dict(dense_reward=True), dict(dense_reward=1)),
    (9003, 5, dict(max_proposed_trades_per_turn=None), dict(max_proposed_trades_per_turn=-1)),
])
def test_cuda_replays_a_fresh_reference_game(seed, env_id, kw, cfg):
    from tests.test_gpu_parity import _GpuAdapter
    g = H.record_game(seed, env_id, max_steps=900, env_kwargs=kw)
    g["cfg"] = cfg
    assert replay_golden(_GpuAdapter(g), g) == len(g["actions"])
