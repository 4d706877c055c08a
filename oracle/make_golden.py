"""TEST INFRASTRUCTURE — generate the golden fixtures under tests/golden/ from the REAL reference.

Needs /root/reference (this container).  The reference is run under the shared-Philox RNG
(oracle/ref_harness.py) with the pinned random-legal sampler; every step's canonical state, packed
observation, packed masks, rewards and done flag are stored.  The GPU box has no reference: there the
fixtures are the reference.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), "tests", "golden")

#: name -> (seed, env_id, n_games, EnvWrapper kwargs)
CASES = {
    "default_s0": (0, 1, 1, {}),
    "default_s1": (1, 8, 1, {}),
    "default_s5_two_games": (5, 36, 2, {}),
    "default_s11_two_games": (11, 78, 2, {}),
    "dense_reward_s3": (3, 22, 1, dict(dense_reward=True)),
    "no_trade_cap_s7": (7, 50, 1, dict(max_proposed_trades_per_turn=None)),
    "no_trades_max_actions_s9": (9, 64, 1, dict(max_proposed_trades_per_turn=0, max_actions_per_turn=6)),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (seed, env_id, n_games, kw) in CASES.items():
        g = H.record_game(seed, env_id, max_steps=12000, env_kwargs=kw, n_games=n_games)
        cfg = dict(
            max_actions_per_turn=-1 if kw.get("max_actions_per_turn") is None else kw["max_actions_per_turn"],
            max_proposed_trades_per_turn=(-1 if ("max_proposed_trades_per_turn" in kw and kw["max_proposed_trades_per_turn"] is None)
                                          else kw.get("max_proposed_trades_per_turn", 4)),
            dense_reward=int(bool(kw.get("dense_reward", False))),
        )
        g["cfg_keys"] = np.array(sorted(cfg.keys()))
        g["cfg_vals"] = np.array([cfg[k] for k in sorted(cfg.keys())], dtype=np.int64)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **g)
        print(name, "steps", len(g["actions"]), "games", int(g["done"].sum()), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
