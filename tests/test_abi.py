"""The C-ABI library loads and exports exactly what include/catan_b200.h declares (CPU; no compute calls)."""
import ctypes as C
import os
import re

import pytest

from settlers_of_catan_rl_b200 import _lib, layout as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "catan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(catan_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.library_path()):
        from settlers_of_catan_rl_b200.build import build_extension
        build_extension()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
        assert n in _lib.ABI, "header declares %s but the Python binding does not know it" % n
    assert sorted(_lib.ABI) == names


def test_layout_constants_agree(lib):
    assert lib.catan_obs_stride() == L.OBS_STRIDE
    assert lib.catan_mask_stride() == L.MASK_STRIDE
    assert lib.catan_info_stride() == L.INFO_STRIDE
    assert lib.catan_action_words() == L.ACTION_WORDS
    assert lib.catan_state_words() == L.STATE_WORDS == 721
    assert lib.catan_record_bytes() % 16 == 0
    assert L.OBS_STRIDE % 16 == 0 and L.MASK_STRIDE % 16 == 0
    assert sum(int(__import__("numpy").prod(s)) for _, s in L.MASK_HEADS) == L.MASK_ENTRIES == 325
    assert L.OBS_FEATURES == 12 + 6 + 19 * 60 + 152 + 3 * 159 == 1787


def test_default_config_is_the_reference_default(lib):
    cfg = _lib.make_config()
    assert (cfg.max_actions_per_turn, cfg.max_proposed_trades_per_turn, cfg.validate_actions, cfg.dense_reward) == (-1, 4, 1, 0)
    assert cfg.win_reward == 500.0 and cfg.reward_annealing_factor == 1.0


def test_no_cpu_fallback(lib):
    """Without a GPU the engine must fail loudly, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.catan_create(4, 0, 0, 0, None, C.byref(h)) != 0
    assert b"no CUDA device" in lib.catan_last_error()
    from settlers_of_catan_rl_b200 import VecCatanEnv
    with pytest.raises(_lib.CatanError):
        VecCatanEnv(4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "settlers_of_catan_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\".*oracle", src, flags=re.M), f
