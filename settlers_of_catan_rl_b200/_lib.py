"""ctypes binding of ``libcatan_b200.so`` (C ABI declared in ``include/catan_b200.h``).

There is deliberately no fallback: if the CUDA library is missing or no GPU is visible, importing
the package still works (so CPU-only tooling can inspect layouts) but every engine call raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import layout as L

# CATAN_B200_LIB lets profiles/phase_profile.py load the instrumented build of the SAME sources; never a fallback
_SO = os.environ.get("CATAN_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcatan_b200.so")


class CatanConfig(C.Structure):
    """``catan_config_t`` == EnvWrapper.__init__ kwargs (env/wrapper.py:12-28)."""

    _fields_ = [
        ("max_actions_per_turn", C.c_int32),
        ("max_proposed_trades_per_turn", C.c_int32),
        ("validate_actions", C.c_int32),
        ("dense_reward", C.c_int32),
        ("auto_reset", C.c_int32),
        ("win_reward", C.c_float),
        ("reward_annealing_factor", C.c_float),
    ]


class CatanRollout(C.Structure):
    """``catan_rollout_t``: device pointers of the time-major rollout buffers."""

    _fields_ = [("obs", C.c_void_p), ("masks", C.c_void_p), ("actions", C.c_void_p), ("logp", C.c_void_p), ("rewards", C.c_void_p),
                ("tmasks", C.c_void_p), ("cursors", C.c_void_p), ("acc", C.c_void_p), ("flags", C.c_void_p),
                ("active_pid", C.c_void_p), ("collecting", C.c_void_p), ("T", C.c_int32), ("N", C.c_int32)]


class CatanError(RuntimeError):
    pass


class CatanMinibatch(C.Structure):
    """catan_minibatch_t"""
    _fields_ = [(n, C.c_void_p) for n in ("obs", "masks", "actions", "logp", "values", "returns", "tmasks", "advantages")]


#: every symbol include/catan_b200.h declares: name -> (restype, argtypes)
_u8p, _i32p, _f32p, _i16p, _u32p, _f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_float, C.c_int16, C.c_uint32, C.c_double))
_vp = C.c_void_p
ABI = {
    "catan_abi_version": (C.c_int, []),
    "catan_obs_stride": (C.c_int, []),
    "catan_mask_stride": (C.c_int, []),
    "catan_info_stride": (C.c_int, []),
    "catan_action_words": (C.c_int, []),
    "catan_state_words": (C.c_int, []),
    "catan_record_bytes": (C.c_int, []),
    "catan_last_error": (C.c_char_p, []),
    "catan_default_config": (None, [C.POINTER(CatanConfig)]),
    "catan_create": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(CatanConfig), C.POINTER(_vp)]),
    "catan_destroy": (C.c_int, [_vp]),
    "catan_num_envs": (C.c_int, [_vp]),
    "catan_set_config": (C.c_int, [_vp, C.POINTER(CatanConfig)]),
    "catan_bind": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "catan_reset": (C.c_int, [_vp, _vp, _vp]),
    "catan_step": (C.c_int, [_vp, _vp, _vp]),
    "catan_step_masked": (C.c_int, [_vp, _vp, _vp, _vp]),
    "catan_rollout_store": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "catan_sample_random": (C.c_int, [_vp, _vp, _vp]),
    "catan_step_sample": (C.c_int, [_vp, _vp, _vp]),
    "catan_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "catan_step_host_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "catan_step_sample_host_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "catan_reset_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "catan_step_sample_host_async_u8": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "catan_step_sample_host_groups": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_longlong)]),
    "catan_export_state": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "catan_import_state": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "catan_read_err_flags": (C.c_int, [_vp, _vp, C.c_int]),
    "catan_randomise_uncertainty": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "catan_read_lr_stats": (C.c_int, [_vp, _vp]),
    "catan_read_lr_histograms": (C.c_int, [_vp, _vp]),
    "catan_set_graphs": (C.c_int, [_vp, C.c_int]),
    "catan_set_timing": (C.c_int, [_vp, C.c_int]),
    "catan_read_timing": (C.c_int, [_vp, _vp]),
    "catan_gae": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp]),
    "catan_adv_stats": (C.c_int, [_vp, C.c_longlong, _vp, _vp]),
    "catan_adv_apply": (C.c_int, [_vp, C.c_longlong, _vp, C.c_double, _vp]),
    "catan_route_by_policy": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "catan_policy_inputs": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "catan_masked_categorical": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "catan_minibatch_gather": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "catan_tile_attention_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "catan_tile_attention_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "catan_ln_small_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_longlong, C.c_int, C.c_float, _vp]),
    "catan_ln_small_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_longlong, C.c_int, _vp]),
}

_lib = None


def library_path() -> str:
    return _SO


def load():
    """Load the shared library and check it agrees with the Python layout constants."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise CatanError(
            "CUDA extension %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc). There is no CPU fallback." % _SO)
    lib = C.CDLL(_SO)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)          # AttributeError here == header and binary disagree
        fn.restype = res
        fn.argtypes = args
    got = (lib.catan_obs_stride(), lib.catan_mask_stride(), lib.catan_info_stride(), lib.catan_action_words(),
           lib.catan_state_words())
    want = (L.OBS_STRIDE, L.MASK_STRIDE, L.INFO_STRIDE, L.ACTION_WORDS, L.STATE_WORDS)
    if got != want:
        raise CatanError("layout mismatch between libcatan_b200.so %r and layout.py %r" % (got, want))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise CatanError(load().catan_last_error().decode("utf-8", "replace"))


def make_config(**kw) -> CatanConfig:
    cfg = CatanConfig()
    load().catan_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError("unknown config field %r" % k)
        setattr(cfg, k, v)
    return cfg
