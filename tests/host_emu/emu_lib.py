"""TEST INFRASTRUCTURE — builds and binds the host emulation of the PRODUCT core
(``settlers_of_catan_rl_b200/csrc/catan_core.cuh`` compiled by g++ with one lane per warp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)
from settlers_of_catan_rl_b200 import layout as L  # noqa: E402
from oracle.oracle_lib import Config, make_config  # noqa: E402  (config struct + reference defaults)

SO = os.path.join(_HERE, "libcatan_emu.so")
_SRCS = [os.path.join(_HERE, "emu.cpp"), os.path.join(_ROOT, "settlers_of_catan_rl_b200", "csrc", "catan_core.cuh"),
         os.path.join(_ROOT, "settlers_of_catan_rl_b200", "csrc", "catan_game.cuh"),
         os.path.join(_ROOT, "include", "catan_layout.h"), os.path.join(_ROOT, "include", "catan_topology.h")]


# "search": the incremental longest-road rule gives up at once (CATAN_LR_FAST_ITERS=1), so every road placement goes through
# the pooled search of the paths through the new road -- the code lr_slow_kernel runs for ~7 % of the updates on the device
SO_SEARCH = os.path.join(_HERE, "libcatan_emu_search.so")


def build(force=False, flavor="default"):
    so, extra = (SO, []) if flavor == "default" else (SO_SEARCH, ["-DCATAN_LR_FAST_ITERS=1"])
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in _SRCS)
    if stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
                               "-Wno-unused-but-set-variable", "-shared"] + extra + ["-o", so, _SRCS[0]])
    return so


_libs = {}


def lib(flavor="default"):
    if flavor not in _libs:
        l = C.CDLL(build(flavor=flavor))
        l.emu_create.restype = C.c_void_p
        l.emu_create.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(Config)]
        l.emu_destroy.argtypes = [C.c_void_p]
        l.emu_set_config.argtypes = [C.c_void_p, C.POINTER(Config)]
        l.emu_reset.argtypes = [C.c_void_p]
        l.emu_step.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
        l.emu_step.restype = C.c_int
        l.emu_sample.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        l.emu_obs.argtypes = [C.c_void_p]
        l.emu_obs.restype = C.POINTER(C.c_uint8)
        l.emu_masks.argtypes = [C.c_void_p]
        l.emu_masks.restype = C.POINTER(C.c_uint8)
        l.emu_export_state.argtypes = [C.c_void_p, C.POINTER(C.c_int16)]
        l.emu_import_state.argtypes = [C.c_void_p, C.POINTER(C.c_int16)]
        l.emu_randomise_uncertainty.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.emu_randomise_uncertainty.restype = C.c_int
        l.emu_longest_path.argtypes = [C.c_void_p, C.c_int]
        l.emu_longest_path.restype = C.c_int
        _libs[flavor] = l
    return _libs[flavor]


class EmuEnv:
    def __init__(self, seed=0, env_id=0, flavor="default", **cfg):
        self.l = lib(flavor)
        self.cfg = make_config(**cfg)
        self.h = self.l.emu_create(seed, env_id, C.byref(self.cfg))

    def __del__(self):
        try:
            self.l.emu_destroy(self.h)
        except Exception:
            pass

    def reset(self):
        self.l.emu_reset(self.h)

    def obs(self):
        return np.ctypeslib.as_array(self.l.emu_obs(self.h), shape=(L.OBS_STRIDE,)).copy()

    def masks(self):
        return np.ctypeslib.as_array(self.l.emu_masks(self.h), shape=(L.MASK_STRIDE,)).copy()

    def state(self):
        s = np.zeros(L.STATE_WORDS, dtype=np.int16)
        self.l.emu_export_state(self.h, s.ctypes.data_as(C.POINTER(C.c_int16)))
        return s

    def import_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.int16)
        self.l.emu_import_state(self.h, s.ctypes.data_as(C.POINTER(C.c_int16)))

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.int32)
        r = np.zeros(4, dtype=np.float32)
        info = np.zeros(L.INFO_STRIDE, dtype=np.uint8)
        err = self.l.emu_step(self.h, a.ctypes.data_as(C.POINTER(C.c_int32)), r.ctypes.data_as(C.POINTER(C.c_float)),
                              info.ctypes.data_as(C.POINTER(C.c_uint8)))
        return err, r, info

    def sample(self):
        a = np.zeros(L.ACTION_WORDS, dtype=np.int32)
        self.l.emu_sample(self.h, a.ctypes.data_as(C.POINTER(C.c_int32)))
        return a

    def randomise_uncertainty(self, controlling_pid, max_attempts=100000):
        return self.l.emu_randomise_uncertainty(self.h, int(controlling_pid), int(max_attempts))

    def longest_path(self, pid):
        return self.l.emu_longest_path(self.h, pid)
