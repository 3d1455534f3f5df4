"""ctypes binding of oracle/libovc_oracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by the product package.  All arrays are host numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libovc_oracle.so")
    src = os.path.join(_HERE, "ovc_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ovc_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-pthread", "-std=gnu11", "-shared", "-o", so, src], cwd=_HERE
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libovc_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.ovo_layout_table_size.restype = ctypes.c_size_t
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


class RandomStart(ctypes.Structure):
    """ovc_random_start_t"""
    _fields_ = [("seed", ctypes.c_uint64), ("obj_threshold", ctypes.c_uint32), ("random_start_pos", ctypes.c_int32),
                ("random_layout", ctypes.c_int32), ("reserved", ctypes.c_int32)]


def random_start(seed, rnd_obj_prob_thresh=0.0, random_start_pos=False, random_layout=False):
    thr = min(int(rnd_obj_prob_thresh * 4294967296.0), 0xFFFFFFFF)
    return RandomStart(int(seed) & 0xFFFFFFFFFFFFFFFF, thr, int(bool(random_start_pos)), int(bool(random_layout)), 0)


def _rs(rs):
    return ctypes.byref(rs) if rs is not None else None


def reset_random(tables, starts, state, rs, env_layout=None, mask=None):
    assert state.dtype == np.int32 and state.flags.c_contiguous
    n, S = state.shape
    tables, starts = np.ascontiguousarray(tables), _i32(starts)
    el = None if env_layout is None else _i32(env_layout)
    mk = None if mask is None else _i32(mask)
    rc = lib().ovo_reset_random(_p(tables), ctypes.c_int(len(tables)), _p(starts), _p(state), None if el is None else _p(el),
                                None if mk is None else _p(mk), ctypes.c_int64(n), ctypes.c_int(S), _rs(rs))
    assert rc == 0


def step(tables, starts, state, actions, horizon=400, flags=0, n_threads=1, rs=None):
    """One transition in place on ``state`` [N,S] int32.  Returns sparse[N], shaped[N,2], done[N], events[N,2]."""
    assert state.dtype == np.int32 and state.flags.c_contiguous
    n, S = state.shape
    actions = _i32(actions)
    sparse = np.zeros(n, np.int32)
    shaped = np.zeros((n, 2), np.int32)
    done = np.zeros(n, np.int32)
    events = np.zeros((n, 2), np.int32)
    tables = np.ascontiguousarray(tables)
    starts = _i32(starts)
    rc = lib().ovo_step(_p(tables), ctypes.c_int(len(tables)), _p(starts), _p(state), _p(actions), _p(sparse),
                        _p(shaped), _p(done), _p(events), ctypes.c_int64(n), ctypes.c_int(S), ctypes.c_int(horizon),
                        ctypes.c_int(flags), ctypes.c_int(n_threads), _rs(rs))
    assert rc == 0
    return sparse, shaped, done, events


def alloc_rollout_out(T, n):
    """Output arrays for rollout(), pages already touched (so a timed call does not pay first-touch faults)."""
    out = (np.empty((T, n), np.int32), np.empty((T, n, 2), np.int32), np.empty((T, n), np.int32), np.empty((T, n, 2), np.int32))
    for o in out:
        o.fill(0)
    return out


def rollout(tables, starts, state, actions, horizon=400, flags=0, n_threads=0, rs=None, out=None):
    """T transitions in place; actions [T,N,2].  Returns sparse[T,N], shaped[T,N,2], done[T,N], events[T,N,2]."""
    assert state.dtype == np.int32 and state.flags.c_contiguous
    n, S = state.shape
    actions = _i32(actions)
    T = actions.shape[0]
    assert actions.shape == (T, n, 2)
    sparse, shaped, done, events = out if out is not None else alloc_rollout_out(T, n)
    tables = np.ascontiguousarray(tables)
    starts = _i32(starts)
    rc = lib().ovo_rollout(_p(tables), ctypes.c_int(len(tables)), _p(starts), _p(state), _p(actions), _p(sparse),
                           _p(shaped), _p(done), _p(events), ctypes.c_int64(n), ctypes.c_int(T), ctypes.c_int(S),
                           ctypes.c_int(horizon), ctypes.c_int(flags), ctypes.c_int(n_threads), _rs(rs))
    assert rc == 0
    return sparse, shaped, done, events


def max_threads():
    return int(lib().ovo_max_threads())


def encode_lossless(tables, state, width, height, horizon=400):
    """int32 [N,2,W,H,26]"""
    state = _i32(state)
    n, S = state.shape
    out = np.zeros((n, 2, width, height, 26), np.int32)
    tables = np.ascontiguousarray(tables)
    rc = lib().ovo_encode_lossless(_p(tables), ctypes.c_int(len(tables)), _p(state), _p(out), ctypes.c_int64(n),
                                   ctypes.c_int(S), ctypes.c_int(width), ctypes.c_int(height), ctypes.c_int(horizon))
    assert rc == 0, "layout shape mismatch"
    return out


def featurize(tables, lut, state, num_pots=2):
    """float64 [N,2,F], F = 2*(10*num_pots+28)"""
    state = _i32(state)
    n, S = state.shape
    F = 2 * (10 * num_pots + 28)
    out = np.zeros((n, 2, F), np.float64)
    tables = np.ascontiguousarray(tables)
    lut = np.ascontiguousarray(lut)
    rc = lib().ovo_featurize(_p(tables), ctypes.c_int(len(tables)), _p(lut), _p(state), _p(out), ctypes.c_int64(n),
                             ctypes.c_int(S), ctypes.c_int(num_pots))
    assert rc == 0
    return out


def potential(tables, pot_tables, cost_lut, gpow, state):
    """float64 [N]: potential_function of every record."""
    state = _i32(state)
    n, S = state.shape
    out = np.zeros(n, np.float64)
    tables, pot_tables, cost_lut = np.ascontiguousarray(tables), np.ascontiguousarray(pot_tables), np.ascontiguousarray(cost_lut)
    gpow = np.ascontiguousarray(gpow, dtype=np.float64)
    rc = lib().ovo_potential(_p(tables), ctypes.c_int(len(tables)), _p(pot_tables), _p(cost_lut), _p(gpow),
                             ctypes.c_int(len(gpow)), _p(state), _p(out), ctypes.c_int64(n), ctypes.c_int(S))
    assert rc == 0
    return out
