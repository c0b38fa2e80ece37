"""ctypes binding of tests/host/libtpe_host.so: the PRODUCT's thread-per-env
physics step (magical_b200/csrc/mg_physics_tpe.h) compiled for the host, so
CPU tests can check the kernel's own source against the oracle."""
import ctypes
import os
import subprocess

import numpy as np

from magical_b200 import scene as sc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host', 'tpe_host.cpp')
LIB = os.path.join(HERE, 'host', 'libtpe_host.so')
CUDA_INC = os.environ.get('CUDA_INC', '/usr/local/cuda/include')
_libs = {}


def lib(max_surv=None):
    """max_surv: build a variant with a smaller cooperative-survivor list (to
    exercise the owner-serial continuation)."""
    global _libs
    LIB = os.path.join(HERE, 'host', 'libtpe_host%s.so' % (
        '' if max_surv is None else '_surv%d' % max_surv))
    _lib = _libs.get(max_surv)
    if _lib is None:
        root = os.path.dirname(HERE)
        csrc = os.path.join(root, 'magical_b200', 'csrc')
        deps = [SRC, os.path.join(root, 'include', 'magical_b200.h')] + [
            os.path.join(csrc, f) for f in os.listdir(csrc)
            if f.endswith(('.h', '.cuh'))]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
                map(os.path.getmtime, deps)):
            subprocess.run(['g++', '-O2', '-fPIC', '-shared', '-std=c++17',
                            '-ffp-contract=off', '-I', CUDA_INC]
                           + ([] if max_surv is None
                              else ['-DTPE_MAX_SURV=%d' % max_surv])
                           + ['-o', LIB, SRC, '-lm'], check=True)
        L = ctypes.CDLL(LIB)
        vp, i32, f64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.tpeh_create.restype = vp
        L.tpeh_create.argtypes = [vp, i32, i32, i32, i32]
        L.tpeh_destroy.argtypes = [vp]
        L.tpeh_reset.argtypes = [vp]
        L.tpeh_step.argtypes = [vp, i32]
        L.tpeh_get_state.argtypes = [vp, vp]
        L.tpeh_set_state.argtypes = [vp, vp]
        L.tpeh_set_pose.argtypes = [vp, i32, f64, f64, f64]
        L.tpeh_kcon.argtypes = [vp]
        L.tpeh_words.argtypes = [vp]
        _libs[max_surv] = _lib = L
    return _lib


class TpeHostEnv:
    def __init__(self, scene_record, kcon=8, spill=True, nitems=32,
                 max_surv=None, scratch_global=True):
        self._lib = lib(max_surv)
        self._scene = np.ascontiguousarray(scene_record).copy()
        self._h = self._lib.tpeh_create(self._scene.ctypes.data, kcon,
                                        int(spill), nitems,
                                        int(scratch_global))
        assert self._h, 'scene rejected by the thread-per-env path'

    @property
    def kcon(self):
        return self._lib.tpeh_kcon(self._h)

    @property
    def words(self):
        return self._lib.tpeh_words(self._h)

    def step(self, action):
        self._lib.tpeh_step(self._h, int(action))

    def reset(self):
        self._lib.tpeh_reset(self._h)

    def set_pose(self, body, x, y, angle):
        self._lib.tpeh_set_pose(self._h, body, x, y, angle)

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=sc.state_dt)
        assert self._lib.tpeh_set_state(self._h, st.ctypes.data) == 0

    def state(self):
        st = np.zeros((), dtype=sc.state_dt)
        self._lib.tpeh_get_state(self._h, st.ctypes.data)
        return st

    def close(self):
        if self._h:
            self._lib.tpeh_destroy(self._h)
            self._h = None

    __del__ = close


def rollout_mismatches(job):
    """Worker for the host sweep: job = (scene_records, actions[T, n]); steps the
    host-compiled kernel source and the oracle side by side and returns, per
    column, None or (step, max |pos difference|, overflow flags) of the first
    step at which poses differ or the kernel flags an overflow."""
    from oracle_lib import OracleEnv
    scenes, actions = job
    out = []
    for e in range(actions.shape[1]):
        orc = OracleEnv(scenes[e], det_sincos=True)
        env = TpeHostEnv(scenes[e], kcon=4, nitems=48)
        bad = None
        for t in range(actions.shape[0]):
            orc.step(int(actions[t, e]))
            env.step(int(actions[t, e]))
            a, b = orc.state(), env.state()
            nb = int(a['n_bodies'])
            if int(b['overflow']) or not (
                    np.array_equal(a['pos'][:nb], b['pos'][:nb])
                    and np.array_equal(a['angle'][:nb], b['angle'][:nb])):
                bad = (t, float(np.abs(a['pos'][:nb] - b['pos'][:nb]).max()),
                       int(b['overflow']))
                break
        out.append(bad)
        orc.close()
        env.close()
    return out
