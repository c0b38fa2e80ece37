"""ctypes binding of the CPU oracle (oracle/libmgo_oracle.so) for tests.

Test infrastructure only: nothing under magical_b200/ imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

from magical_b200 import scene as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, 'oracle')
LIB_PATH = os.path.join(ORACLE_DIR, 'libmgo_oracle.so')

_lib = None


def build():
    subprocess.run(['make', '-C', ORACLE_DIR, '--quiet'], check=True)


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR)
                if f.endswith(('.c', '.h'))]
        if (not os.path.exists(LIB_PATH)
                or os.path.getmtime(LIB_PATH) < max(map(os.path.getmtime,
                                                        srcs))):
            build()
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_double
        L.mgo_create.restype = vp
        L.mgo_create.argtypes = [vp]
        L.mgo_destroy.argtypes = [vp]
        L.mgo_reset.argtypes = [vp]
        L.mgo_set_det_sincos.argtypes = [vp, i32]
        L.mgo_set_pair_permutation.argtypes = [vp, vp]
        L.mgo_set_action.argtypes = [vp, i32]
        L.mgo_robot_update.argtypes = [vp]
        L.mgo_space_step.argtypes = [vp, f64]
        L.mgo_phys_steps_on_frame.argtypes = [vp]
        L.mgo_step.argtypes = [vp, i32, vp, vp, vp]
        L.mgo_get_state.argtypes = [vp, vp]
        L.mgo_set_state.argtypes = [vp, vp]
        L.mgo_set_pose.argtypes = [vp, i32, f64, f64, f64]
        L.mgo_score.restype = f64
        L.mgo_score.argtypes = [vp]
        L.mgo_debug_reward.restype = f64
        L.mgo_debug_reward.argtypes = [vp]
        L.mgo_block_in_goal.restype = i32
        L.mgo_block_in_goal.argtypes = [vp, i32, i32]
        L.mgo_render_view.argtypes = [vp, i32, i32, vp]
        L.mgo_collide.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.mgo_downsample4.argtypes = [vp, i32, vp]
        L.mgo_sizeof_scene.restype = ctypes.c_int64
        L.mgo_sizeof_state.restype = ctypes.c_int64
        assert L.mgo_sizeof_scene() == sc.scene_dt.itemsize, \
            (L.mgo_sizeof_scene(), sc.scene_dt.itemsize)
        assert L.mgo_sizeof_state() == sc.state_dt.itemsize, \
            (L.mgo_sizeof_state(), sc.state_dt.itemsize)
        _lib = L
    return _lib


class OracleEnv:
    """One environment stepped by the CPU oracle."""

    def __init__(self, scene_record, det_sincos=False):
        self._lib = lib()
        self._scene = np.ascontiguousarray(scene_record).copy()
        self._h = self._lib.mgo_create(self._scene.ctypes.data)
        assert self._h
        if det_sincos:
            self._lib.mgo_set_det_sincos(self._h, 1)

    def close(self):
        if self._h:
            self._lib.mgo_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        self._lib.mgo_reset(self._h)

    def set_pair_permutation(self, perm):
        if perm is None:
            self._lib.mgo_set_pair_permutation(self._h, None)
        else:
            perm = np.ascontiguousarray(perm, dtype=np.int32)
            self._lib.mgo_set_pair_permutation(self._h, perm.ctypes.data)

    def step(self, action):
        rew = ctypes.c_float()
        done = ctypes.c_uint8()
        score = ctypes.c_float()
        self._lib.mgo_step(self._h, int(action), ctypes.byref(rew),
                           ctypes.byref(done), ctypes.byref(score))
        return rew.value, bool(done.value), score.value

    def state(self):
        st = np.zeros((), dtype=sc.state_dt)
        self._lib.mgo_get_state(self._h, st.ctypes.data)
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=sc.state_dt)
        self._lib.mgo_set_state(self._h, st.ctypes.data)

    def set_pose(self, body, x, y, angle):
        self._lib.mgo_set_pose(self._h, body, x, y, angle)

    def score(self):
        return self._lib.mgo_score(self._h)

    def block_in_goal(self, block, goal):
        return bool(self._lib.mgo_block_in_goal(self._h, block, goal))

    def collide(self, sa, sb):
        """Oracle narrowphase for shapes (sa, sb): returns (a, b, n, count,
        p1[2,2], p2[2,2], hash[2]) with (a, b) in Chipmunk's type order."""
        ia, ib, cnt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        n = np.zeros(2)
        p1 = np.zeros((2, 2))
        p2 = np.zeros((2, 2))
        h = np.zeros(2, dtype=np.uint32)
        self._lib.mgo_collide(self._h, sa, sb, ctypes.byref(ia),
                              ctypes.byref(ib), n.ctypes.data,
                              ctypes.byref(cnt), p1.ctypes.data,
                              p2.ctypes.data, h.ctypes.data)
        return ia.value, ib.value, n, cnt.value, p1, p2, h

    def render_view(self, view, res=384):
        out = np.zeros((res, res, 3), dtype=np.uint8)
        self._lib.mgo_render_view(self._h, view, res, out.ctypes.data)
        return out

    def render_lores(self, view):
        """384 render + 4x4 INTER_AREA mean -> (96, 96, 3)."""
        full = self.render_view(view, 384)
        out = np.zeros((96, 96, 3), dtype=np.uint8)
        self._lib.mgo_downsample4(full.ctypes.data, 96, out.ctypes.data)
        return out


def downsample4(img):
    n = img.shape[0] // 4
    img = np.ascontiguousarray(img)
    out = np.zeros((n, n, 3), dtype=np.uint8)
    lib().mgo_downsample4(img.ctypes.data, n, out.ctypes.data)
    return out


def rollout_scores(job):
    """Worker for the score sweeps: job = (scene_records, actions[T, n]); steps
    one oracle per column through its whole action stream (physics + score,
    no render) and returns (score at the done step [n] f32, step of done [n],
    final positions [n, MAX_BODIES, 2])."""
    scenes, actions = job
    n = actions.shape[1]
    score = np.full(n, np.nan, dtype=np.float32)
    done_at = np.full(n, -1, dtype=np.int32)
    pos = None
    for e in range(n):
        orc = OracleEnv(scenes[e], det_sincos=True)
        for t in range(actions.shape[0]):
            _, d, s = orc.step(int(actions[t, e]))
            if d and done_at[e] < 0:
                done_at[e], score[e] = t, np.float32(s)
        st = orc.state()
        if pos is None:
            pos = np.zeros((n,) + st['pos'].shape)
        pos[e] = st['pos']
        orc.close()
    return score, done_at, pos


def rollout_frames(job):
    """Worker for the render sweep: job = (scene_records, actions[T, n]); steps
    one oracle per column through its action stream and returns the 96x96
    allo and ego frames of the final state, [n, 2, 96, 96, 3] u8."""
    scenes, actions = job
    n = actions.shape[1]
    out = np.zeros((n, 2, 96, 96, 3), dtype=np.uint8)
    for e in range(n):
        orc = OracleEnv(scenes[e], det_sincos=True)
        for t in range(actions.shape[0]):
            orc.step(int(actions[t, e]))
        out[e, 0] = orc.render_lores(0)
        out[e, 1] = orc.render_lores(1)
        orc.close()
    return out
