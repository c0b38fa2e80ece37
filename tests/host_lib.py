"""ctypes binding of tests/host/libmg_host.so: the PRODUCT's per-lane device
functions (narrowphase, scene-aux builder) compiled for the host, so CPU tests
can exercise them without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host', 'narrowphase_host.cpp')
LIB = os.path.join(HERE, 'host', 'libmg_host.so')
_lib = None


def lib():
    global _lib
    if _lib is None:
        root = os.path.dirname(HERE)
        deps = [SRC] + [os.path.join(root, 'magical_b200', 'csrc', f)
                        for f in ('mg_narrowphase.h', 'mg_scene_aux.h',
                                  'mg_sincos.h')]
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
                map(os.path.getmtime, deps)):
            subprocess.run(['g++', '-O2', '-fPIC', '-shared', '-std=c++17',
                            '-ffp-contract=off', '-o', LIB, SRC, '-lm'],
                           check=True)
        L = ctypes.CDLL(LIB)
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        L.mgh_collide.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.mgh_gjk.argtypes = [vp, vp, i32, i32, vp]
        _lib = L
    return _lib


def collide(scene, poses, sa, sb):
    scene = np.ascontiguousarray(scene)
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    out = np.zeros(12)
    h = np.zeros(2, dtype=np.uint32)
    ab = np.zeros(2, dtype=np.int32)
    rc = lib().mgh_collide(scene.ctypes.data, poses.ctypes.data, sa, sb,
                           out.ctypes.data, h.ctypes.data, ab.ctypes.data)
    assert rc == 0
    cnt = int(out[0])
    return (int(ab[0]), int(ab[1]), out[1:3].copy(), cnt,
            out[3:11].reshape(2, 4)[:, :2].copy(),
            out[3:11].reshape(2, 4)[:, 2:].copy(), h, bool(out[11]))


def gjk(scene, poses, sa, sb):
    scene = np.ascontiguousarray(scene)
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    out = np.zeros(7)
    rc = lib().mgh_gjk(scene.ctypes.data, poses.ctypes.data, sa, sb,
                       out.ctypes.data)
    assert rc == 0
    return out[0], out[1:3].copy(), out[3:5].copy(), out[5:7].copy()
