"""The shared deterministic sin/cos (magical_b200/csrc/mg_sincos.h) against libm."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_det_sincos_accuracy():
    src = '#include "%s/magical_b200/csrc/mg_sincos.h"\nvoid sc(double a, double* s, double* c) { mg_det_sincos(a, s, c); }\n' % ROOT
    with tempfile.TemporaryDirectory() as d:
        cfile, so = os.path.join(d, 's.c'), os.path.join(d, 's.so')
        open(cfile, 'w').write(src)
        subprocess.run(['gcc', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-o', so, cfile, '-lm'], check=True)
        lib = ctypes.CDLL(so)
        lib.sc.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        rng = np.random.RandomState(0)
        xs = np.concatenate([rng.uniform(-10, 10, 4000), rng.uniform(-400, 400, 4000),
                             np.arange(-8, 9) * np.pi / 4, [0.0, -0.0, 1e-300, 0.55 * np.pi, -2.13]])
        s, c = ctypes.c_double(), ctypes.c_double()
        worst = 0.0
        for x in xs:
            lib.sc(float(x), ctypes.byref(s), ctypes.byref(c))
            worst = max(worst, abs(s.value - np.sin(x)), abs(c.value - np.cos(x)))
            assert abs(s.value ** 2 + c.value ** 2 - 1) < 4e-16
        assert worst < 2.3e-16, worst
