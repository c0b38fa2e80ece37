#!/usr/bin/env python3
"""Regression pins of the CPU oracle: final poses, scores and a render checksum of every Demo task after a
seeded random rollout, as computed by the oracle at the commit that was validated against the reference's own
renders and closed forms (tests/test_oracle_golden.py, tests/test_oracle_physics.py).  They do not add
evidence about the reference -- they freeze the oracle so later edits cannot silently change what "parity"
means.  Regenerate deliberately:  python tests/golden/make_oracle_pins.py
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def compute():
    from conftest import demo_tasks, make_demo_task
    from oracle_lib import OracleEnv
    pins = {}
    for name in demo_tasks():
        rec = make_demo_task(name).build_scene()
        orc = OracleEnv(rec, det_sincos=True)
        rng = np.random.RandomState(1234)
        n = int(rec['max_steps'])
        score = 0.0
        for _ in range(n):
            a = int(rng.randint(18)) if rng.rand() < 0.5 else int(rng.choice([1, 4, 7, 10, 13, 16]))
            _, done, score = orc.step(a)
        st = orc.state()
        nb = int(st['n_bodies'])
        pins[name] = {
            'steps': n, 'done': bool(done), 'score': float(score),
            'pos': st['pos'][:nb].tolist(), 'angle': st['angle'][:nb].tolist(),
            'ego_crc32': int(zlib.crc32(orc.render_lores(1).tobytes())),
            'allo_crc32': int(zlib.crc32(orc.render_lores(0).tobytes())),
        }
        orc.close()
    return pins


if __name__ == '__main__':
    with open(os.path.join(HERE, 'oracle_pins.json'), 'w') as fh:
        json.dump(compute(), fh, indent=1)
    print('wrote oracle_pins.json')
