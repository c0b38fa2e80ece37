"""The oracle is frozen: seeded full-episode rollouts of all 8 Demo tasks must
reproduce the committed poses / scores / frame checksums (tests/golden/
oracle_pins.json, written by make_oracle_pins.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_its_pins():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'make_oracle_pins', os.path.join(HERE, 'golden', 'make_oracle_pins.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(HERE, 'golden', 'oracle_pins.json')) as fh:
        want = json.load(fh)
    got = mod.compute()
    assert set(got) == set(want)
    for name in want:
        w, g = want[name], got[name]
        assert g['done'] == w['done'] and g['steps'] == w['steps']
        assert g['score'] == w['score'], name
        # identical arithmetic -> identical bits; 1e-12 leaves room for a libm
        # whose pow/exp (scene constants only) differs in the last place
        assert np.allclose(g['pos'], w['pos'], rtol=0, atol=1e-12), name
        assert np.allclose(g['angle'], w['angle'], rtol=0, atol=1e-12), name
        assert g['ego_crc32'] == w['ego_crc32'], name
        assert g['allo_crc32'] == w['allo_crc32'], name
