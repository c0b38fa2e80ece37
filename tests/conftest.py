import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line(
        'markers', 'allow_overflow: the test provokes a physics capacity overflow on purpose')


@pytest.fixture(scope='session')
def built():
    """Native pieces compiled once per session (nvcc cross-compiles on CPU)."""
    import __graft_entry__ as entry
    entry.build_cuda()
    entry.build_oracle()
    return True


DEMO_KW = dict(res_hw=(384, 384), fps=8, phys_steps=10, phys_iter=10)


def demo_tasks():
    from magical_b200.benchmarks import (cluster, find_dupe, fix_colour,
                                         make_line, match_regions,
                                         move_to_corner, move_to_region)
    return {
        'MoveToCorner': (move_to_corner.MoveToCornerEnv, 80),
        'MoveToRegion': (move_to_region.MoveToRegionEnv, 40),
        'MatchRegions': (match_regions.MatchRegionsEnv, 120),
        'MakeLine': (make_line.MakeLineEnv, 180),
        'FindDupe': (find_dupe.FindDupeEnv, 100),
        'FixColour': (fix_colour.FixColourEnv, 60),
        'ClusterColour': (cluster.ClusterColourEnv, 240),
        'ClusterShape': (cluster.ClusterShapeEnv, 240),
    }


def make_demo_task(name, **extra):
    cls, ep_len = demo_tasks()[name]
    kw = dict(DEMO_KW, max_episode_steps=ep_len)
    kw.update(extra)
    return cls(**kw)


@pytest.fixture(autouse=True)
def _no_physics_overflow(request, monkeypatch):
    """Bit-exactness claims must fail loudly when the physics hits a capacity
    limit: every `MagicalVecEnv` a GPU test closes is
    checked for `overflow_count() == 0` first.  Tests that provoke an
    overflow on purpose carry `@pytest.mark.allow_overflow`."""
    if request.node.get_closest_marker('gpu') is None \
            or request.node.get_closest_marker('allow_overflow') is not None:
        yield
        return
    from magical_b200 import vec_env
    seen = []
    orig_close = vec_env.MagicalVecEnv.close

    def checked_close(self):
        if getattr(self, '_h', None):
            seen.append(self.overflow_count())
        orig_close(self)

    monkeypatch.setattr(vec_env.MagicalVecEnv, 'close', checked_close)
    yield
    assert all(v == 0 for v in seen), \
        f'physics capacity overflow in {sum(1 for v in seen if v)} handle(s): {seen}'
