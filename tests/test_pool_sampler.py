"""Background scene sampler (magical_b200/pool_sampler.py): the scene stream
depends on (task, seed) only, not on the number of workers, and every scene
is a valid collision-free layout."""
import numpy as np

from magical_b200.env import make_task
from magical_b200.pool_sampler import ScenePoolSampler, sample_chunk

ENV_ID = 'MatchRegions-TestAll-LoResStack-v0'


def _bytes(scenes):
    return np.ascontiguousarray(scenes).view(np.uint8)


def test_stream_is_independent_of_worker_count():
    task, _ = make_task(ENV_ID)
    with ScenePoolSampler(task, workers=1, seed=11, chunk=4) as one, \
            ScenePoolSampler(task, workers=3, seed=11, chunk=4) as three:
        a = np.concatenate([one.take(6), one.take(5)])
        b = three.take(11)
    assert a.shape == b.shape == (11,)
    assert np.array_equal(_bytes(a), _bytes(b))
    # and equal to sampling the chunks in this process
    ref = np.concatenate([sample_chunk(task, 11, k, 4) for k in range(3)])[:11]
    assert np.array_equal(_bytes(a), _bytes(ref))
    # another seed gives another stream; scenes within a stream differ
    other = sample_chunk(task, 12, 0, 4)
    assert not np.array_equal(_bytes(other), _bytes(ref[:4]))
    assert len({_bytes(s).tobytes() for s in a}) == len(a)


def test_non_blocking_take_and_ready():
    task, _ = make_task(ENV_ID)
    with ScenePoolSampler(task, workers=2, seed=1, chunk=2, prefetch=1) as s:
        assert s.take(10 ** 6, block=False) is None
        got = s.take(3)
        assert len(got) == 3
        assert s.ready() >= 1      # the tail of the second chunk
        assert len(s.take(1, block=False)) == 1
