"""Background scene sampling for the randomised Test* variants.

The reference samples a fresh layout inside every `reset()` on the host
(`base_env.py:177-234` -> task `on_reset` -> `geom.pm_randomise_all_poses`,
`geom.py:294-384`).  Here resets happen on the device, thousands per second,
and draw from a pool of pre-sampled scenes (`MagicalVecEnv`); this module keeps
that pool fresh: worker processes run the rejection sampler
(`magical_b200/placement.py`) ahead of the GPU, and
`MagicalVecEnv.refresh_pool(sampler)` streams the finished scenes into the
half of the pool that is not being drawn from.

The stream of scenes is a function of (task, seed) only: chunk k is always
sampled with `RandomState(SeedSequence([seed, k]))`, and chunks are handed out
in order, so the number of workers changes the rate, not the result.
"""
import collections
import multiprocessing as mp
import pickle

import numpy as np

_task = None


def _init_worker(task_bytes):
    global _task
    _task = pickle.loads(task_bytes)


def _chunk_seed(seed, k):
    return int(np.random.SeedSequence([int(seed), int(k)]).generate_state(1)[0])


def sample_chunk(task, seed, k, n):
    """Chunk k of the scene stream of (task, seed): n scene records."""
    task.seed(_chunk_seed(seed, k))
    return np.stack([task.build_scene() for _ in range(n)])


def _work(seed, k, n):
    return sample_chunk(_task, seed, k, n)


class ScenePoolSampler:
    """Samples scenes of `task` in `workers` processes, `chunk` scenes per job,
    keeping `prefetch` jobs per worker in flight."""

    def __init__(self, task, workers=4, seed=0, chunk=16, prefetch=2,
                 start_method='spawn'):
        self.seed, self.chunk = int(seed), int(chunk)
        self._next = 0
        self._pending = collections.deque()
        self._left = None     # unread tail of the oldest finished chunk
        # spawn: the parent usually holds a CUDA context, which must not be forked
        self._pool = mp.get_context(start_method).Pool(
            workers, initializer=_init_worker,
            initargs=(pickle.dumps(task),))
        self._depth = max(1, workers * prefetch)
        self._fill()

    def _fill(self):
        while len(self._pending) < self._depth:
            self._pending.append(self._pool.apply_async(
                _work, (self.seed, self._next, self.chunk)))
            self._next += 1

    def ready(self):
        """Scenes that `take(block=False)` could return right now."""
        n = 0 if self._left is None else len(self._left)
        for job in self._pending:
            if not job.ready():
                break
            n += self.chunk
        return n

    def take(self, n, block=True, timeout=600.0):
        """The next n scenes of the stream; None if block=False and fewer than
        n are finished.  A chunk that does not arrive within `timeout` seconds
        raises RuntimeError (workers that cannot start -- e.g. a parent whose
        __main__ cannot be re-imported under the spawn start method -- would
        otherwise block forever)."""
        if not block and self.ready() < n:
            return None
        out = []
        have = 0
        while have < n:
            if self._left is None:
                try:
                    self._left = self._pending[0].get(timeout)
                except mp.TimeoutError:
                    raise RuntimeError(
                        'ScenePoolSampler: no scenes from the worker processes '
                        f'within {timeout:.0f} s') from None
                self._pending.popleft()
                self._fill()
            part = self._left[:n - have]
            self._left = self._left[len(part):]
            if len(self._left) == 0:
                self._left = None
            out.append(part)
            have += len(part)
        return np.concatenate(out)

    def close(self):
        if self._pool is not None:
            self._pool.terminate()
            self._pool.join()
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    __del__ = close
