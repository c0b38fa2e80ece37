"""Many-episode sweep of the PRODUCT's physics source (host build of
mg_physics_tpe.h) against the oracle on the CPU: the first 128 of the 1000
push-biased ClusterShape episodes of the GPU score sweep
(tests/test_gpu_fullsize.py).  Episode 14 of that stream piles up more than 16
simultaneous contacts, which overflowed the kernel's first contact capacity --
long contact-rich rollouts find what 200-step parity runs do not."""
import multiprocessing as mp
import os

import numpy as np

from magical_b200.env import make_task


def test_contact_rich_episodes_bit_exact_and_no_overflow():
    import oracle_lib
    import tpe_host_lib
    oracle_lib.lib()
    tpe_host_lib.lib()          # build once here, not in every worker
    task, _ = make_task('ClusterShape-Demo-LoRes4E-v0')
    scene = task.build_scene()
    n_steps, total, n = task.max_episode_steps, 1000, 128
    # the same action streams as test_score_sweep_1000_episodes
    rng = np.random.RandomState(101)
    base = rng.randint(0, 18, size=(n_steps, total)).astype(np.int32)
    push = rng.choice([1, 4, 7, 10, 13, 16], size=(n_steps, total))
    acts = np.where(rng.rand(n_steps, total) < 0.5, base, push).astype(np.int32)[:, :n]
    chunk = 8
    jobs = [([scene] * chunk, acts[:, lo:lo + chunk]) for lo in range(0, n, chunk)]
    with mp.get_context('spawn').Pool(min(os.cpu_count() or 1, 16)) as pool:
        res = pool.map(tpe_host_lib.rollout_mismatches, jobs)
    bad = [(i, b) for i, b in enumerate(x for r in res for x in r) if b]
    assert not bad, bad[:5]


def test_randomised_layouts_bit_exact_and_no_overflow():
    """The same check over sampled layouts of three randomised variants
    (64 layouts each, one push-biased episode per layout)."""
    import oracle_lib
    import tpe_host_lib
    oracle_lib.lib()
    tpe_host_lib.lib()
    jobs = []
    for env_id in ('ClusterColour-TestAll-LoRes4E-v0',
                   'MatchRegions-TestAll-LoRes4E-v0',
                   'MakeLine-TestCountPlus-LoRes4E-v0'):
        task, _ = make_task(env_id)
        task.seed(31)
        n, n_steps = 64, task.max_episode_steps
        scenes = [task.build_scene() for _ in range(n)]
        rng = np.random.RandomState(7)
        base = rng.randint(0, 18, size=(n_steps, n)).astype(np.int32)
        push = rng.choice([1, 4, 7, 10, 13, 16], size=(n_steps, n))
        acts = np.where(rng.rand(n_steps, n) < 0.5, base, push).astype(np.int32)
        jobs += [(scenes[lo:lo + 8], acts[:, lo:lo + 8]) for lo in range(0, n, 8)]
    with mp.get_context('spawn').Pool(min(os.cpu_count() or 1, 16)) as pool:
        res = pool.map(tpe_host_lib.rollout_mismatches, jobs)
    bad = [(i, b) for i, b in enumerate(x for r in res for x in r) if b]
    assert not bad, bad[:5]
