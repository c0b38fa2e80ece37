#!/usr/bin/env python3
"""The TRUE CPU baseline: env-steps/s of the real reference (qxcv/magical on pymunk 5.6 / pyglet 1.5 / gym 0.17),
measured the way its own `magical/misc/benchmark_env_perf.py:12-18` rolls out (reset, random actions until done),
without the cProfile wrapper, in one process per core.  Prints one JSON line in bench.py's vocabulary.

Neither the build container nor the GPU box of this project has those packages, so there
`bench.py --impl reference` times the CPU oracle instead; this script exists for machines that do have them:

    xvfb-run -a python tools/bench_reference_real.py ClusterColour-Demo-LoRes4E-v0 --seconds 20 --procs 16

Without the reference stack it prints {"impl": "reference-real", "unavailable": "..."} and exits 0.
"""
import argparse
import json
import multiprocessing as mp
import os
import time


def _rollout(args):
    env_id, seed, seconds = args
    import gym
    import magical
    magical.register_envs()
    env = gym.make(env_id)
    env.seed(seed)
    env.action_space.seed(seed)
    steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.reset()
        done = False
        while not done:
            _, _, done, _ = env.step(env.action_space.sample())
            steps += 1
    dt = time.perf_counter() - t0
    env.close()
    return steps, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('env_id')
    ap.add_argument('--seconds', type=float, default=20.0, help='wall-clock budget per process')
    ap.add_argument('--procs', type=int, default=os.cpu_count() or 1)
    ap.add_argument('--seed', type=int, default=42)
    a = ap.parse_args()
    try:
        import gym  # noqa: F401
        import magical  # noqa: F401
        import pymunk  # noqa: F401
    except Exception as ex:  # noqa: BLE001
        print(json.dumps({'impl': 'reference-real', 'unavailable': f'{type(ex).__name__}: {ex}'}))
        return
    jobs = [(a.env_id, a.seed + i, a.seconds) for i in range(a.procs)]
    if a.procs == 1:
        res = [_rollout(jobs[0])]
    else:
        with mp.get_context('spawn').Pool(a.procs) as pool:     # one GL context per process, never forked
            res = pool.map(_rollout, jobs)
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    print(json.dumps({'impl': 'reference-real', 'metric': 'env_steps_per_sec', 'unit': 'env-steps/s',
                      'value': steps / wall, 'cores': a.procs, 'steps': steps, 'seconds': wall,
                      'config': {'workload': a.env_id}, 'higher_is_better': True}))


if __name__ == '__main__':
    main()
