#!/usr/bin/env python3
"""How far do end-of-episode SCORES move when the arbiter (contact pair) order changes?

The one place where the oracle knowingly departs from Chipmunk is the order in which colliding shape pairs
reach the solver: canonical pair list here, cpBBTree traversal order there (unknowable offline).  Poses
diverge under a different order (zero-length finger PinJoints amplify rounding), so
pose-level agreement with pymunk is ill-posed; what a user of the benchmark sees is the score.  This tool
plays N push-biased random episodes per Demo task twice -- canonical order and a random permutation of the
pair list (a fresh permutation per episode) -- with identical actions and reports the distribution of
score differences.  CPU only (oracle = test infrastructure).

    python tools/score_sensitivity.py [episodes_per_task] [out.json]
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

TASKS = ['MoveToCorner', 'MoveToRegion', 'MatchRegions', 'MakeLine', 'FindDupe', 'FixColour', 'ClusterColour',
         'ClusterShape']


def episode_pair(job):
    """(task name, seed) -> (canonical score, permuted score, max final pose distance)."""
    name, seed = job
    import magical_b200 as magical
    from oracle_lib import OracleEnv
    task, spec = magical.make_task(f'{name}-Demo-v0')
    rec = task.build_scene()
    rng = np.random.RandomState(seed)
    n = spec.max_episode_steps
    acts = [int(rng.randint(18)) if rng.rand() < 0.5 else int(rng.choice([1, 4, 7, 10, 13, 16])) for _ in range(n)]
    perm = rng.permutation(int(rec['n_bpairs'])).astype(np.int32)
    out = []
    for p in (None, perm):
        orc = OracleEnv(rec, det_sincos=True)
        orc.set_pair_permutation(p)
        score = None
        for a in acts:
            _, done, s = orc.step(a)
            if done:
                score = s
        st = orc.state()
        out.append((float(score), st['pos'][:int(st['n_bodies'])].copy()))
        orc.close()
    return out[0][0], out[1][0], float(np.abs(out[0][1] - out[1][1]).max())


def sweep(n_episodes, tasks=TASKS, workers=None):
    import oracle_lib
    oracle_lib.lib()   # build once before the workers start
    jobs = [(t, 1000 * k + i) for k, t in enumerate(tasks) for i in range(n_episodes)]
    with mp.get_context('spawn').Pool(workers or os.cpu_count()) as pool:
        res = pool.map(episode_pair, jobs, chunksize=4)
    report = {}
    for k, t in enumerate(tasks):
        r = np.array(res[k * n_episodes:(k + 1) * n_episodes])
        d = r[:, 1] - r[:, 0]
        report[t] = {
            'episodes': n_episodes,
            'mean_score_canonical': float(r[:, 0].mean()), 'mean_score_permuted': float(r[:, 1].mean()),
            'mean_delta': float(d.mean()), 'mean_abs_delta': float(np.abs(d).mean()),
            'frac_episodes_score_changed': float((d != 0).mean()),
            'max_abs_delta': float(np.abs(d).max()),
            'abs_delta_quantiles_50_90_99': [float(q) for q in np.quantile(np.abs(d), [0.5, 0.9, 0.99])],
            'stderr_of_mean_delta': float(d.std(ddof=1) / np.sqrt(len(d))) if len(d) > 1 else None,
            'median_final_pose_distance': float(np.median(r[:, 2])),
            'frac_episodes_pose_changed': float((r[:, 2] > 0).mean()),
        }
    return report


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rep = sweep(n)
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'profiles', 'r02_score_sensitivity.json')
    with open(out, 'w') as fh:
        json.dump({'what': 'score deltas, random pair-order permutation vs canonical, push-biased random episodes',
                   'tasks': rep}, fh, indent=1)
    for t, r in rep.items():
        print(f"{t:14s} changed {100 * r['frac_episodes_score_changed']:5.1f}% of episodes, mean delta "
              f"{r['mean_delta']:+.4f} (+- {r['stderr_of_mean_delta']:.4f}), mean |delta| {r['mean_abs_delta']:.4f}, "
              f"max |delta| {r['max_abs_delta']:.3f}, poses changed in {100 * r['frac_episodes_pose_changed']:.0f}%")
