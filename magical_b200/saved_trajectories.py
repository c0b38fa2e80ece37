"""Recorded-demonstration files and their replay.

The reference stores human demonstrations as gzip'd pickles of
`{'env_name': str, 'trajectory': MAGICALTrajectory(acts, obs, rews, infos),
'score': float}` (magical/__main__.py:150-159) and loads them with a class-
rewriting unpickler (magical/saved_trajectories.py:14-49).  This module reads
and writes the same files and, because the Demo variants are deterministic
(misc/re_record_demos.py:29-30 relies on that), replays the recorded actions
through the GPU engine -- all demonstrations of one env id as ONE batch -- so a
recorded `score` can be compared with the score this implementation reaches on
the same actions.  With the public magical-data demos that is a ground-truth
check of the whole physics + scoring path; none are available offline
(reference_demos.py:13 downloads them), so the tests use self-recorded files.
"""
import collections
import gzip
import pickle
from typing import List, NamedTuple, Optional

import numpy as np


class MAGICALTrajectory(NamedTuple):
    """Same fields as the reference's class (saved_trajectories.py:14-21)."""
    acts: np.ndarray
    obs: dict
    rews: np.ndarray
    infos: Optional[List[dict]]


_TRAJ_CLASSES = {
    ('magical.saved_trajectories', 'MAGICALTrajectory'),
    ('magical_b200.saved_trajectories', 'MAGICALTrajectory'),
    ('imitation.util.rollout', 'Trajectory'),
    ('milbench.baselines.saved_trajectories', 'MILBenchTrajectory'),
}


class _TrajRewriteUnpickler(pickle.Unpickler):
    """Maps every historical trajectory class onto MAGICALTrajectory so the
    files load without the packages that wrote them."""

    def find_class(self, module, name):
        if (module, name) in _TRAJ_CLASSES:
            return MAGICALTrajectory
        return super().find_class(module, name)


def load_demos(demo_paths, verbose=False):
    """Yield the demo dictionaries of a sequence of `.pkl.gz` paths."""
    n_demos = len(demo_paths)
    for d_num, d_path in enumerate(demo_paths, start=1):
        if verbose:
            print(f"Loading '{d_path}' ({d_num}/{n_demos})")
        with gzip.GzipFile(d_path, 'rb') as fp:
            yield _TrajRewriteUnpickler(fp).load()


def save_demo(path, env_name, trajectory, score):
    """Write one demonstration in the reference's on-disk format."""
    with gzip.GzipFile(path, 'wb') as fp:
        pickle.dump({'env_name': env_name, 'trajectory': trajectory,
                     'score': float(score)}, fp)


def replay_demos(demo_dicts, device=0, preproc='LoRes4E'):
    """Replay the recorded actions of every demo on the GPU engine.

    Demos are grouped by env id; each group runs as one batch (one environment
    per demo, shorter demos idle after their last action).  Returns a list of
    dicts `{env_name, recorded_score, replayed_score, n_actions}` in input
    order."""
    import torch
    from magical_b200 import benchmarks
    from magical_b200.env import make_vec
    benchmarks.register_envs()
    demo_dicts = list(demo_dicts)
    groups = collections.OrderedDict()
    for idx, demo in enumerate(demo_dicts):
        groups.setdefault(demo['env_name'], []).append(idx)
    results = [None] * len(demo_dicts)
    for env_name, members in groups.items():
        ename = benchmarks.EnvName(env_name)
        if ename.is_test:
            raise ValueError(
                f"'{env_name}' is a randomised variant: its layout is not "
                "recoverable from the recorded actions, only Demo variants "
                "replay deterministically")
        run_name = env_name if ename.preproc is not None else \
            benchmarks.update_magical_env_name(env_name, preproc=preproc)
        acts = [np.asarray(demo_dicts[i]['trajectory'].acts).astype(np.int32)
                .reshape(-1) for i in members]
        horizon = max(len(a) for a in acts)
        venv = make_vec(run_name, len(members), device=device,
                        auto_reset=False)
        try:
            venv.reset()
            scores = np.zeros(len(members), dtype=np.float64)
            for t in range(horizon):
                step_acts = np.asarray([a[t] if t < len(a) else 0
                                        for a in acts], dtype=np.int32)
                _, _, _, info = venv.step(torch.from_numpy(step_acts).to(
                    venv.device))
                # a demo's score is the eval_score of its LAST recorded step
                last = np.asarray([t == len(a) - 1 for a in acts])
                if last.any():
                    now = venv.eval_score().cpu().numpy()
                    scores[last] = now[last]
            for k, i in enumerate(members):
                results[i] = {'env_name': env_name,
                              'recorded_score': float(demo_dicts[i]['score']),
                              'replayed_score': float(scores[k]),
                              'n_actions': int(len(acts[k]))}
        finally:
            venv.close()
    return results
