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


def load_demos(demo_paths, rewrite_traj_cls=True, verbose=False):
    """Yield the demo dictionaries of a sequence of `.pkl.gz` paths.  With
    `rewrite_traj_cls` (the default, as in the reference) every historical
    trajectory class is mapped onto `MAGICALTrajectory`."""
    n_demos = len(demo_paths)
    for d_num, d_path in enumerate(demo_paths, start=1):
        if verbose:
            print(f"Loading '{d_path}' ({d_num}/{n_demos})")
        with gzip.GzipFile(d_path, 'rb') as fp:
            unpickler = _TrajRewriteUnpickler(fp) if rewrite_traj_cls \
                else pickle.Unpickler(fp)
            yield unpickler.load()


def splice_in_preproc_name(base_env_name, preproc_name):
    """'MoveToCorner-Demo-v0' + 'LoResStack' -> 'MoveToCorner-Demo-LoResStack-v0'
    (reference saved_trajectories.py:53-62); the preprocessor must exist."""
    from magical_b200 import benchmarks
    assert preproc_name in benchmarks.DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS, \
        f"no preprocessor named '{preproc_name}', options are " \
        f"{', '.join(benchmarks.DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS)}"
    return benchmarks.update_magical_env_name(base_env_name,
                                              preproc=preproc_name)


def area_mean_4x4(frames):
    """cv2.resize(..., INTER_AREA) for an exact 4x reduction: the mean of each
    4x4 block per channel, rounded half to even (what the LoRes preprocessors
    apply, benchmarks/__init__.py:159-169, 234).  frames: u8 [..., H, W, C]
    with H, W multiples of 4."""
    frames = np.asarray(frames)
    *lead, h, w, c = frames.shape
    s = frames.reshape(*lead, h // 4, 4, w // 4, 4, c).astype(np.uint32).sum(
        axis=(-4, -2))
    return ((s + 7 + ((s >> 4) & 1)) >> 4).astype(np.uint8)


def _stack_last(frames, depth):
    """[T, H, W, C] -> [T, H, W, depth * C]: at time t the `depth` most recent
    frames, oldest first, the first frame repeated at the start (the deque of
    FlattenFrameStack / EagerDictFrameStack, benchmarks/__init__.py:46-136)."""
    frames = np.asarray(frames)
    padded = np.concatenate([np.repeat(frames[:1], depth - 1, axis=0), frames])
    return np.concatenate([padded[k:k + len(frames)] for k in range(depth)],
                          axis=-1)


def preprocess_demos_with_wrapper(trajectories, orig_env_name,
                                  preproc_name=None, wrapper=None):
    """Re-preprocess recorded raw trajectories (dict observations with the
    full-resolution 'allo' and 'ego' views, one more observation than actions)
    with one of the built-in LoRes pipelines; returns trajectories of the same
    type whose `obs` is what the `<env>-<preproc>-v0` id would have returned
    during that episode (reference saved_trajectories.py:90-160, which replays
    the recorded frames through the gym wrappers; here the same arithmetic is
    applied to the whole episode at once).  Custom `wrapper` callables are a
    gym-wrapper mechanism and are not supported."""
    from magical_b200 import benchmarks
    if wrapper is not None:
        raise NotImplementedError(
            'only the built-in preprocessors (preproc_name=...) are supported')
    assert preproc_name in benchmarks.DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS, \
        preproc_name
    benchmarks.register_envs()
    if orig_env_name not in benchmarks.ENV_SPECS:
        raise KeyError(f"no registered MAGICAL env with id '{orig_env_name}'")
    out = []
    for traj in trajectories:
        if isinstance(traj.obs, dict):
            allo, ego = np.asarray(traj.obs['allo']), np.asarray(traj.obs['ego'])
        else:   # a sequence of per-step dicts
            allo = np.stack([o['allo'] for o in traj.obs])
            ego = np.stack([o['ego'] for o in traj.obs])
        lo_a, lo_e = area_mean_4x4(allo), area_mean_4x4(ego)
        if preproc_name == 'LoResStack':
            obs = collections.OrderedDict([('allo', _stack_last(lo_a, 4)),
                                           ('ego', _stack_last(lo_e, 4))])
        elif preproc_name == 'LoRes4A':
            obs = _stack_last(lo_a, 4)
        elif preproc_name == 'LoRes3EA':
            obs = np.concatenate([lo_a, _stack_last(lo_e, 3)], axis=-1)
        else:
            obs = _stack_last(lo_e, 4)
            if preproc_name == 'LoResCHW4E':
                obs = np.moveaxis(obs, -1, 1)
        out.append(type(traj)(acts=np.asarray(traj.acts), obs=obs,
                              rews=np.asarray(traj.rews), infos=traj.infos))
    return out


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
