"""Demo files in the reference's on-disk format and their batched replay
(magical_b200/saved_trajectories.py, the N4 row of SURVEY §8(f))."""
import gzip
import pickle
import sys
import types

import numpy as np
import pytest

from magical_b200 import saved_trajectories as st


def test_round_trip_and_class_rewrite(tmp_path):
    traj = st.MAGICALTrajectory(acts=np.arange(5), obs={'ego': np.zeros((6, 2))},
                                rews=np.zeros(5), infos=[{'eval_score': 0.0}] * 5)
    path = str(tmp_path / 'demo-a.pkl.gz')
    st.save_demo(path, 'MoveToCorner-Demo-v0', traj, 0.25)
    # a file written by the reference names ITS trajectory class
    fake = types.ModuleType('magical')
    fake_st = types.ModuleType('magical.saved_trajectories')
    cls = type('MAGICALTrajectory', (tuple,), {
        '__module__': 'magical.saved_trajectories',
        '__new__': lambda c, *a: tuple.__new__(c, a),
        '__reduce__': lambda self: (type(self), tuple(self))})
    fake_st.MAGICALTrajectory = cls
    sys.modules['magical'] = fake
    sys.modules['magical.saved_trajectories'] = fake_st
    try:
        path_b = str(tmp_path / 'demo-b.pkl.gz')
        with gzip.GzipFile(path_b, 'wb') as fp:
            pickle.dump({'env_name': 'MoveToCorner-Demo-v0',
                         'trajectory': cls(*traj), 'score': 0.5}, fp)
    finally:
        del sys.modules['magical.saved_trajectories'], sys.modules['magical']
    demos = list(st.load_demos([path, path_b]))
    assert [d['score'] for d in demos] == [0.25, 0.5]
    for d in demos:
        assert isinstance(d['trajectory'], st.MAGICALTrajectory)
        assert np.array_equal(d['trajectory'].acts, np.arange(5))


def _oracle_demo(env_name, actions, with_obs=False):
    """A demonstration recorded by the CPU ORACLE: scripted actions, the oracle's physics and score
    function, and (with_obs) its brute-force 384x384 renders of both views, in the reference's
    trajectory layout (one more observation than actions)."""
    import magical_b200 as magical
    from oracle_lib import OracleEnv
    task, spec = magical.make_task(env_name)
    orc = OracleEnv(task.build_scene(), det_sincos=True)
    obs = {'allo': [], 'ego': []}

    def snap():
        if with_obs:
            obs['allo'].append(orc.render_view(0, 384))
            obs['ego'].append(orc.render_view(1, 384))
    snap()
    for a in actions:
        orc.step(int(a))
        snap()
    score = float(np.float32(orc.score()))
    traj = st.MAGICALTrajectory(
        acts=np.asarray(actions, dtype=np.int64),
        obs={k: np.stack(v) for k, v in obs.items()} if with_obs else {},
        rews=np.zeros(len(actions)), infos=None)
    orc.close()
    return traj, score


def _search_scored_demo(env_name, length, first_seed, max_tries=80):
    """Random demos mostly score 0; look for an action stream on which the ORACLE scores > 0."""
    for k in range(max_tries):
        rng = np.random.RandomState(first_seed + k)
        acts = _pushy(rng, length)
        traj, score = _oracle_demo(env_name, acts)
        if score > 0:
            return traj, score
    raise AssertionError(f'no scoring demo found for {env_name}')


def _pushy(rng, n):
    return [int(rng.randint(18)) if rng.rand() < 0.5 else int(rng.choice([1, 4, 7, 10, 13, 16]))
            for _ in range(n)]


def test_preprocess_demos_matches_oracle_downsample():
    """area_mean_4x4 / preprocess_demos_with_wrapper against the oracle's INTER_AREA statement (itself
    pinned to real cv2 in test_oracle_golden.py) on oracle-rendered frames, all five layouts."""
    from oracle_lib import downsample4
    rng = np.random.RandomState(4)
    traj, _ = _oracle_demo('MoveToCorner-Demo-v0', _pushy(rng, 6), with_obs=True)
    lo = {k: np.stack([downsample4(f) for f in traj.obs[k]]) for k in ('allo', 'ego')}
    assert np.array_equal(st.area_mean_4x4(traj.obs['ego']), lo['ego'])

    def stack(frames, depth, t):
        idx = [max(t - k, 0) for k in range(depth - 1, -1, -1)]
        return np.concatenate([frames[i] for i in idx], axis=-1)

    outs = {p: st.preprocess_demos_with_wrapper([traj], 'MoveToCorner-Demo-v0', preproc_name=p)[0]
            for p in ('LoRes4E', 'LoRes4A', 'LoRes3EA', 'LoResStack', 'LoResCHW4E')}
    for t in range(len(traj.acts) + 1):
        assert np.array_equal(outs['LoRes4E'].obs[t], stack(lo['ego'], 4, t))
        assert np.array_equal(outs['LoRes4A'].obs[t], stack(lo['allo'], 4, t))
        assert np.array_equal(outs['LoRes3EA'].obs[t],
                              np.concatenate([lo['allo'][t], stack(lo['ego'], 3, t)], axis=-1))
        assert np.array_equal(outs['LoResStack'].obs['allo'][t], stack(lo['allo'], 4, t))
        assert np.array_equal(outs['LoResCHW4E'].obs[t], np.moveaxis(stack(lo['ego'], 4, t), -1, 0))
    assert outs['LoRes4E'].obs.shape == (7, 96, 96, 12) and outs['LoResCHW4E'].obs.shape == (7, 12, 96, 96)
    assert st.splice_in_preproc_name('MoveToCorner-Demo-v0', 'LoResStack') == 'MoveToCorner-Demo-LoResStack-v0'
    with pytest.raises(AssertionError):
        st.splice_in_preproc_name('MoveToCorner-Demo-v0', 'NoSuchPreproc')


@pytest.mark.gpu
def test_replay_of_oracle_recorded_demos(built, tmp_path):
    """demonstrations recorded by the ORACLE (its physics, its score functions),
    written in the reference's file format, replayed through the GPU engine in one batch per env id:
    the replayed score equals the recorded one for every demo, and the observations the GPU env returns
    during the replay equal the recorded raw frames pushed through preprocess_demos_with_wrapper."""
    import torch
    import magical_b200 as magical
    rng = np.random.RandomState(12)
    specs = [('MoveToRegion-Demo-v0', 40), ('ClusterColour-Demo-v0', 150), ('MoveToCorner-Demo-v0', 80),
             ('MatchRegions-Demo-v0', 110), ('MoveToRegion-Demo-v0', 25), ('FixColour-Demo-v0', 60),
             ('FindDupe-Demo-v0', 100), ('MakeLine-Demo-v0', 140), ('ClusterShape-Demo-v0', 200)]
    paths, demos = [], []
    for n, (env_name, length) in enumerate(specs):
        traj, score = _oracle_demo(env_name, _pushy(rng, length), with_obs=(n == 2))
        demos.append((env_name, traj, score))
    # demos that SCORE (random ones almost never do): searched with the oracle
    for env_name, length, seed in (('MoveToRegion-Demo-v0', 40, 100), ('FixColour-Demo-v0', 60, 200),
                                   ('MoveToRegion-Demo-v0', 40, 300)):
        traj, score = _search_scored_demo(env_name, length, seed)
        demos.append((env_name, traj, score))
    for n, (env_name, traj, score) in enumerate(demos):
        paths.append(str(tmp_path / f'demo-{n}.pkl.gz'))
        st.save_demo(paths[-1], env_name, traj, score)
    out = st.replay_demos(st.load_demos(paths))
    assert [o['n_actions'] for o in out] == [len(d[1].acts) for d in demos]
    for o in out:
        assert np.float32(o['replayed_score']) == np.float32(o['recorded_score']), o
    assert sum(1 for o in out if o['recorded_score'] > 0) >= 3   # not all trivially zero
    # observations: recorded raw frames -> LoRes4E == what the GPU env shows on the same actions
    env_name, traj, _ = demos[2]
    want = st.preprocess_demos_with_wrapper([traj], env_name, preproc_name='LoRes4E')[0].obs
    venv = magical.make_vec(st.splice_in_preproc_name(env_name, 'LoRes4E'), 1, auto_reset=False)
    assert np.array_equal(venv.reset()[0].cpu().numpy(), want[0])
    for t, a in enumerate(traj.acts[:30]):
        obs = venv.step(torch.tensor([int(a)], dtype=torch.int32, device='cuda'))[0]
        assert np.array_equal(obs[0].cpu().numpy(), want[t + 1]), t
    venv.close()


@pytest.mark.gpu
def test_replay_reproduces_recorded_scores(built, tmp_path):
    """Self-recorded demonstrations (random actions on the GPU engine, written
    in the reference's format) replay to exactly the recorded scores, batched
    per env id, for demos of different lengths."""
    import torch
    import magical_b200 as magical
    rng = np.random.RandomState(8)
    paths = []
    for n, (env_name, length) in enumerate([('MoveToRegion-Demo-v0', 40),
                                            ('MoveToCorner-Demo-v0', 80),
                                            ('MoveToRegion-Demo-v0', 25)]):
        venv = magical.make_vec(env_name.replace('-v0', '-LoRes4E-v0'), 1,
                                auto_reset=False)
        venv.reset()
        acts = rng.randint(0, 18, size=length).astype(np.int64)
        for a in acts:
            venv.step(torch.tensor([int(a)], dtype=torch.int32, device='cuda'))
        score = float(venv.eval_score()[0].item())
        venv.close()
        traj = st.MAGICALTrajectory(acts=acts, obs={}, rews=np.zeros(length),
                                    infos=None)
        paths.append(str(tmp_path / f'demo-{n}.pkl.gz'))
        st.save_demo(paths[-1], env_name, traj, score)
    out = st.replay_demos(st.load_demos(paths))
    assert [o['n_actions'] for o in out] == [40, 80, 25]
    for o in out:
        assert o['replayed_score'] == o['recorded_score'], o
