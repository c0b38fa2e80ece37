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
