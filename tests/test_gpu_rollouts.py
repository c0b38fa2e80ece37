"""The reference's own test (tests/test_rollout_preproc.py:17-35), through the
drop-in gym surface on the GPU path: every registered env id rolls out
trajectories of exactly `max_episode_steps` steps with observations inside the
declared observation space.  Plus `render(mode='rgb_array')`, which returns
the full-resolution allo/ego dict under every preprocessor
(base_env.py:309-338), checked against the oracle's rasteriser."""
import numpy as np
import pytest

import magical_b200 as magical

pytestmark = pytest.mark.gpu

N_ROLLOUTS = 2

magical.register_envs()


def test_registered_envs():
    assert len(magical.ALL_REGISTERED_ENVS) > 8


@pytest.mark.parametrize('env_name', magical.ALL_REGISTERED_ENVS)
def test_rollouts(built, env_name):
    env = magical.make(env_name)
    try:
        env.seed(7)
        env.action_space.seed(42)
        obs = env.reset()
        assert env.observation_space.contains(obs)
        for _ in range(N_ROLLOUTS):
            done = False
            traj_len = 0
            while not done:
                action = env.action_space.sample()
                obs, rew, done, info = env.step(action)
                traj_len += 1
                if 'DebugReward' not in env_name:   # base_env.py:290: reward is always 0
                    assert rew == 0.0
                assert done or info['eval_score'] == 0.0
            assert traj_len == env.max_episode_steps
            assert 0.0 <= info['eval_score'] <= 1.0
            assert env.observation_space.contains(obs)
            env.reset()
    finally:
        env.close()


@pytest.mark.parametrize('env_name', ['MatchRegions-Demo-LoRes4E-v0',
                                      'ClusterShape-TestAll-LoResStack-v0',
                                      'MoveToCorner-Demo-v0'])
def test_render_rgb_array_is_full_resolution_dict(built, env_name):
    from oracle_lib import OracleEnv
    env = magical.make(env_name)
    try:
        env.seed(3)
        env.action_space.seed(5)
        env.reset()
        actions = [env.action_space.sample() for _ in range(9)]
        for a in actions:
            env.step(a)
        views = env.render(mode='rgb_array')
        assert list(views.keys()) == ['allo', 'ego']
        orc = OracleEnv(env._venv.scenes[0], det_sincos=True)
        for a in actions:
            orc.step(int(a))
        for v, name in enumerate(('allo', 'ego')):
            assert views[name].shape == (384, 384, 3)
            assert views[name].dtype == np.uint8
            assert np.array_equal(views[name], orc.render_view(v)), name
        with pytest.raises(NotImplementedError):
            env.render(mode='human')
    finally:
        env.close()
