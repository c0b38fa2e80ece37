"""Batched evaluation protocol (magical_b200/evaluation.py, the N3 row of
SURVEY §8(f); reference magical/evaluation.py:13-98)."""
import numpy as np
import pytest

from magical_b200.evaluation import score_statistics


def test_score_statistics_match_t_interval():
    scores = [0.0, 1.0, 0.25, 0.75, 0.5, 1.0, 0.0, 0.3]
    mean, (lo, hi), std = score_statistics(scores)
    assert mean == pytest.approx(np.mean(scores))
    assert std == pytest.approx(np.std(scores, ddof=1))
    # t_{0.975, 7} = 2.364624...
    half = 2.3646242510102993 * std / np.sqrt(len(scores))
    assert lo == pytest.approx(mean - half) and hi == pytest.approx(mean + half)


@pytest.mark.gpu
def test_batched_protocol_runs_all_variants(built):
    import torch
    import magical_b200 as magical
    from magical_b200.evaluation import BatchedEvaluationProtocol
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1)

    def policy(obs):
        return torch.randint(0, 18, (obs.shape[0],), dtype=torch.int32,
                             device=obs.device, generator=gen)

    proto = BatchedEvaluationProtocol('MoveToRegion-Demo-LoRes4E-v0', 24,
                                      policy, run_id='random')
    frame = proto.do_eval()
    want = ['MoveToRegion-Demo-LoRes4E-v0',
            *magical.DEMO_ENVS_TO_TEST_ENVS_MAP['MoveToRegion-Demo-LoRes4E-v0']]
    assert list(frame['test_env']) == want
    assert list(frame.columns) == ['demo_env', 'test_env', 'mean_score',
                                   'ci95_lower', 'ci95_upper', 'std_score',
                                   'run_id']
    assert ((frame['mean_score'] >= 0) & (frame['mean_score'] <= 1)).all()


def test_protocol_contract_and_latex_table():
    """The reference's subclassing contract (evaluation.py:13-98): obtain_scores is the user's hook;
    too few scores raise, extra ones are dropped with a warning; latexify_results renders one row per
    run id."""
    import warnings
    from magical_b200.evaluation import EvaluationProtocol, latexify_results

    class Fixed(EvaluationProtocol):
        run_id = 'fixed'

        def __init__(self, n, k):
            super().__init__('MoveToRegion-Demo-v0', n)
            self.k = k

        def obtain_scores(self, env_name):
            return np.linspace(0.0, 1.0, self.k)

    proto = Fixed(4, 4)
    frame = proto.do_eval()
    n_envs = len(proto.test_env_names)
    assert n_envs >= 4 and len(frame) == n_envs and (frame['run_id'] == 'fixed').all()
    assert frame['mean_score'].iloc[0] == pytest.approx(0.5)
    with pytest.raises(ValueError, match='returned only 3 scores'):
        Fixed(4, 3).do_eval()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        frame = Fixed(4, 6).do_eval()
    assert w and 'truncate' in str(w[0].message)
    assert frame['mean_score'].iloc[0] == pytest.approx(np.mean(np.linspace(0, 1, 6)[:4]))
    tex = latexify_results(frame)
    assert tex.count('\\textbf{MoveToRegion') == n_envs and '\\textbf{fixed} & 0.30 ($\\pm$ 0.26)' in tex


@pytest.mark.gpu
def test_batched_protocol_scores_equal_oracle_episodes(built):
    """with a fixed scripted policy the per-episode scores the protocol collects
    on the GPU equal the scores of the same episodes stepped by the ORACLE, for the Demo variant and the
    randomised variants (each rollout on its own sampled layout); mean / CI / std then follow."""
    import torch
    from magical_b200.evaluation import BatchedEvaluationProtocol, score_statistics
    from oracle_lib import OracleEnv
    n = 12
    rng = np.random.RandomState(5)
    script = np.array([[int(rng.randint(18)) if rng.rand() < 0.5 else int(rng.choice([1, 4, 7, 10, 13, 16]))
                        for _ in range(n)] for _ in range(120)], dtype=np.int32)    # [T, n]
    clock = {'t': 0}

    def policy(obs):
        a = torch.from_numpy(script[clock['t']]).to(obs.device)
        clock['t'] += 1
        return a

    proto = BatchedEvaluationProtocol('MatchRegions-Demo-LoRes4E-v0', n, policy, run_id='scripted', seed=3)
    rows = {}
    for env_name in proto.test_env_names:
        clock['t'] = 0
        scores = proto.obtain_scores(env_name)
        scenes = proto.scenes[env_name]
        want = []
        for e in range(n):
            orc = OracleEnv(scenes[e % len(scenes)], det_sincos=True)
            for t in range(120):
                _, done, sc_ = orc.step(int(script[t, e]))
            assert done
            want.append(np.float32(sc_))
            orc.close()
        assert np.array_equal(scores.astype(np.float32), np.array(want)), env_name
        rows[env_name] = score_statistics(scores)
    clock['t'] = 0

    def policy2(obs):
        a = torch.from_numpy(script[clock['t'] % 120]).to(obs.device)
        clock['t'] += 1
        return a
    proto.policy = policy2
    frame = proto.do_eval()
    for _, r in frame.iterrows():
        assert r['mean_score'] == pytest.approx(rows[r['test_env']][0])
