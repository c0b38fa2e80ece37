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
