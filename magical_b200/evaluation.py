"""Batched multi-task evaluation protocol.

Counterpart of the reference's `EvaluationProtocol` (magical/evaluation.py:13-98):
for a Demo env it evaluates the Demo variant and every registered Test variant
(`DEMO_ENVS_TO_TEST_ENVS_MAP`, benchmarks/__init__.py:1001-1019) with
`n_rollouts` rollouts each and reports mean score, 95 % two-sided t confidence
interval and sample standard deviation per variant, in the same column layout.
The difference is the rollout engine: the `n_rollouts` episodes of a variant
run as ONE batch on the GPU (each on its own freshly sampled scene for the
randomised variants), so the user supplies a batched policy instead of
`obtain_scores`.
"""
import collections

import numpy as np

from magical_b200 import benchmarks
from magical_b200.env import make_vec


def score_statistics(scores):
    """mean, (ci95_lower, ci95_upper), std(ddof=1) as the reference computes
    them (statsmodels DescrStatsW.tconfint_mean(0.05) == mean -+ t * sem)."""
    from scipy import stats
    scores = np.asarray(scores, dtype=np.float64)
    n = len(scores)
    mean = float(np.mean(scores))
    std = float(np.std(scores, ddof=1)) if n > 1 else float('nan')
    if n > 1:
        half = float(stats.t.ppf(0.975, n - 1)) * std / np.sqrt(n)
    else:
        half = float('nan')
    return mean, (mean - half, mean + half), std


class BatchedEvaluationProtocol:
    """policy(obs) -> int actions [n_rollouts] (torch tensors on the env's
    device; `obs` is the batched observation of the chosen preprocessor)."""

    def __init__(self, demo_env_name, n_rollouts, policy, run_id='policy',
                 device=0, seed=0):
        benchmarks.register_envs()
        self.demo_env_name = demo_env_name
        self.n_rollouts = int(n_rollouts)
        self.policy = policy
        self.run_id = run_id
        self.device = device
        self.seed = seed
        self.test_env_names = [
            demo_env_name,
            *benchmarks.DEMO_ENVS_TO_TEST_ENVS_MAP[demo_env_name]]

    def obtain_scores(self, env_name):
        """One batch of `n_rollouts` full episodes; returns their eval_scores."""
        import torch
        is_test = benchmarks.EnvName(env_name).is_test
        venv = make_vec(env_name, self.n_rollouts, device=self.device,
                        auto_reset=False, seed=self.seed,
                        n_scenes=self.n_rollouts if is_test else 1)
        try:
            scene_ids = np.arange(self.n_rollouts) % venv.n_scenes
            obs = venv.reset(scene_ids=scene_ids if venv.n_scenes > 1 else None)
            scores = None
            for _ in range(venv.max_episode_steps):
                actions = self.policy(obs)
                obs, _, done, info = venv.step(actions)
            assert bool(torch.all(done != 0)), 'episodes did not end together'
            scores = info['eval_score'].cpu().numpy().astype(np.float64)
        finally:
            venv.close()
        return scores

    def do_eval(self, verbose=False):
        import pandas as pd
        records = []
        for env_name in self.test_env_names:
            scores = self.obtain_scores(env_name)[:self.n_rollouts]
            mean, (lo, hi), std = score_statistics(scores)
            records.append(collections.OrderedDict([
                ('demo_env', self.demo_env_name), ('test_env', env_name),
                ('mean_score', mean), ('ci95_lower', lo), ('ci95_upper', hi),
                ('std_score', std), ('run_id', self.run_id)]))
        frame = pd.DataFrame.from_records(records)
        if verbose:
            print(f"Final mean scores for '{self.run_id}':")
            print(frame[['test_env', 'mean_score', 'ci95_lower', 'ci95_upper']])
        return frame
