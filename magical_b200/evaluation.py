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
import abc
import collections
import io
import warnings

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


class EvaluationProtocol(abc.ABC):
    """The reference's subclassing contract (magical/evaluation.py:13-98):
    a subclass provides `run_id` and `obtain_scores(env_name)` (at least
    `n_rollouts` eval_scores of rollouts on the named env, extra ones are
    dropped with a warning); `do_eval()` evaluates the Demo variant and every
    registered Test variant and returns one DataFrame row per variant."""
    _called_init = False

    def __init__(self, demo_env_name, n_rollouts):
        benchmarks.register_envs()
        self.n_rollouts = int(n_rollouts)
        self.demo_env_name = demo_env_name
        self.test_env_names = [
            demo_env_name,
            *benchmarks.DEMO_ENVS_TO_TEST_ENVS_MAP[demo_env_name]]
        self._called_init = True

    @property
    @abc.abstractmethod
    def run_id(self):
        """String naming the evaluated model / algorithm (`run_id` column)."""

    @abc.abstractmethod
    def obtain_scores(self, env_name):
        """`self.n_rollouts` (or more) eval_scores on the env `env_name`."""

    def do_eval(self, verbose=False):
        import pandas as pd
        if not self._called_init:
            raise ValueError(
                "EvaluationProtocol.__init__() was not called. Did you "
                "include a super().__init__(...) call in your subclass?")
        records = []
        for env_name in self.test_env_names:
            scores = self.obtain_scores(env_name)
            if len(scores) < self.n_rollouts:
                raise ValueError(
                    f".obtain_scores() returned only {len(scores)} scores, "
                    f"but we asked for {self.n_rollouts} scores")
            if len(scores) > self.n_rollouts:
                warnings.warn(
                    f"Asked for {self.n_rollouts} scores but got "
                    f"{len(scores)} scores instead. Will truncate to only "
                    f"consider the first {self.n_rollouts} scores.")
                scores = scores[:self.n_rollouts]
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


class BatchedEvaluationProtocol(EvaluationProtocol):
    """policy(obs) -> int actions [n_rollouts] (torch tensors on the env's
    device; `obs` is the batched observation of the chosen preprocessor)."""

    run_id = None  # set per instance

    def __init__(self, demo_env_name, n_rollouts, policy, run_id='policy',
                 device=0, seed=0):
        super().__init__(demo_env_name, n_rollouts)
        self.policy = policy
        self.run_id = run_id
        self.device = device
        self.seed = seed
        # compiled scenes each variant's rollouts ran on (rollout i played
        # scene i % len), kept so that a run can be reproduced / audited
        self.scenes = {}

    def obtain_scores(self, env_name):
        """One batch of `n_rollouts` full episodes; returns their eval_scores."""
        import torch
        is_test = benchmarks.EnvName(env_name).is_test
        venv = make_vec(env_name, self.n_rollouts, device=self.device,
                        auto_reset=False, seed=self.seed,
                        n_scenes=self.n_rollouts if is_test else 1)
        self.scenes[env_name] = venv.scenes.copy()
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



def latexify_results(eval_data, id_column='run_id'):
    """LaTeX table from the frame(s) `do_eval()` returns: one column per test
    env, one row per algorithm (distinct value of `id_column`), each cell
    `mean ($\\pm$ std)`; same layout as the reference's `latexify_results`
    (magical/evaluation.py:101-153)."""
    test_envs = eval_data['test_env'].unique()
    headers = [r'\textbf{%s}' % e for e in test_envs]
    out = io.StringIO()
    print(r'\centering', file=out)
    print(r'\begin{tabular}{l@{\hspace{1em}}%s}' % ('c' * len(headers)), file=out)
    print(r'\toprule', file=out)
    print(r'\textbf{Randomisation} & ' + ' & '.join(headers) + '\\\\', file=out)
    print(r'\midrule', file=out)
    for alg in eval_data[id_column].unique():
        cells = []
        for env_name in test_envs:
            rows = eval_data[(eval_data[id_column] == alg)
                             & (eval_data['test_env'] == env_name)]
            if len(rows) != 1:
                raise ValueError(
                    f"got {len(rows)} rows corresponding to {id_column}={alg} "
                    f"and test_env={env_name}, but expected one (maybe IDs in "
                    f"column {id_column} aren't unique?)")
            row = rows.iloc[0]
            cells.append(f'{row["mean_score"]:.2f} ($\\pm$ {row["std_score"]:.2f})')
        print(r'\textbf{%s} & ' % alg + ' & '.join(cells) + '\\\\', file=out)
        print(r'\bottomrule', file=out)
        print(r'\end{tabular}', file=out)
    return out.getvalue()
