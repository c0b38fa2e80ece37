"""BASELINE.json's full batch sizes through size-independent properties (the
oracle cannot step 65 536 environments): environments that receive the same
actions stay bit-identical across the whole batch (no cross-environment
interference at full occupancy), two half-size handles equal one full-size
handle (the 1-GPU == N-GPU shard property, on one device), and a sampled
environment still equals the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_pool():
    """Worker processes for the oracle side of the sweeps.  Spawned, not
    forked: this process holds a CUDA context and its threads, and the
    workers only need numpy + the oracle's ctypes binding."""
    import multiprocessing as mp
    import os
    import oracle_lib
    oracle_lib.lib()      # build once here, not concurrently in every worker
    return mp.get_context('spawn').Pool(min(os.cpu_count() or 1, 32))


def _actions(rng, n_steps, batch, n_groups):
    """n_groups distinct action streams, env i follows stream i % n_groups."""
    base = rng.randint(0, 18, size=(n_steps, n_groups)).astype(np.int32)
    # forward-biased so that contacts happen
    push = rng.choice([1, 4, 7, 10, 13, 16], size=(n_steps, n_groups))
    base = np.where(rng.rand(n_steps, n_groups) < 0.5, base, push).astype(np.int32)
    return base[:, np.arange(batch) % n_groups]


@pytest.mark.parametrize('env_id,batch', [
    ('ClusterColour-Demo-LoRes4E-v0', 65536),   # BASELINE configs[2]
    ('MoveToCorner-Demo-LoRes4E-v0', 4096),     # BASELINE configs[1]
])
def test_full_batch_replicas_stay_identical_and_match_oracle(built, env_id,
                                                             batch):
    import torch
    import magical_b200 as magical
    from oracle_lib import OracleEnv
    n_groups, n_steps = 7, 40
    rng = np.random.RandomState(9)
    acts = _actions(rng, n_steps, batch, n_groups)
    venv = magical.make_vec(env_id, batch, auto_reset=False)
    venv.reset()
    orc = OracleEnv(venv.scenes[0], det_sincos=True)
    probe = batch - 3                       # an env in the last warp / block
    for t in range(n_steps):
        obs, rew, done, info = venv.step(torch.from_numpy(acts[t]).cuda())
        orc.step(int(acts[t, probe]))
    # replicas: every env of a group equals the group's first env, pixels included
    groups = torch.arange(batch, device=obs.device) % n_groups
    first = obs[:n_groups]
    assert torch.equal(obs, first[groups])
    st, ost = venv.get_state(probe), orc.state()
    nb = int(st['n_bodies'])
    assert int(st['overflow']) == 0
    assert np.array_equal(st['pos'][:nb], ost['pos'][:nb])
    assert np.array_equal(st['angle'][:nb], ost['angle'][:nb])
    assert np.array_equal(obs[probe, :, :, 9:12].cpu().numpy(),
                          orc.render_lores(1))
    for e in (n_groups + (probe % n_groups), batch // 2 + (probe - batch // 2) % n_groups):
        if e % n_groups == probe % n_groups:
            assert np.array_equal(venv.get_state(e)['pos'][:nb], st['pos'][:nb])
    venv.close()


def test_two_half_handles_equal_one_full_handle(built):
    """Shard equivalence: environments are independent, so splitting the batch
    over handles (= over GPUs, magical_b200/dist.py) must not change anything."""
    import torch
    import magical_b200 as magical
    env_id, batch, n_steps = 'MatchRegions-Demo-LoRes4E-v0', 8192, 30
    rng = np.random.RandomState(4)
    acts = _actions(rng, n_steps, batch, 11)
    full = magical.make_vec(env_id, batch, auto_reset=True)
    halves = [magical.make_vec(env_id, batch // 2, auto_reset=True)
              for _ in range(2)]
    full.reset()
    for h in halves:
        h.reset()
    for t in range(n_steps):
        a = torch.from_numpy(acts[t]).cuda()
        obs, rew, done, info = full.step(a)
        parts = [h.step(a[i * (batch // 2):(i + 1) * (batch // 2)])
                 for i, h in enumerate(halves)]
    assert torch.equal(obs, torch.cat([p[0] for p in parts]))
    assert torch.equal(done, torch.cat([p[2] for p in parts]))
    assert torch.equal(info['eval_score'],
                       torch.cat([p[3]['eval_score'] for p in parts]))
    for v in (full, *halves):
        v.close()


SWEEP = [(name + '-Demo-LoRes4E-v0', 0) for name in (
    'MoveToCorner', 'MoveToRegion', 'MatchRegions', 'MakeLine', 'FindDupe',
    'FixColour', 'ClusterColour', 'ClusterShape')] + [
    ('MatchRegions-TestAll-LoRes4E-v0', 250),
    ('ClusterColour-TestAll-LoRes4E-v0', 125),
]


@pytest.mark.parametrize('env_id,n_scenes', SWEEP)
def test_score_sweep_1000_episodes(built, env_id, n_scenes):
    """end-of-episode scores, done steps and final poses of
    1000 random-action episodes per task, CUDA against the oracle (stepped on
    all host cores), exact.  The two randomised variants run 1000 episodes
    over a pool of distinct sampled layouts."""
    import multiprocessing as mp
    import os
    import torch
    import magical_b200 as magical
    from oracle_lib import rollout_scores
    batch = 1000
    kw = dict(n_scenes=n_scenes, seed=5) if n_scenes else {}
    venv = magical.make_vec(env_id, batch, auto_reset=False, **kw)
    scene_ids = np.arange(batch) % max(n_scenes, 1)
    venv.reset(scene_ids=scene_ids) if n_scenes else venv.reset()
    n_steps = venv.max_episode_steps
    rng = np.random.RandomState(101)
    acts = _actions(rng, n_steps, batch, batch)
    # the oracle side runs on all host cores
    chunk = 25
    jobs = [([venv.scenes[scene_ids[e]] for e in range(lo, lo + chunk)],
             acts[:, lo:lo + chunk]) for lo in range(0, batch, chunk)]
    with _oracle_pool() as pool:
        res = pool.map(rollout_scores, jobs)
    o_score = np.concatenate([r[0] for r in res])
    o_done = np.concatenate([r[1] for r in res])
    o_pos = np.concatenate([r[2] for r in res])
    score = np.full(batch, np.nan, dtype=np.float32)
    done_at = np.full(batch, -1, dtype=np.int32)
    for t in range(n_steps):
        rew, done, info = venv.step_physics(torch.from_numpy(acts[t]).cuda())
        d = done.cpu().numpy().astype(bool) & (done_at < 0)
        done_at[d] = t
        score[d] = info['eval_score'].cpu().numpy()[d]
    assert np.array_equal(done_at, o_done)
    assert (done_at == n_steps - 1).all()
    assert np.array_equal(score, o_score), \
        np.flatnonzero(score != o_score)[:10]
    for e in range(batch):
        st = venv.get_state(e)
        nb = int(st['n_bodies'])
        assert int(st['overflow']) == 0
        assert np.array_equal(st['pos'][:nb], o_pos[e][:nb]), (env_id, e)
    assert venv.overflow_count() == 0
    # the sweep must exercise the score function, not just zeros
    if env_id.startswith('MoveToCorner') or n_scenes:
        assert len(np.unique(score)) > 1, np.unique(score)
    venv.close()


@pytest.mark.parametrize('env_id,n_scenes,n_steps', [
    ('MatchRegions-TestAll-LoResStack-v0', 128, 25),
    ('ClusterShape-TestAll-LoResStack-v0', 128, 60),
    ('FindDupe-TestAll-LoResStack-v0', 128, 40),
    ('MakeLine-TestCountPlus-LoResStack-v0', 128, 50),
    ('FixColour-TestAll-LoResStack-v0', 128, 15),
    ('MoveToRegion-TestAll-LoResStack-v0', 128, 33),
])
def test_render_sweep_random_layouts(built, env_id, n_scenes, n_steps):
    """Pixels over many layouts: 512 environments on 128 sampled scenes
    (random shapes, colours, counts, poses), stepped with random actions;
    the newest allo and ego frame of EVERY environment equals the oracle's
    brute-force rasteriser + 4x4 mean, bit for bit."""
    import multiprocessing as mp
    import os
    import torch
    import magical_b200 as magical
    from oracle_lib import rollout_frames
    batch = 512
    venv = magical.make_vec(env_id, batch, auto_reset=False,
                            n_scenes=n_scenes, seed=17)
    scene_ids = np.arange(batch) % n_scenes
    venv.reset(scene_ids=scene_ids)
    rng = np.random.RandomState(23)
    acts = _actions(rng, n_steps, batch, batch)
    chunk = 16
    jobs = [([venv.scenes[scene_ids[e]] for e in range(lo, lo + chunk)],
             acts[:, lo:lo + chunk]) for lo in range(0, batch, chunk)]
    with _oracle_pool() as pool:
        want = np.concatenate(pool.map(rollout_frames, jobs))
    for t in range(n_steps):
        obs, _, _, _ = venv.step(torch.from_numpy(acts[t]).cuda())
    got = obs[:, :, :, :, 9:12].cpu().numpy()        # [view, env, 96, 96, 3]
    for v, name in enumerate(('allo', 'ego')):
        bad = np.flatnonzero((got[v] != want[:, v]).reshape(batch, -1).any(1))
        assert len(bad) == 0, (env_id, name, bad[:10],
                               int((got[v][bad[0]] != want[bad[0], v]).sum()))
    venv.close()
