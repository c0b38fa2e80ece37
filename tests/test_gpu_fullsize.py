"""BASELINE.json's full batch sizes through size-independent properties (the
oracle cannot step 65 536 environments): environments that receive the same
actions stay bit-identical across the whole batch (no cross-environment
interference at full occupancy), two half-size handles equal one full-size
handle (the 1-GPU == N-GPU shard property, on one device), and a sampled
environment still equals the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
