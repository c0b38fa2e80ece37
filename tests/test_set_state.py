"""mg_set_state / checkpoint-resume: a snapshot
taken with get_state carries everything Chipmunk carries between steps (poses,
velocities, bias velocities, joint accumulators, arbiter cache).  CPU half:
the product's EnvState <-> mg_state_t conversion (mg_state_io.h, shared with
mg_api.cu) on the host build of the kernel source, against the oracle.  GPU
half (through the C ABI): get -> perturb -> set -> step == oracle bit for bit."""
import numpy as np
import pytest

from conftest import make_demo_task
from oracle_lib import OracleEnv
from tpe_host_lib import TpeHostEnv
from test_tpe_host import _compare


def _pushy_actions(seed, n):
    rng = np.random.RandomState(seed)
    return [int(rng.randint(18)) if rng.rand() < 0.5
            else int(rng.choice([1, 4, 7, 10, 13, 16])) for _ in range(n)]


def _perturb(st):
    """Kick two blocks and nudge the robot: a state no rollout from reset passed through."""
    st = st.copy()
    nb = int(st['n_bodies'])
    st['vel'][nb - 1] += (0.31, -0.17)
    st['angvel'][nb - 2] += 0.9
    st['pos'][0] += (0.013, -0.021)
    st['angle'][0] += 0.11
    return st


@pytest.mark.parametrize('task_name', ['ClusterColour', 'MatchRegions'])
def test_host_snapshot_roundtrip_and_cross_restore(task_name):
    rec = make_demo_task(task_name).build_scene()
    acts = _pushy_actions(3, 150)
    env, orc = TpeHostEnv(rec), OracleEnv(rec, det_sincos=True)
    seen_cached = 0
    for t, a in enumerate(acts[:90]):
        env.step(a)
        orc.step(a)
    snap, osnap = env.state(), orc.state()
    # the two snapshots agree entry for entry (cache order: newest contacts first, canonical order)
    n = int(snap['n_cache'])
    assert n == int(osnap['n_cache']) and n >= 1
    for key in ('cache_shapes', 'cache_hash', 'cache_age', 'cache_jn', 'cache_jt'):
        assert np.array_equal(snap[key][:n], osnap[key][:n]), key
    for key in ('bias_vel', 'bias_angvel'):
        assert np.array_equal(snap[key], osnap[key]), key
    assert int(snap['stamp']) == int(osnap['stamp']) == 900
    seen_cached += int((snap['cache_age'][:n] > 0).sum())
    # restore the ORACLE's snapshot into a fresh kernel env and the KERNEL's into a fresh oracle
    env2, orc2 = TpeHostEnv(rec), OracleEnv(rec, det_sincos=True)
    env2.set_state(osnap)
    orc2.set_state(snap)
    for t, a in enumerate(acts[90:]):
        for e in (env, env2):
            e.step(a)
        for o in (orc, orc2):
            o.step(a)
        _compare(env2.state(), orc.state(), ('restored kernel', t))
        _compare(env.state(), orc2.state(), ('restored oracle', t))
    for e in (env, env2, orc, orc2):
        e.close()


def test_host_perturbed_state_continues_like_the_oracle():
    rec = make_demo_task('ClusterColour').build_scene()
    acts = _pushy_actions(5, 140)
    env, orc = TpeHostEnv(rec), OracleEnv(rec, det_sincos=True)
    for a in acts[:70]:
        env.step(a)
        orc.step(a)
    st = _perturb(env.state())
    env.set_state(st)
    orc.set_state(st)
    most = 0
    for t, a in enumerate(acts[70:]):
        env.step(a)
        orc.step(a)
        most = max(most, _compare(env.state(), orc.state(), t))
    assert most >= 1
    env.close()
    orc.close()


def test_host_rejects_a_snapshot_of_another_scene():
    a = TpeHostEnv(make_demo_task('ClusterColour').build_scene())
    b = TpeHostEnv(make_demo_task('MoveToCorner').build_scene())
    with pytest.raises(AssertionError):
        b.set_state(a.state())
    a.close()
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize('env_id', ['ClusterColour-Demo-LoRes4E-v0',
                                    'MatchRegions-Demo-LoRes4E-v0'])
def test_gpu_get_perturb_set_step_equals_oracle(env_id):
    import torch
    import magical_b200 as magical
    from magical_b200 import _native
    B = 6
    venv = magical.make_vec(env_id, batch=B, device=0, auto_reset=False)
    venv.reset()
    rec = venv.scenes[0]
    acts = _pushy_actions(11, 150)
    orc = OracleEnv(rec, det_sincos=True)
    for a in acts[:80]:
        venv.step(torch.full((B,), a, dtype=torch.int32, device='cuda'))
        orc.step(a)
    snap = venv.get_state(2)
    _compare(snap, orc.state(), 'before')
    assert int(snap['n_cache']) >= 1
    # (1) plain restore into another env of the batch: env 4 is first driven somewhere else
    venv.set_pose(4, 0, 0.3, 0.3, 1.0)
    venv.set_state(4, snap)
    # (2) perturbed restore into env 1 and into the oracle
    pert = _perturb(snap)
    venv.set_state(1, pert)
    orc_p = OracleEnv(rec, det_sincos=True)
    orc_p.set_state(pert)
    most = 0
    for t, a in enumerate(acts[80:]):
        venv.step(torch.full((B,), a, dtype=torch.int32, device='cuda'))
        orc.step(a)
        orc_p.step(a)
        _compare(venv.get_state(4), orc.state(), ('restored', t))
        _compare(venv.get_state(2), orc.state(), ('untouched', t))
        most = max(most, _compare(venv.get_state(1), orc_p.state(), ('perturbed', t)))
    assert most >= 1
    assert venv.overflow_count() == 0
    # a snapshot of a different scene is refused
    other = magical.make_vec('MoveToCorner-Demo-LoRes4E-v0', batch=1, device=0, auto_reset=False)
    other.reset()
    with pytest.raises(_native.NativeError, match='does not match'):
        venv.set_state(0, other.get_state(0))
    other.close()
    venv.close()
