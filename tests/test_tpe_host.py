"""The thread-per-environment physics step (mg_physics_tpe.h, the source the
sm_100a kernel k_physics_tpe executes) compiled for the host and compared with
the CPU oracle: every body's pose and velocity, joint and contact impulses
BIT-EXACT over random and contact-rich rollouts of all 8 Demo tasks and of
randomised Test scenes, with and without the contact spill area."""
import numpy as np
import pytest

import magical_b200 as magical
from conftest import demo_tasks, make_demo_task
from oracle_lib import OracleEnv
from tpe_host_lib import TpeHostEnv

TASKS = list(demo_tasks())


def _compare(st, ost, tag):
    nb, nj = int(st['n_bodies']), int(st['n_joints'])
    assert int(st['overflow']) == 0, tag
    for key in ('pos', 'angle', 'vel', 'angvel'):
        assert np.array_equal(st[key][:nb], ost[key][:nb]), (
            tag, key, np.abs(st[key][:nb] - ost[key][:nb]).max())
    assert np.array_equal(st['joint_acc'][:nj], ost['joint_acc'][:nj]), (
        tag, 'joint_acc',
        np.abs(st['joint_acc'][:nj] - ost['joint_acc'][:nj]).max())
    assert int(st['n_contacts']) == int(ost['n_contacts']), tag
    nc = int(st['n_contacts'])
    if nc:
        assert np.array_equal(st['contact_shapes'][:nc],
                              ost['contact_shapes'][:nc]), tag
        assert np.array_equal(st['contact_jn'][:nc], ost['contact_jn'][:nc])
        assert np.array_equal(st['contact_jt'][:nc], ost['contact_jt'][:nc])
    return nc


def _rollout(rec, actions, kcon=8, spill=True, nitems=32, max_surv=None,
             scratch_global=True):
    env = TpeHostEnv(rec, kcon=kcon, spill=spill, nitems=nitems,
                     max_surv=max_surv, scratch_global=scratch_global)
    orc = OracleEnv(rec, det_sincos=True)
    most = 0
    for t, a in enumerate(actions):
        env.step(int(a))
        orc.step(int(a))
        most = max(most, _compare(env.state(), orc.state(), t))
    env.close()
    orc.close()
    return most


@pytest.mark.parametrize('task_name', TASKS)
def test_random_rollout_bit_exact(task_name):
    rec = make_demo_task(task_name).build_scene()
    rng = np.random.RandomState(7)
    n = {'MoveToRegion': 200}.get(task_name, 130)
    _rollout(rec, rng.randint(0, 18, size=n))


@pytest.mark.parametrize('task_name', ['ClusterColour', 'MatchRegions',
                                       'FindDupe', 'MakeLine'])
def test_contact_rich_rollout_bit_exact(task_name):
    """Forward-biased actions push the robot through the block field."""
    rec = make_demo_task(task_name).build_scene()
    for seed in range(3):
        rng = np.random.RandomState(seed)
        acts = [int(rng.randint(18)) if rng.rand() < 0.5
                else int(rng.choice([1, 4, 7, 10, 13, 16]))
                for _ in range(240)]
        most = _rollout(rec, acts)
    assert most >= 2


def test_spill_area_and_tiny_private_capacity():
    """With room for only 2 contacts in the private words the rest lives in
    the spill area; results must not change."""
    rec = make_demo_task('ClusterColour').build_scene()
    rng = np.random.RandomState(1)
    acts = [int(rng.randint(18)) if rng.rand() < 0.5
            else int(rng.choice([1, 4, 7, 10, 13, 16])) for _ in range(240)]
    env = TpeHostEnv(rec, kcon=2, spill=True)
    assert env.kcon <= 3
    env.close()
    most = _rollout(rec, acts, kcon=2, spill=True)
    assert most > 3


@pytest.mark.parametrize('nitems', [1, 2, 3])
def test_item_word_overflow_takes_the_serial_tail(nitems):
    """More candidate shape pairs than item words: the remaining pairs are
    processed serially by the owner, same results."""
    for name in ('ClusterColour', 'MatchRegions'):
        rec = make_demo_task(name).build_scene()
        rng = np.random.RandomState(2)
        acts = [int(rng.randint(18)) if rng.rand() < 0.5
                else int(rng.choice([1, 4, 7, 10, 13, 16]))
                for _ in range(200)]
        most = _rollout(rec, acts, nitems=nitems)
        assert most >= 2


def test_items_and_sep_cache_in_private_words_layout():
    """The alternative layout (work items + separation cache in the private
    words instead of the per-environment scratch record) gives the same bits."""
    rec = make_demo_task('ClusterColour').build_scene()
    rng = np.random.RandomState(6)
    acts = [int(rng.randint(18)) if rng.rand() < 0.5
            else int(rng.choice([1, 4, 7, 10, 13, 16])) for _ in range(120)]
    _rollout(rec, acts, scratch_global=False)


def test_survivor_list_overflow_continues_serially():
    """More GJK survivors than the packed cooperative list holds: the owner
    finishes them itself, in canonical order."""
    rec = make_demo_task('ClusterColour').build_scene()
    rng = np.random.RandomState(4)
    acts = [int(rng.randint(18)) if rng.rand() < 0.5
            else int(rng.choice([1, 4, 7, 10, 13, 16])) for _ in range(240)]
    most = _rollout(rec, acts, max_surv=1)
    assert most >= 3


def test_capacity_overflow_is_flagged_without_spill():
    rec = make_demo_task('ClusterColour').build_scene()
    rng = np.random.RandomState(1)
    acts = [int(rng.randint(18)) if rng.rand() < 0.5
            else int(rng.choice([1, 4, 7, 10, 13, 16])) for _ in range(240)]
    env = TpeHostEnv(rec, kcon=2, spill=False)
    for a in acts:
        env.step(a)
    assert int(env.state()['overflow']) & 2
    env.close()


@pytest.mark.parametrize('env_id', ['MatchRegions-TestAll-v0',
                                    'ClusterShape-TestAll-v0',
                                    'FindDupe-TestAll-v0',
                                    'FixColour-TestAll-v0',
                                    'MoveToCorner-TestAll-v0'])
def test_randomised_scenes_bit_exact(env_id):
    task, _ = magical.make_task(env_id)
    task.seed(5)
    rng = np.random.RandomState(3)
    for _ in range(3):
        rec = task.build_scene()
        _rollout(rec, rng.randint(0, 18, size=60))
