"""Reset-time pose randomisation of the Test* variants (magical_b200.placement
restating reference geom.py:116-384): non-overlap, arena containment, jitter
bounds, determinism under seeding."""
import math

import numpy as np
import pytest

import magical_b200 as magical
from magical_b200 import benchmarks, placement
from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.phys_vars import PhysicsVariables
from oracle_lib import OracleEnv

magical.register_envs()
TEST_IDS = sorted(i for i in benchmarks.ENV_SPECS
                  if benchmarks.EnvName(i).is_test and 'LoRes' not in i)


def _solid_pairs(rec):
    groups = rec['cgroups'][:rec['n_cgroups']]
    for ga, gb in rec['bpairs'][:rec['n_bpairs']]:
        for i in range(groups[ga]['nshape']):
            for k in range(groups[gb]['nshape']):
                yield int(groups[ga]['shape0']) + i, int(groups[gb]['shape0']) + k


# variants that re-place every entity (the others keep the Demo layout, where
# e.g. a randomised shape type may legitimately start in contact)
LAYOUT_IDS = [i for i in TEST_IDS
              if benchmarks.EnvName(i).variant in (
                  'TestJitter', 'TestLayout', 'TestCountPlus', 'TestAll')]


@pytest.mark.parametrize('env_id', LAYOUT_IDS)
def test_randomised_scene_is_collision_free_and_inside_arena(env_id):
    task, _ = magical.make_task(env_id)
    task.seed(11)
    for _ in range(2):
        rec = task.build_scene()
        orc = OracleEnv(rec)
        for sa, sb in _solid_pairs(rec):
            *_, cnt, _p1, _p2, _h = orc.collide(sa, sb)
            assert cnt == 0, (env_id, sa, sb)
        nb = int(rec['n_bodies'])
        pos = rec['bodies']['p0'][:nb]
        dyn = rec['bodies']['m_inv'][:nb] > 0
        # eye bodies ride along at the robot's shifted origin offset; every
        # body that owns a shape must be inside the arena
        owners = set(int(b) for b in rec['shapes']['body'][:rec['n_shapes']]
                     if b >= 0)
        for b in owners:
            assert dyn[b]
            assert np.all(np.abs(pos[b]) <= 1.0), (env_id, b, pos[b])
        orc.close()


def test_seeding_is_deterministic_and_seeds_differ():
    a, _ = magical.make_task('ClusterColour-TestAll-v0')
    b, _ = magical.make_task('ClusterColour-TestAll-v0')
    a.seed(3)
    b.seed(3)
    ra, rb = a.build_scene(), b.build_scene()
    assert ra.tobytes() == rb.tobytes()
    b.seed(4)
    assert b.build_scene().tobytes() != ra.tobytes()


@pytest.mark.parametrize('env_id', [
    'MoveToCorner-TestJitter-v0', 'MatchRegions-TestJitter-v0',
    'ClusterShape-TestJitter-v0', 'MakeLine-TestJitter-v0'])
def test_jitter_stays_within_reference_bounds(env_id):
    demo_id = env_id.replace('TestJitter', 'Demo')
    demo, _ = magical.make_task(demo_id)
    drec = demo.build_scene()
    task, _ = magical.make_task(env_id)
    task.seed(5)
    rec = task.build_scene()
    nb = int(rec['n_bodies'])
    assert nb == int(drec['n_bodies'])
    pos_bound = task.JITTER_POS_BOUND + 1e-12
    rot_bound = task.JITTER_ROT_BOUND + 1e-12
    main = [int(rec['robot_body'])] + [int(b) for b in
                                       rec['blocks']['body'][:rec['n_blocks']]]
    moved = False
    for b in main:
        d = np.abs(rec['bodies']['p0'][b] - drec['bodies']['p0'][b])
        assert np.all(d <= pos_bound), (b, d)
        da = abs(rec['bodies']['a0'][b] - drec['bodies']['a0'][b])
        assert da <= rot_bound
        moved |= bool(d.max() > 0)
    assert moved


def test_shift_entity_is_rigid():
    task, _ = magical.make_task('MoveToCorner-TestJitter-v0')
    task.seed(1)
    task.build_scene()
    # rebuild by hand to keep the builder alive
    task._entities = []
    task._builder = sc.SceneBuilder(task.TASK_ID, 80,
                                    PhysicsVariables.defaults())
    task.add_entities([en.ArenaBoundaries(-1, 1, 1, -1)])
    robot = task._make_robot((0.1, -0.2), 0.3)
    task.add_entities([robot])
    before = placement.entity_poses(task._builder, robot)
    placement.shift_entity(task._builder, robot, position=(-0.4, 0.5),
                           angle=0.3 + math.pi / 2)
    after = placement.entity_poses(task._builder, robot)
    assert after[0][0] == pytest.approx((-0.4, 0.5))
    for (p0, a0), (p1, a1) in zip(before, after):
        assert a1 - a0 == pytest.approx(math.pi / 2)
        rel0 = np.subtract(p0, before[0][0])
        rel1 = np.subtract(p1, after[0][0])
        # rotated by +90 degrees
        assert rel1 == pytest.approx((-rel0[1], rel0[0]), abs=1e-12)


def test_core_distance_against_brute_force():
    rng = np.random.RandomState(0)
    for _ in range(200):
        na, nb = rng.randint(1, 7, size=2)

        def poly(n):
            c = rng.uniform(-1, 1, size=2)
            if n == 1:
                return [tuple(c)]
            if n == 2:
                return [tuple(c), tuple(c + rng.uniform(-0.5, 0.5, size=2))]
            ang = np.sort(rng.uniform(0, 2 * math.pi, size=n))
            r = rng.uniform(0.1, 0.5)
            return [(c[0] + r * math.cos(t), c[1] + r * math.sin(t))
                    for t in ang]
        A, B = poly(na), poly(nb)
        d = placement.core_distance(A, B)

        def sample(P, m=60):
            if len(P) == 1:
                return np.asarray(P)
            pts = []
            E = placement._edges(P)
            for p, q in E:
                t = np.linspace(0, 1, m)[:, None]
                pts.append(np.asarray(p) * (1 - t) + np.asarray(q) * t)
            return np.concatenate(pts)
        SA, SB = sample(A), sample(B)
        brute = np.sqrt(((SA[:, None] - SB[None]) ** 2).sum(-1)).min()
        if d > 0:
            assert d <= brute + 1e-9
            assert d >= brute - 0.03
