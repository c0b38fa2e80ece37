"""Scene tables against the constants derived from the reference source
(entities.py / geom.py / base_env.py formulas)."""
import math

import numpy as np
import pytest

from conftest import demo_tasks, make_demo_task
from magical_b200 import entities as en
from magical_b200 import geom as g
from magical_b200 import scene as sc


def test_action_table():
    # reference entities.py:162-190: id = 9*grip + 3*lr + ud
    assert len(en.ACTION_NUMS_FLAGS_NAMES) == 18
    names = [n for _, _, n in en.ACTION_NUMS_FLAGS_NAMES]
    assert names[:4] == ['Open', 'UpOpen', 'DownOpen', 'LeftOpen']
    assert names[9] == 'Close' and names[17] == 'DownRightClose'
    A = en.RobotAction
    assert en.ACTION_ID_TO_FLAGS[4] == (A.UP, A.LEFT, A.OPEN)
    assert en.ACTION_ID_TO_FLAGS[16] == (A.UP, A.RIGHT, A.CLOSE)
    for a, flags, _ in en.ACTION_NUMS_FLAGS_NAMES:
        assert en.FLAGS_TO_ACTION_ID[flags] == a


def test_appendix_b_constants():
    R, S = 0.2, 0.12
    fl = g.make_finger_vertices(1.1 * R, 0.7 * R, 0.25 * R, -1)
    fr = g.make_finger_vertices(1.1 * R, 0.7 * R, 0.25 * R, 1)
    assert g.moment_for_poly(1 / 8, fl[0] + fl[1]) == pytest.approx(0.0050065, abs=5e-8)
    assert g.moment_for_poly(1 / 8, fr[0] + fr[1]) == pytest.approx(0.0049830, abs=5e-8)
    assert np.allclose(fl[1], [(0.0748, 0.3302), (0.0286, 0.3493), (-0.025, 0.22), (0.0212, 0.2009)], atol=5e-5)
    side = math.sqrt(math.pi) * S
    assert side == pytest.approx(0.2126945, abs=5e-8)
    hw = side / 2
    assert 0.5 * g.moment_for_poly(1.0, [(hw, -hw), (hw, hw), (-hw, hw), (-hw, -hw)]) == pytest.approx(0.0037699, abs=5e-8)
    pside = g.regular_poly_circ_rad_to_side_length(5, S)
    assert pside == pytest.approx(0.162156, abs=5e-7)
    assert g.regular_poly_circumrad(5, pside) == pytest.approx(0.137938, abs=5e-7)
    assert g.moment_for_poly(0.5, g.compute_regular_poly_verts(5, pside)) == pytest.approx(0.0036611, abs=5e-8)
    assert 0.8 * g.regular_poly_circ_rad_to_side_length(3, S) == pytest.approx(0.258581, abs=5e-7)
    star = g.compute_star_verts(5, 1.3 * S, 0.65 * S)
    assert g.moment_for_poly(0.5, g.to_convex_hull(star, 1e-5)) == pytest.approx(0.0046827, abs=5e-8)
    assert g.moment_for_circle(0.5, 0, S) == pytest.approx(0.0036)
    assert g.moment_for_circle(1.0, 0, R) == pytest.approx(0.02)


def test_convex_hull_order_and_decomposition():
    # Chipmunk's QuickHull: CCW from the leftmost (lowest) vertex
    pent = g.compute_regular_poly_verts(5, 0.1)
    hull = g.convex_hull(pent)
    assert len(hull) == 5 and hull[0] == min(pent)
    assert g.area_for_poly(hull) > 0
    # exact convex partition of the star: parts are convex, CCW, areas add up
    star = g.compute_star_verts(5, 0.156, 0.078)
    parts = g.convex_decomposition(star + star[:1], 0)
    assert sum(g.area_for_poly(p) for p in parts) == pytest.approx(g.area_for_poly(star), rel=1e-12)
    for p in parts:
        n = len(p)
        assert n >= 3
        for i in range(n):
            a, b, c = p[i], p[(i + 1) % n], p[(i + 2) % n]
            assert g.vcross(g.vsub(b, a), g.vsub(c, b)) > -1e-12  # convex, CCW


@pytest.mark.parametrize('task_name,bodies,joints,blocks,goals', [
    ('MoveToRegion', 6, 10, 0, 1), ('MoveToCorner', 7, 12, 1, 0),
    ('MatchRegions', 11, 20, 5, 1), ('MakeLine', 10, 18, 4, 0),
    ('FindDupe', 13, 24, 7, 1), ('FixColour', 9, 16, 3, 3),
    ('ClusterColour', 14, 26, 8, 0), ('ClusterShape', 14, 26, 8, 0)])
def test_demo_scene_inventory(task_name, bodies, joints, blocks, goals):
    """Counts from SURVEY.md §8(a): 6 robot bodies + 1 per block; 10 robot
    joints + 2 per block; 4 wall segments + 5 robot shapes + block shapes."""
    rec = make_demo_task(task_name).build_scene()
    assert int(rec['n_bodies']) == bodies
    assert int(rec['n_joints']) == joints
    assert int(rec['n_blocks']) == blocks
    assert int(rec['n_goals']) == goals
    assert int(rec['max_steps']) == demo_tasks()[task_name][1]
    shapes = rec['shapes'][:int(rec['n_shapes'])]
    assert (shapes['kind'][:4] == sc.SHAPE_SEGMENT).all() and (shapes['body'][:4] == -1).all()
    # joints: robot chain kinds in insertion order (entities.py:255-354)
    kinds = list(rec['joints']['kind'][:joints])
    robot_chain = [sc.JOINT_PIVOT, sc.JOINT_GEAR, sc.JOINT_ROTARY_SPRING, sc.JOINT_ROTARY_SPRING,
                   sc.JOINT_PIN, sc.JOINT_ROTARY_LIMIT, sc.JOINT_MOTOR,
                   sc.JOINT_PIN, sc.JOINT_ROTARY_LIMIT, sc.JOINT_MOTOR]
    if task_name in ('MoveToCorner', 'MoveToRegion'):
        assert kinds[:10] == robot_chain        # robot added first
    else:
        assert kinds[-10:] == robot_chain       # robot added last (drawn on top)
    # default force limits (base_env.py:53-57)
    j = rec['joints'][:joints]
    assert set(np.round(j['max_force'][j['kind'] == sc.JOINT_MOTOR], 6)) == {4.0}
    piv = j[j['kind'] == sc.JOINT_PIVOT]['max_force']
    assert set(np.round(piv, 6)) <= {3.0, 1.5}


def test_scene_record_is_deterministic_for_demo():
    a = make_demo_task('ClusterColour').build_scene()
    b = make_demo_task('ClusterColour').build_scene()
    assert a.tobytes() == b.tobytes()


def test_rand_dynamics_samples_in_range():
    from magical_b200.benchmarks.move_to_corner import MoveToCornerEnv
    from conftest import DEMO_KW
    env = MoveToCornerEnv(rand_dynamics=True, max_episode_steps=80, **DEMO_KW)
    env.seed(3)
    forces = set()
    for _ in range(5):
        rec = env.build_scene()
        j = rec['joints'][:int(rec['n_joints'])]
        f = float(j[j['kind'] == sc.JOINT_MOTOR]['max_force'][0])
        assert 2.5 <= f <= 4.5
        forces.add(f)
    assert len(forces) == 5
