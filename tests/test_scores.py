"""Oracle score functions against straight Python restatements of the
reference's score_on_end_of_traj (benchmarks/*.py) on hand-built layouts."""
import numpy as np
import pytest

from conftest import make_demo_task
from magical_b200.benchmarks.make_line import longest_line
from oracle_lib import OracleEnv


def _place(orc, rec, block, x, y, a=0.0):
    orc.set_pose(int(rec['blocks'][block]['body']), x, y, a)


def test_move_to_corner_score():
    rec = make_demo_task('MoveToCorner').build_scene()
    orc = OracleEnv(rec)
    for (x, y), want in [((-1.0, 1.0), 1.0), ((-0.5, 0.5), 1.0), ((0.0, 0.0), 0.0), ((1.0, -1.0), 0.0),
                         ((-0.3, 0.3), None)]:
        _place(orc, rec, 0, x, y)
        d = np.hypot(-1 - x, 1 - y)
        ref = min(1.0, max(0.0, np.sqrt(2) - d) / (np.sqrt(2) - np.sqrt(2) / 2))
        assert orc.score() == pytest.approx(ref, abs=1e-12)
        if want is not None:
            assert orc.score() == pytest.approx(want, abs=1e-12)


def test_move_to_region_score_is_point_in_box():
    rec = make_demo_task('MoveToRegion').build_scene()
    orc = OracleEnv(rec)
    g = rec['goals'][0]
    rb = int(rec['robot_body'])
    for dx, dy, want in [(0, 0, 1.0), (g['w'] / 2 - 1e-9, 0, 1.0), (g['w'] / 2 + 1e-6, 0, 0.0),
                         (0, -g['h'] / 2 - 1e-6, 0.0), (0, g['h'] / 2 - 1e-9, 1.0)]:
        orc.set_pose(rb, float(g['cx'] + dx), float(g['cy'] + dy), 0.3)
        assert orc.score() == want


def test_match_regions_score():
    rec = make_demo_task('MatchRegions').build_scene()
    orc = OracleEnv(rec)
    g = rec['goals'][0]
    roles = [int(rec['blocks'][i]['role']) for i in range(int(rec['n_blocks']))]
    targets = [i for i, r in enumerate(roles) if r == 1]
    distractors = [i for i, r in enumerate(roles) if r == 2]
    assert len(targets) == 2 and len(distractors) == 3
    cx, cy = float(g['cx']), float(g['cy'])
    # everything far away
    for i in range(5):
        _place(orc, rec, i, -0.8, -0.8 + 0.01 * i)
    assert orc.score() == 0.0
    _place(orc, rec, targets[0], cx - 0.1, cy)
    assert orc.score() == pytest.approx(0.5)
    _place(orc, rec, targets[1], cx + 0.1, cy)
    assert orc.score() == pytest.approx(1.0)
    _place(orc, rec, distractors[0], cx, cy + 0.15)
    assert orc.score() == pytest.approx(1.0 * (1 - 1 / 3))
    # centre of mass outside the box => not "in" even though the shape overlaps it
    _place(orc, rec, distractors[0], cx + float(g['w']) / 2 + 0.01, cy)
    assert orc.block_in_goal(distractors[0], 0) is False
    assert orc.score() == pytest.approx(1.0)


def test_make_line_score_matches_python_restatement():
    rec = make_demo_task('MakeLine').build_scene()
    orc = OracleEnv(rec)
    rng = np.random.RandomState(0)
    n = int(rec['n_blocks'])
    seen = set()
    for trial in range(300):
        if trial % 3 == 0:  # near-collinear layouts
            t = np.sort(rng.uniform(-0.8, 0.8, size=n))
            pts = np.stack([t, 0.3 * t + rng.normal(0, rng.choice([0.01, 0.08, 0.2]), size=n)], axis=1)
        else:
            pts = rng.uniform(-0.9, 0.9, size=(n, 2))
        for i in range(n):
            _place(orc, rec, i, pts[i, 0], pts[i, 1])
        line_len = longest_line(pts, 0.12 * 1.5, 0.12 * 3.5)
        min_len = max(n - 2, 2)
        want = max(line_len - min_len, 0) / (n - min_len)
        assert orc.score() == pytest.approx(want, abs=1e-12), (trial, pts)
        seen.add(want)
    assert seen == {0.0, 0.5, 1.0}


def test_cluster_score_matches_python_restatement():
    for task_name in ('ClusterColour', 'ClusterShape'):
        rec = make_demo_task(task_name).build_scene()
        orc = OracleEnv(rec)
        n, nvals = int(rec['n_blocks']), int(rec['n_labels'])
        labels = [int(rec['blocks'][i]['label']) for i in range(n)]
        rng = np.random.RandomState(1)
        seen = set()
        for trial in range(200):
            centres = rng.uniform(-0.7, 0.7, size=(nvals, 2))
            spread = rng.choice([0.01, 0.05, 0.3])
            pts = np.array([centres[l] + rng.normal(0, spread, size=2) for l in labels])
            for i in range(n):
                _place(orc, rec, i, pts[i, 0], pts[i, 1])
            cent = np.array([pts[[i for i in range(n) if labels[i] == c]].mean(axis=0) for c in range(nvals)])
            n_correct = 0
            for i in range(n):
                sses = ((pts[i] - cent) ** 2).sum(axis=1)
                true_sse = sses[labels[i]]
                bad = np.min(np.delete(sses, labels[i]))
                n_correct += int(np.sqrt(true_sse) < np.sqrt(bad) - 2.0 * true_sse)
            want = max(n_correct / n - 0.75, 0) / 0.25
            assert orc.score() == pytest.approx(want, abs=1e-9)
            seen.add(round(want, 3))
        assert 0.0 in seen and 1.0 in seen


def test_fix_colour_and_find_dupe_scores():
    rec = make_demo_task('FixColour').build_scene()
    orc = OracleEnv(rec)
    # Demo: blocks 0,1 match their regions, block 2 (blue) sits in the red region
    assert [int(g['expect_block']) for g in rec['goals'][:3]] == [0, 1, -1]
    assert orc.score() == 0.0                      # initial state: the odd block is still inside
    _place(orc, rec, 2, 0.8, -0.8)
    assert orc.score() == 1.0                      # odd one removed, others untouched
    _place(orc, rec, 0, 0.8, 0.8)
    assert orc.score() == 0.0                      # a correct block was removed too

    rec = make_demo_task('FindDupe').build_scene()
    orc = OracleEnv(rec)
    roles = [int(rec['blocks'][i]['role']) for i in range(int(rec['n_blocks']))]
    assert roles.count(1) == 2 and roles.count(2) == 5
    g = rec['goals'][0]
    assert orc.score() == 0.0                      # only the query block is inside
    dupe = roles.index(1)
    _place(orc, rec, dupe, float(g['cx']) + 0.15, float(g['cy']) + 0.1)
    assert orc.score() == pytest.approx(1.0)
    distractor = roles.index(2)
    _place(orc, rec, distractor, float(g['cx']) - 0.15, float(g['cy']) + 0.1)
    assert orc.score() == pytest.approx(1 - 1 / 3)
