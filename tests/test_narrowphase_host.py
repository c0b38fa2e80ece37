"""CPU tests of the product's narrowphase (the code the GPU lanes run,
compiled for the host) against (1) the oracle's independent recursive
restatement, bit for bit, and (2) brute-force geometry."""
import numpy as np
import pytest

import host_lib
from conftest import make_demo_task
from oracle_lib import OracleEnv


def _cross2(a, b):
    return a[0] * b[1] - a[1] * b[0]


def _random_poses(rng, scene, spread):
    nb = int(scene['n_bodies'])
    poses = np.zeros((nb, 3))
    poses[:, :2] = rng.uniform(-spread, spread, size=(nb, 2))
    poses[:, 2] = rng.uniform(-np.pi, np.pi, size=nb)
    return poses


@pytest.mark.parametrize('task_name', ['ClusterColour', 'FindDupe',
                                       'MatchRegions'])
def test_product_narrowphase_matches_oracle_bitwise(task_name):
    scene = make_demo_task(task_name).build_scene()
    orc = OracleEnv(scene, det_sincos=True)
    rng = np.random.RandomState(5)
    ns = int(scene['n_shapes'])
    n_hits = 0
    n_two = 0
    for trial in range(60):
        # small spread => many overlapping shapes (EPA path), incl. wall hits
        poses = _random_poses(rng, scene, spread=rng.choice([0.25, 0.6, 1.05]))
        for b in range(len(poses)):
            orc.set_pose(b, *poses[b])
        for sa in range(ns):
            for sb in range(sa + 1, ns):
                ba = int(scene['shapes'][sa]['body'])
                bb = int(scene['shapes'][sb]['body'])
                if ba == bb:
                    continue
                got = host_lib.collide(scene, poses, sa, sb)
                want = orc.collide(sa, sb)
                assert got[0] == want[0] and got[1] == want[1]
                assert got[3] == want[3], (trial, sa, sb, got[3], want[3])
                cnt = got[3]
                if cnt:
                    n_hits += 1
                    n_two += cnt == 2
                    assert np.array_equal(got[2], want[2])
                    assert np.array_equal(got[4][:cnt], want[4][:cnt])
                    assert np.array_equal(got[5][:cnt], want[5][:cnt])
                    assert np.array_equal(got[6][:cnt], want[6][:cnt])
    assert n_hits > 200 and n_two > 50, (n_hits, n_two)


def _world_poly(scene, poses, si):
    sh = scene['shapes'][si]
    v = scene['cverts'][int(sh['vert0']):int(sh['vert0']) + int(sh['nvert'])]
    b = int(sh['body'])
    if b < 0:
        return v.copy()
    x, y, a = poses[b]
    c, s = np.cos(a), np.sin(a)
    return np.stack([c * v[:, 0] - s * v[:, 1] + x,
                     s * v[:, 0] + c * v[:, 1] + y], axis=1)


def _seg_dist(p, a, b):
    d = b - a
    t = np.clip(np.dot(p - a, d) / np.dot(d, d), 0, 1)
    return np.linalg.norm(p - (a + t * d))


def _poly_distance_bruteforce(P, Q):
    """Separation distance of two convex polygons (0 if they overlap)."""
    def inside(p, poly):
        n = len(poly)
        return all(_cross2(poly[(i + 1) % n] - poly[i], p - poly[i]) >= 0
                   for i in range(n))
    if any(inside(p, Q) for p in P) or any(inside(q, P) for q in Q):
        return 0.0
    best = np.inf
    for A, B in ((P, Q), (Q, P)):
        for p in A:
            for i in range(len(B)):
                best = min(best, _seg_dist(p, B[i], B[(i + 1) % len(B)]))
    # edge crossings without contained vertices
    for i in range(len(P)):
        for j in range(len(Q)):
            a, b, c, d = P[i], P[(i + 1) % len(P)], Q[j], Q[(j + 1) % len(Q)]
            d1 = _cross2(b - a, c - a) * _cross2(b - a, d - a)
            d2 = _cross2(d - c, a - c) * _cross2(d - c, b - c)
            if d1 < 0 and d2 < 0:
                return 0.0
    return best


def _sat_penetration(P, Q):
    """Minimum translation distance of overlapping convex polygons (SAT)."""
    best = np.inf
    for A, B in ((P, Q), (Q, P)):
        for i in range(len(A)):
            e = A[(i + 1) % len(A)] - A[i]
            n = np.array([e[1], -e[0]]) / np.linalg.norm(e)
            depth = np.max(A @ n) - np.min(B @ n)
            best = min(best, depth)
    return best


def test_gjk_epa_distance_against_bruteforce():
    """GJK separation distance == brute-force vertex/edge distance, and EPA
    penetration depth == SAT minimum translation, for the polygon shapes of
    the ClusterShape scene (squares, pentagons, star parts, finger boxes)."""
    scene = make_demo_task('ClusterShape').build_scene()
    polys = [i for i in range(int(scene['n_shapes']))
             if int(scene['shapes'][i]['kind']) == 2]
    rng = np.random.RandomState(9)
    n_sep = n_pen = 0
    for trial in range(40):
        poses = _random_poses(rng, scene, spread=rng.choice([0.15, 0.5]))
        for _ in range(40):
            sa, sb = rng.choice(polys, size=2, replace=False)
            if scene['shapes'][sa]['body'] == scene['shapes'][sb]['body']:
                continue
            d, n, pa, pb = host_lib.gjk(scene, poses, int(sa), int(sb))
            P, Q = _world_poly(scene, poses, sa), _world_poly(scene, poses, sb)
            if d > 1e-9:
                want = _poly_distance_bruteforce(P, Q)
                assert abs(d - want) < 1e-9, (trial, sa, sb, d, want)
                assert abs(np.linalg.norm(pb - pa) - d) < 1e-9
                n_sep += 1
            elif d < -1e-9:
                want = _sat_penetration(P, Q)
                assert abs(-d - want) < 1e-9, (trial, sa, sb, d, want)
                n_pen += 1
            assert abs(np.linalg.norm(n) - 1) < 1e-12
    assert n_sep > 100 and n_pen > 100, (n_sep, n_pen)
