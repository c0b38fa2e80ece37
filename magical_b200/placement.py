"""Reset-time pose randomisation for the Test* variants (host side).

Restates the reference's rejection sampler -- `pm_randomise_pose`
(magical/geom.py:116-264), `pm_randomise_all_poses` (geom.py:281-341) and
`pm_shift_bodies` (geom.py:362-384) -- on top of a `scene.SceneBuilder`
instead of a live `pymunk.Space`.  The draw order of random numbers per try
(x, y, angle from `rng.uniform`) and the try / retry limits are the
reference's; the overlap predicate that pymunk answers with
`space.shape_query` (Chipmunk `cpSpaceShapeQuery`: filter reject by group,
then an exact narrowphase, sensors included in the result list) is restated
here as an exact convex-distance test `dist(core_a, core_b) <= r_a + r_b`
over the three shape kinds the reference creates (circle, fat segment,
convex polygon with bevel radius).

This is scene *construction* (it runs once per pre-sampled scene, like
`Entity.setup`), not the per-step hot path, so it is plain Python.
"""
import math

from magical_b200 import entities as en
from magical_b200 import scene as sc

MAX_TRIES = 10000


class PlacementError(Exception):
    """No non-colliding pose found (reference geom.py:111-113)."""


# ---------------------------------------------------------------------------
# exact overlap predicate for convex "cores" inflated by a radius
# ---------------------------------------------------------------------------

def _seg_seg_dist2(p1, q1, p2, q2):
    """Squared distance between segments p1q1 and p2q2 (either may be a
    point)."""
    d1x, d1y = q1[0] - p1[0], q1[1] - p1[1]
    d2x, d2y = q2[0] - p2[0], q2[1] - p2[1]
    rx, ry = p1[0] - p2[0], p1[1] - p2[1]
    a = d1x * d1x + d1y * d1y
    e = d2x * d2x + d2y * d2y
    f = d2x * rx + d2y * ry
    if a <= 0.0 and e <= 0.0:
        return rx * rx + ry * ry
    if a <= 0.0:
        s = 0.0
        t = min(1.0, max(0.0, f / e))
    else:
        c = d1x * rx + d1y * ry
        if e <= 0.0:
            t = 0.0
            s = min(1.0, max(0.0, -c / a))
        else:
            b = d1x * d2x + d1y * d2y
            denom = a * e - b * b
            s = min(1.0, max(0.0, (b * f - c * e) / denom)) \
                if denom > 0.0 else 0.0
            t = (b * s + f) / e
            if t < 0.0:
                t = 0.0
                s = min(1.0, max(0.0, -c / a))
            elif t > 1.0:
                t = 1.0
                s = min(1.0, max(0.0, (b - c) / a))
    cx = rx + d1x * s - d2x * t
    cy = ry + d1y * s - d2y * t
    return cx * cx + cy * cy


def _orient(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def _segments_cross(p1, q1, p2, q2):
    """Proper crossing (interiors intersect), decided by orientation signs so
    that rounding in the distance formula cannot turn a crossing into a tiny
    positive gap."""
    o1, o2 = _orient(p1, q1, p2), _orient(p1, q1, q2)
    o3, o4 = _orient(p2, q2, p1), _orient(p2, q2, q1)
    return ((o1 > 0.0) != (o2 > 0.0) and (o3 > 0.0) != (o4 > 0.0)
            and o1 != 0.0 and o2 != 0.0 and o3 != 0.0 and o4 != 0.0)


def _edges(core):
    n = len(core)
    if n == 1:
        return [(core[0], core[0])]
    if n == 2:
        return [(core[0], core[1])]
    return [(core[i], core[(i + 1) % n]) for i in range(n)]


def _inside_convex(p, poly):
    """p inside or on a CCW convex polygon (n >= 3)."""
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        if (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) < 0.0:
            return False
    return True


def core_distance(core_a, core_b):
    """Distance between two convex cores (point, segment or CCW polygon);
    zero when they intersect."""
    if len(core_a) >= 3 and _inside_convex(core_b[0], core_a):
        return 0.0
    if len(core_b) >= 3 and _inside_convex(core_a[0], core_b):
        return 0.0
    best = math.inf
    for p1, q1 in _edges(core_a):
        for p2, q2 in _edges(core_b):
            if _segments_cross(p1, q1, p2, q2):
                return 0.0
            d2 = _seg_seg_dist2(p1, q1, p2, q2)
            if d2 < best:
                best = d2
                if best == 0.0:
                    return 0.0
    return math.sqrt(best)


class WorldShape:
    """One collision shape in world coordinates."""
    __slots__ = ('core', 'radius', 'group', 'bb')

    def __init__(self, core, radius, group):
        self.core = core
        self.radius = radius
        self.group = group
        xs = [p[0] for p in core]
        ys = [p[1] for p in core]
        self.bb = (min(xs) - radius, min(ys) - radius, max(xs) + radius,
                   max(ys) + radius)


def shapes_collide(a, b):
    """ShapeQuery of Chipmunk's cpSpaceShapeQuery: group filter, bounding
    boxes, then the exact test."""
    if a.group != 0 and a.group == b.group:
        return False
    if a.bb[0] > b.bb[2] or b.bb[0] > a.bb[2] or a.bb[1] > b.bb[3] \
            or b.bb[1] > a.bb[3]:
        return False
    # the slack only ever rejects a placement that is within rounding of
    # touching, which the sampler then simply redraws
    return core_distance(a.core, b.core) <= a.radius + b.radius + 1e-12


# ---------------------------------------------------------------------------
# entity <-> builder plumbing
# ---------------------------------------------------------------------------

def _rot(v, c, s):
    return (v[0] * c - v[1] * s, v[0] * s + v[1] * c)


def entity_poses(builder, ent):
    """[(pos, angle)] of the entity's bodies, main body first (the order of
    `Entity.bodies` in the reference)."""
    if isinstance(ent, en.GoalRegion):
        g = builder.goals[ent.goal_index]
        return [((g['cx'], g['cy']), 0.0)]
    return [(tuple(builder.bodies[i]['p0']), builder.bodies[i]['a0'])
            for i in ent.body_ids]


def entity_pose(builder, ent):
    return entity_poses(builder, ent)[0]


def _set_poses(builder, ent, poses):
    if isinstance(ent, en.GoalRegion):
        (cx, cy), _ = poses[0]
        ent.move_to(builder, cx, cy)
        return
    for i, (pos, angle) in zip(ent.body_ids, poses):
        builder.bodies[i]['p0'] = (float(pos[0]), float(pos[1]))
        builder.bodies[i]['a0'] = float(angle)


def shift_entity(builder, ent, position=None, angle=None):
    """Rigid transform of all the entity's bodies so that the main body lands
    on (position, angle) -- pm_shift_bodies (geom.py:362-384)."""
    poses = entity_poses(builder, ent)
    root_pos, root_angle = poses[0]
    if angle is None:
        angle = root_angle
    if position is None:
        position = root_pos
    position = (float(position[0]), float(position[1]))
    angle = float(angle)
    dc, ds = math.cos(angle - root_angle), math.sin(angle - root_angle)
    new = []
    for pos, ang in poses:
        delta = _rot((pos[0] - root_pos[0], pos[1] - root_pos[1]), dc, ds)
        new.append(((position[0] + delta[0], position[1] + delta[1]),
                    angle + (ang - root_angle)))
    _set_poses(builder, ent, new)


def world_shapes(builder, ent):
    """The entity's collision shapes at the builder's current poses."""
    if isinstance(ent, en.GoalRegion):
        g = builder.goals[ent.goal_index]
        hw, hh = g['w'] / 2, g['h'] / 2
        cx, cy = g['cx'], g['cy']
        core = [(cx - hw, cy - hh), (cx + hw, cy - hh), (cx + hw, cy + hh),
                (cx - hw, cy + hh)]
        return [WorldShape(core, 0.0, 0)]
    out = []
    for gi in ent.cgroup_ids:
        grp = builder.cgroups[gi]
        for si in range(grp['shape0'], grp['shape0'] + grp['nshape']):
            sh = builder.shapes[si]
            verts = builder.cverts[sh['vert0']:sh['vert0'] + sh['nvert']]
            if sh['body'] >= 0:
                body = builder.bodies[sh['body']]
                c, s = math.cos(body['a0']), math.sin(body['a0'])
                px, py = body['p0']
                verts = [(px + v[0] * c - v[1] * s, py + v[0] * s + v[1] * c)
                         for v in verts]
            out.append(WorldShape(list(verts), sh['radius'], sh['group']))
    return out


# ---------------------------------------------------------------------------
# the samplers
# ---------------------------------------------------------------------------

def _listify(value, n):
    if isinstance(value, (list, tuple)):
        assert len(value) == n, (len(value), n)
        return list(value)
    return [value] * n


def randomise_pose(builder, ent, arena_lrbt, rng, rand_pos=True,
                   rand_rot=True, rel_pos_linf_limit=None, rel_rot_limit=None,
                   ignore_ents=None, disabled=()):
    """Rejection-sample a pose for `ent` that touches nothing else in the
    scene (pm_randomise_pose, geom.py:116-264).  `disabled` are entities whose
    collisions are switched off (not placed yet); `ignore_ents` are entities
    whose shapes are removed from the query result.  Returns the number of
    rejected tries."""
    assert rand_pos or rand_rot, \
        "need to randomise at least one thing, or placement may be impossible"
    saved = entity_poses(builder, ent)
    orig_pos, orig_angle = saved[0]
    skip = {id(ent)}
    skip.update(id(e) for e in (ignore_ents or ()))
    skip.update(id(e) for e in disabled)
    others = []
    for other in builder.entities:
        if id(other) not in skip:
            others.extend(world_shapes(builder, other))

    arena_l, arena_r, arena_b, arena_t = arena_lrbt
    if rel_pos_linf_limit is not None:
        assert 0 <= rel_pos_linf_limit
        pos_x = (max(arena_l, orig_pos[0] - rel_pos_linf_limit),
                 min(arena_r, orig_pos[0] + rel_pos_linf_limit))
        pos_y = (max(arena_b, orig_pos[1] - rel_pos_linf_limit),
                 min(arena_t, orig_pos[1] + rel_pos_linf_limit))
    else:
        pos_x = (arena_l, arena_r)
        pos_y = (arena_b, arena_t)
    if rel_rot_limit is not None:
        assert 0 <= rel_rot_limit
        rot_min = orig_angle - rel_rot_limit
        rot_max = orig_angle + rel_rot_limit
    else:
        rot_min, rot_max = -math.pi, math.pi

    n_tries = 0
    while n_tries < MAX_TRIES:
        if rand_pos:
            new_pos = (rng.uniform(*pos_x), rng.uniform(*pos_y))
        else:
            new_pos = orig_pos
        new_angle = rng.uniform(rot_min, rot_max) if rand_rot else orig_angle
        shift_entity(builder, ent, position=new_pos, angle=new_angle)
        mine = world_shapes(builder, ent)
        if not any(shapes_collide(a, b) for a in mine for b in others):
            return n_tries
        n_tries += 1
    _set_poses(builder, ent, saved)
    raise PlacementError(
        f"could not place {type(ent).__name__} after {n_tries} attempts "
        f"(rand_pos={rand_pos}, rand_rot={rand_rot}, arena={arena_lrbt})")


def randomise_all_poses(builder, entities, arena_lrbt, rng, rand_pos=True,
                        rand_rot=True, rel_pos_linf_limits=None,
                        rel_rot_limits=None, ignore_ents=None,
                        max_retries=10):
    """Place the entities one after the other, each avoiding everything
    placed before it and everything outside the list
    (pm_randomise_all_poses, geom.py:281-341)."""
    entities = list(entities)
    n = len(entities)
    pos_limits = _listify(rel_pos_linf_limits, n)
    rot_limits = _listify(rel_rot_limits, n)
    rand_pos = _listify(rand_pos, n)
    rand_rot = _listify(rand_rot, n)
    for retry in range(max_retries):
        try:
            for k, ent in enumerate(entities):
                randomise_pose(builder, ent, arena_lrbt, rng,
                               rand_pos=rand_pos[k], rand_rot=rand_rot[k],
                               rel_pos_linf_limit=pos_limits[k],
                               rel_rot_limit=rot_limits[k],
                               ignore_ents=ignore_ents,
                               disabled=entities[k + 1:])
            return
        except PlacementError:
            if retry == max_retries - 1:
                raise
