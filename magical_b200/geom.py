"""Geometry helpers for the MAGICAL world: vertex generators, convex hulls,
convex decomposition and mass moments.

Two groups of functions live here:

* restatements of the reference's own pure-math helpers in
  `magical/geom.py:13-63,101-108` and `magical/entities.py:193-214`
  (regular polygons, stars, rectangles, finger outlines), and
* restatements of the *published algorithms* of the third-party routines the
  reference calls through pymunk 5.6 / Chipmunk2D 7.0.x (`moment_for_poly`,
  `moment_for_circle`, the QuickHull used by `Poly(...)` and
  `autogeometry.to_convex_hull`, and `autogeometry.convex_decomposition`).
  That library is NOT vendored under /root/reference, so these follow the
  upstream algorithm descriptions and are "parity
  unpinned" with respect to real pymunk.

All functions use plain Python floats (IEEE double) and tuples so that the
operation order is explicit; results feed the compiled scene tables consumed by
both the CUDA library and the CPU oracle.
"""
import math

DBL_MIN = 2.2250738585072014e-308


# ---------------------------------------------------------------------------
# small vector helpers (tuples of floats)
# ---------------------------------------------------------------------------

def vrotated(v, angle):
    """Rotate `v` CCW by `angle` (same op order as pymunk's Vec2d.rotated)."""
    c = math.cos(angle)
    s = math.sin(angle)
    return (v[0] * c - v[1] * s, v[0] * s + v[1] * c)


def vadd(a, b):
    return (a[0] + b[0], a[1] + b[1])


def vsub(a, b):
    return (a[0] - b[0], a[1] - b[1])


def vcross(a, b):
    return a[0] * b[1] - a[1] * b[0]


def vdot(a, b):
    return a[0] * b[0] + a[1] * b[1]


def vlerp(a, b, t):
    return (a[0] * (1.0 - t) + b[0] * t, a[1] * (1.0 - t) + b[1] * t)


def vnormalize(v):
    inv = 1.0 / (math.sqrt(vdot(v, v)) + DBL_MIN)
    return (v[0] * inv, v[1] * inv)


# ---------------------------------------------------------------------------
# reference geom.py:13-63, 101-108
# ---------------------------------------------------------------------------

def regular_poly_circumrad(n_sides, side_length):
    return side_length / (2 * math.sin(math.pi / n_sides))


def regular_poly_circ_rad_to_side_length(n_sides, rad):
    """Side length of the regular n-gon with the area of a radius-`rad` disc
    (reference geom.py:18-22)."""
    p_n = math.pi / n_sides
    return 2 * rad * math.sqrt(p_n * math.tan(p_n))


def regular_poly_apothem_to_side_length(n_sides, apothem):
    return 2 * apothem * math.tan(math.pi / n_sides)


def regular_poly_side_length_to_apothem(n_sides, side_length):
    return side_length / (2 * math.tan(math.pi / n_sides))


def compute_regular_poly_verts(n_sides, side_length):
    """CCW vertices, first one straight up (reference geom.py:35-46)."""
    assert n_sides >= 3
    step = 2 * math.pi / n_sides
    top = (0.0, regular_poly_circumrad(n_sides, side_length))
    return [vrotated(top, k * step) for k in range(n_sides)]


def compute_star_verts(n_points, out_radius, in_radius):
    """Alternating tip / notch vertices, CCW (reference geom.py:49-63)."""
    assert n_points >= 3
    tip = (0.0, out_radius)
    notch = (0.0, in_radius)
    verts = []
    for k in range(n_points):
        verts.append(vrotated(tip, k * 2 * math.pi / n_points))
        verts.append(vrotated(notch, (2 * k + 1) * math.pi / n_points))
    return verts


def rect_verts(w, h):
    """CCW from top right (reference geom.py:101-108)."""
    return [(w / 2, h / 2), (-w / 2, h / 2), (-w / 2, -h / 2), (w / 2, -h / 2)]


def make_finger_vertices(upper_arm_len, forearm_len, thickness, side_sign):
    """Two rectangles (upper arm, forearm) of one gripper finger, in the
    finger body's frame: origin at the root of the upper arm, upper arm
    pointing along +y, forearm bent inwards by pi/8
    (reference entities.py:193-214)."""
    up_shift = upper_arm_len / 2
    upper = rect_verts(thickness, upper_arm_len)
    fore = rect_verts(thickness, forearm_len)
    upper_start = (side_sign * thickness / 2, upper_arm_len / 2)
    fore_off = (-side_sign * thickness / 2, forearm_len / 2)
    rot = side_sign * math.pi / 8
    trans = vadd(upper_start, vrotated(fore_off, rot))
    trans = (trans[0], trans[1] + up_shift)
    fore_final = [vadd(vrotated(v, rot), trans) for v in fore]
    upper_final = [(v[0], v[1] + up_shift) for v in upper]
    return upper_final, fore_final


# ---------------------------------------------------------------------------
# Chipmunk2D mass helpers (published formulas)
# ---------------------------------------------------------------------------

def moment_for_circle(mass, inner_radius, outer_radius, offset=(0.0, 0.0)):
    return mass * (0.5 * (inner_radius * inner_radius
                          + outer_radius * outer_radius) + vdot(offset, offset))


def moment_for_poly(mass, verts, offset=(0.0, 0.0)):
    """sum(a_i*b_i)*m / (6*sum(a_i)) over consecutive vertex pairs; the vertex
    list is used literally (no hull), as Chipmunk's cpMomentForPoly does."""
    n = len(verts)
    sum1 = 0.0
    sum2 = 0.0
    for i in range(n):
        v1 = vadd(verts[i], offset)
        v2 = vadd(verts[(i + 1) % n], offset)
        a = vcross(v2, v1)
        b = vdot(v1, v1) + vdot(v1, v2) + vdot(v2, v2)
        sum1 += a * b
        sum2 += a
    return (mass * sum1) / (6.0 * sum2)


def area_for_poly(verts):
    n = len(verts)
    area = 0.0
    for i in range(n):
        area += vcross(verts[i], verts[(i + 1) % n])
    return area / 2.0


# ---------------------------------------------------------------------------
# Chipmunk2D QuickHull (vertex ORDER matters: support-point tie-breaks and
# contact hashes index into it).  Output is CCW starting from the vertex with
# minimum x (ties: minimum y).
# ---------------------------------------------------------------------------

def _loop_indexes(verts):
    start = end = 0
    vmin = vmax = verts[0]
    for i in range(1, len(verts)):
        v = verts[i]
        if v[0] < vmin[0] or (v[0] == vmin[0] and v[1] < vmin[1]):
            vmin = v
            start = i
        elif v[0] > vmax[0] or (v[0] == vmax[0] and v[1] > vmax[1]):
            vmax = v
            end = i
    return start, end


def _qhull_partition(verts, lo, count, a, b, tol):
    """In-place partition of verts[lo:lo+count]: points strictly to the right
    of a->b first, with the farthest one moved to the front."""
    if count == 0:
        return 0
    best = 0.0
    pivot = 0
    delta = vsub(b, a)
    value_tol = tol * math.sqrt(vdot(delta, delta))
    head = 0
    tail = count - 1
    while head <= tail:
        value = vcross(vsub(verts[lo + head], a), delta)
        if value > value_tol:
            if value > best:
                best = value
                pivot = head
            head += 1
        else:
            verts[lo + head], verts[lo + tail] = verts[lo + tail], verts[lo + head]
            tail -= 1
    if pivot != 0:
        verts[lo], verts[lo + pivot] = verts[lo + pivot], verts[lo]
    return head


def _qhull_reduce(tol, verts, lo, count, a, pivot, b, out):
    if count < 0:
        return
    if count == 0:
        out.append(pivot)
        return
    left = _qhull_partition(verts, lo, count, a, pivot, tol)
    _qhull_reduce(tol, verts, lo + 1, left - 1, a, verts[lo], pivot, out)
    out.append(pivot)
    right = _qhull_partition(verts, lo + left, count - left, pivot, b, tol)
    _qhull_reduce(tol, verts, lo + left + 1, right - 1, pivot,
                  verts[lo + left] if right > 0 else None, b, out)


def convex_hull(verts, tol=0.0, return_first=False):
    """QuickHull with Chipmunk's traversal: returns the hull CCW from the
    leftmost-lowest input vertex; collinear points are dropped when tol=0."""
    work = [tuple(map(float, v)) for v in verts]
    count = len(work)
    start, end = _loop_indexes(work)
    if start == end:
        return ([work[0]], 0) if return_first else [work[0]]
    work[0], work[start] = work[start], work[0]
    second = start if end == 0 else end
    work[1], work[second] = work[second], work[1]
    a = work[0]
    b = work[1]
    out = [a]
    _qhull_reduce(tol, work, 2, count - 2, a, b, a, out)
    return (out, start) if return_first else out


# ---------------------------------------------------------------------------
# Chipmunk2D convex decomposition (deepest notch + Steiner point), used by the
# reference for star blocks (entities.py:653-654, 724-728).
# ---------------------------------------------------------------------------

def _deepest_notch(verts, hull, first):
    count = len(verts)
    hcount = len(hull)
    notch = dict(i=0, d=0.0, v=(0.0, 0.0), n=(0.0, 0.0))
    j = (first + 1) % count
    for i in range(hcount):
        a = hull[i]
        b = hull[(i + 1) % hcount]
        # inward normal of hull edge a->b
        n = vnormalize((vsub(a, b)[1], -vsub(a, b)[0]))
        d = vdot(n, a)
        v = verts[j]
        while v != b:
            depth = vdot(n, v) - d
            if depth > notch['d']:
                notch = dict(i=j, d=depth, v=v, n=n)
            j = (j + 1) % count
            v = verts[j]
        j = (j + 1) % count
    return notch


def _find_steiner(verts, notch):
    count = len(verts)
    best = math.inf
    feature = -1.0
    for i in range(1, count - 1):
        index = (notch['i'] + i) % count
        seg_a = verts[index]
        seg_b = verts[(index + 1) % count]
        thing_a = vcross(notch['n'], vsub(seg_a, notch['v']))
        thing_b = vcross(notch['n'], vsub(seg_b, notch['v']))
        if thing_a * thing_b <= 0.0:
            t = thing_a / (thing_a - thing_b)
            dist = vdot(notch['n'], vsub(vlerp(seg_a, seg_b, t), notch['v']))
            if 0.0 <= dist <= best:
                best = dist
                feature = index + t
    return feature


def _decompose(verts, tol, out):
    count = len(verts)
    hull, first = convex_hull(verts, 0.0, return_first=True)
    if len(hull) != count:
        notch = _deepest_notch(verts, hull, first)
        if notch['d'] > tol:
            steiner_it = _find_steiner(verts, notch)
            if steiner_it >= 0.0:
                steiner_i = int(steiner_it)
                steiner_t = math.fmod(steiner_it, 1.0)
                steiner = vlerp(verts[steiner_i % count],
                                verts[(steiner_i + 1) % count], steiner_t)
                sub1 = (steiner_i - notch['i'] + count) % count + 1
                sub2 = count - (steiner_i - notch['i'] + count) % count
                part1 = [verts[(notch['i'] + k) % count] for k in range(sub1)]
                part1.append(steiner)
                _decompose(part1, tol, out)
                part2 = [verts[(steiner_i + 1 + k) % count]
                         for k in range(sub2)]
                part2.append(steiner)
                _decompose(part2, tol, out)
                return
    out.append(hull)


def convex_decomposition(closed_polyline, tol=0.0, dedupe_eps=1e-12):
    """Exact convex partition of a CCW closed polyline (last vertex repeats the
    first).  Returns a list of convex CCW vertex lists (not closed).

    Deviation, documented in DESIGN.md: a star's notch ray passes (to within
    rounding) through the opposite tip, so the Steiner point can land a few
    ulps from an existing vertex; parts are re-hulled after merging vertices
    closer than `dedupe_eps`, and parts that degenerate to <3 vertices or to
    (numerically) zero area are dropped.  The union of the parts is unchanged.
    """
    assert closed_polyline[0] == closed_polyline[-1], "polyline must be closed"
    verts = [tuple(map(float, v)) for v in closed_polyline[:-1]]
    assert area_for_poly(verts) >= 0.0, "winding must be CCW"
    raw = []
    _decompose(verts, tol, raw)
    parts = []
    for part in raw:
        merged = []
        for v in part:
            if all(abs(v[0] - u[0]) + abs(v[1] - u[1]) > dedupe_eps
                   for u in merged):
                merged.append(v)
        if len(merged) < 3:
            continue
        hull = convex_hull(merged, 0.0)
        if len(hull) < 3 or abs(area_for_poly(hull)) < 1e-12:
            continue
        parts.append(hull)
    return parts


def to_convex_hull(verts, tol):
    """autogeometry.to_convex_hull: closed hull (first vertex repeated)."""
    hull = convex_hull(verts, tol)
    return hull + hull[:1]
