/*
 * mg_narrowphase.h — per-lane convex narrowphase for the sm_100a physics
 * kernel (also compilable for the host so it can be unit-tested on CPU).
 *
 * Replaces what the reference reaches through pymunk: Chipmunk2D's cpCollide
 * (circle/segment/convex-poly pairs, GJK closest points + EPA penetration,
 * two-point edge clipping) as used by `pm.Space.step` (base_env.py:243) and
 * `space.shape_query` (entities.py:837).
 *
 * GPU-first structure (differs from the recursive CPU formulation in
 * oracle/mgo_physics.c, which it must match bit for bit):
 *   - no recursion: GJK and EPA are bounded loops;
 *   - no per-thread world-vertex arrays: a shape is a view (pointer to the
 *     scene's local vertices in global memory + the body's pose in
 *     registers) and world vertices are re-derived on demand, which is
 *     deterministic and keeps the lane free of local-memory arrays;
 *   - EPA's hull keeps only (ab, id) per Minkowski vertex; the two support
 *     points are re-derived from the id when the closest edge is found.
 * All arithmetic is IEEE double in the literal order written (the TU is
 * compiled with -fmad=false for the parity build).
 */
#ifndef MG_NARROWPHASE_H
#define MG_NARROWPHASE_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#if defined(MG_NP_FORCE_INLINE)
/* GJK / EPA inlined into their (single, itself out-of-line) caller: the shape views then live in
 * registers instead of being read through references from local memory at every vertex access */
#define MG_HD_NOINLINE __host__ __device__ __forceinline__
#else
#define MG_HD_NOINLINE static __host__ __device__ __noinline__
#endif
#else
#define MG_HD static inline
#define MG_HD_NOINLINE static
#endif

#define MG_DBL_MIN 2.2250738585072014e-308
#define MG_INF (__builtin_huge_val())
#define MG_MAX_GJK_ITERATIONS 30
#define MG_MAX_EPA_ITERATIONS 30
#define MG_EPA_HULL_CAP 34

struct d2 {
  double x, y;
};

MG_HD d2 D2(double x, double y) { d2 r; r.x = x; r.y = y; return r; }
MG_HD d2 dadd(d2 a, d2 b) { return D2(a.x + b.x, a.y + b.y); }
MG_HD d2 dsub(d2 a, d2 b) { return D2(a.x - b.x, a.y - b.y); }
MG_HD d2 dneg(d2 a) { return D2(-a.x, -a.y); }
MG_HD d2 dmul(d2 a, double s) { return D2(a.x * s, a.y * s); }
MG_HD double ddot(d2 a, d2 b) { return a.x * b.x + a.y * b.y; }
MG_HD double dcross(d2 a, d2 b) { return a.x * b.y - a.y * b.x; }
MG_HD d2 dperp(d2 a) { return D2(-a.y, a.x); }
MG_HD d2 drperp(d2 a) { return D2(a.y, -a.x); }
MG_HD d2 drotate(d2 a, d2 b) { return D2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
MG_HD double dlengthsq(d2 a) { return ddot(a, a); }
MG_HD double dlength(d2 a) { return sqrt(ddot(a, a)); }
MG_HD d2 dlerp(d2 a, d2 b, double t) { return dadd(dmul(a, 1.0 - t), dmul(b, t)); }
MG_HD d2 dnormalize(d2 a) { return dmul(a, 1.0 / (dlength(a) + MG_DBL_MIN)); }
MG_HD double dmaxf(double a, double b) { return (a > b) ? a : b; }
MG_HD double dminf(double a, double b) { return (a < b) ? a : b; }
MG_HD double dclamp(double f, double lo, double hi) { return dminf(dmaxf(f, lo), hi); }
MG_HD double dclamp01(double f) { return dmaxf(0.0, dminf(f, 1.0)); }
MG_HD d2 dvclamp(d2 v, double len) { return (ddot(v, v) > len * len) ? dmul(dnormalize(v), len) : v; }

/* A collision shape seen by one lane: local geometry in (global) scene memory,
 * pose of the owning body in registers. */
struct ShapeView {
  int kind;          /* MG_SHAPE_* (0 circle, 1 segment, 2 poly) */
  int nvert;
  const double* lv;  /* local vertices, interleaved x,y */
  const double* ln;  /* local plane normals, interleaved (segments: 1 normal) */
  double rc, rs;     /* body rotation (cos, sin) */
  double px, py;     /* body position */
  double radius;
  int index;         /* shape index in the scene (contact feature ids) */
};

MG_HD d2 sv_vert(const ShapeView& s, int i) {
  double x = s.lv[2 * i], y = s.lv[2 * i + 1];
  return D2(s.rc * x - s.rs * y + s.px, s.rs * x + s.rc * y + s.py);
}
MG_HD d2 sv_normal(const ShapeView& s, int i) {
  double x = s.ln[2 * i], y = s.ln[2 * i + 1];
  return D2(s.rc * x - s.rs * y, s.rs * x + s.rc * y);
}
/* l, b, r, t exactly as the shape's cached bounding box */
MG_HD void sv_bb(const ShapeView& s, double bb[4]) {
  if (s.kind == 0) {
    d2 c = sv_vert(s, 0);
    bb[0] = c.x - s.radius; bb[2] = c.x + s.radius; bb[1] = c.y - s.radius; bb[3] = c.y + s.radius;
  } else if (s.kind == 1) {
    d2 ta = sv_vert(s, 0), tb = sv_vert(s, 1);
    double l, r, b, t;
    if (ta.x < tb.x) { l = ta.x; r = tb.x; } else { l = tb.x; r = ta.x; }
    if (ta.y < tb.y) { b = ta.y; t = tb.y; } else { b = tb.y; t = ta.y; }
    bb[0] = l - s.radius; bb[1] = b - s.radius; bb[2] = r + s.radius; bb[3] = t + s.radius;
  } else {
    double l = MG_INF, r = -MG_INF, b = MG_INF, t = -MG_INF;
#pragma unroll 1
    for (int i = 0; i < s.nvert; i++) {
      d2 v = sv_vert(s, i);
      l = dminf(l, v.x); r = dmaxf(r, v.x); b = dminf(b, v.y); t = dmaxf(t, v.y);
    }
    bb[0] = l - s.radius; bb[1] = b - s.radius; bb[2] = r + s.radius; bb[3] = t + s.radius;
  }
}
MG_HD bool bb_intersects(const double* a, const double* b) {
  return a[0] <= b[2] && b[0] <= a[2] && a[1] <= b[3] && b[1] <= a[3];
}

MG_HD int sv_support_index(const ShapeView& s, d2 n) {
  if (s.kind == 0) return 0;
  if (s.kind == 1) return (ddot(sv_vert(s, 0), n) > ddot(sv_vert(s, 1), n)) ? 0 : 1;
  double best = -MG_INF;
  int index = 0;
#pragma unroll 1
  for (int i = 0; i < s.nvert; i++) {
    double d = ddot(sv_vert(s, i), n);
    if (d > best) { best = d; index = i; }
  }
  return index;
}

struct MinkPoint {
  d2 ab;
  unsigned id; /* (index on shape1) << 8 | index on shape2 */
};
MG_HD MinkPoint mk_support(const ShapeView& s1, const ShapeView& s2, d2 n) {
  int ia = sv_support_index(s1, dneg(n));
  int ib = sv_support_index(s2, n);
  MinkPoint m;
  m.ab = dsub(sv_vert(s2, ib), sv_vert(s1, ia));
  m.id = ((unsigned)ia & 0xFFu) << 8 | ((unsigned)ib & 0xFFu);
  return m;
}

struct ClosestPts {
  d2 a, b, n;
  double d;
};

MG_HD double closest_t(d2 a, d2 b) {
  d2 delta = dsub(b, a);
  return -dclamp(ddot(delta, dadd(a, b)) / dlengthsq(delta), -1.0, 1.0);
}
MG_HD d2 lerp_t(d2 a, d2 b, double t) {
  double ht = 0.5 * t;
  return dadd(dmul(a, 0.5 - ht), dmul(b, 0.5 + ht));
}
MG_HD double closest_dist(d2 v0, d2 v1) { return dlengthsq(lerp_t(v0, v1, closest_t(v0, v1))); }
MG_HD bool check_point_greater(d2 a, d2 b, d2 c) {
  return (b.y - a.y) * (a.x + b.x - 2 * c.x) > (b.x - a.x) * (a.y + b.y - 2 * c.y);
}
MG_HD bool check_axis(d2 v0, d2 v1, d2 p, d2 n) { return ddot(p, n) <= dmaxf(ddot(v0, n), ddot(v1, n)); }

MG_HD ClosestPts closest_points_new(const ShapeView& s1, const ShapeView& s2, MinkPoint v0, MinkPoint v1) {
  double t = closest_t(v0.ab, v1.ab);
  d2 p = lerp_t(v0.ab, v1.ab, t);
  d2 pa = lerp_t(sv_vert(s1, (int)(v0.id >> 8)), sv_vert(s1, (int)(v1.id >> 8)), t);
  d2 pb = lerp_t(sv_vert(s2, (int)(v0.id & 0xFFu)), sv_vert(s2, (int)(v1.id & 0xFFu)), t);
  d2 delta = dsub(v1.ab, v0.ab);
  d2 n = dnormalize(drperp(delta));
  double d = ddot(n, p);
  ClosestPts pts;
  pts.a = pa; pts.b = pb;
  if (d <= 0.0 || (-1.0 < t && t < 1.0)) {
    pts.n = n; pts.d = d;
  } else {
    double d2_ = dlength(p);
    pts.n = dmul(p, 1.0 / (d2_ + MG_DBL_MIN));
    pts.d = d2_;
  }
  return pts;
}

/* GJK closest points with the EPA penetration search folded into the same loop.  The loop is a small state
 * machine (0, 1: the two initial supports; 2: a GJK step; 3: an EPA step) so that the support mapping --
 * the bulk of the code and of the run time -- exists ONCE: lanes that are in different stages of different
 * shape pairs stay converged through it, and the instruction footprint stays small (the kernels that
 * inline this are instruction-fetch sensitive).  Every step performs exactly the arithmetic of Chipmunk's
 * recursive GJKRecurse / EPARecurse, in the same order. */
MG_HD_NOINLINE ClosestPts mg_gjk(const ShapeView& s1, const ShapeView& s2, const double* bb1, const double* bb2) {
  d2 c1 = dlerp(D2(bb1[0], bb1[1]), D2(bb1[2], bb1[3]), 0.5);
  d2 c2 = dlerp(D2(bb2[0], bb2[1]), D2(bb2[2], bb2[3]), 0.5);
  const d2 axis = dperp(dsub(c1, c2));
  MinkPoint v0, v1;             /* GJK simplex */
  MinkPoint bufA[MG_EPA_HULL_CAP], bufB[MG_EPA_HULL_CAP]; /* EPA hull, ping-pong */
  MinkPoint* hull = bufA;
  MinkPoint* hull2 = bufB;
  int count = 0, mini = 0;
  MinkPoint e0, e1;             /* closest EPA edge / the pair handed to closest_points_new */
  v0.ab = v1.ab = e0.ab = e1.ab = D2(0, 0);
  v0.id = v1.id = e0.id = e1.id = 0u;
  int phase = 0, iteration = 1, eiter = 1;
  d2 dir = axis;
  for (;;) {
    if (phase == 2) {
      if (iteration > MG_MAX_GJK_ITERATIONS) { e0 = v0; e1 = v1; break; }
      if (check_point_greater(v1.ab, v0.ab, D2(0, 0))) { MinkPoint t = v0; v0 = v1; v1 = t; }
      double t = closest_t(v0.ab, v1.ab);
      dir = (-1.0 < t && t < 1.0) ? dperp(dsub(v1.ab, v0.ab)) : dneg(lerp_t(v0.ab, v1.ab, t));
    } else if (phase == 3) {
      mini = 0;
      double min_dist = MG_INF;
#pragma unroll 1
      for (int j = 0, i = count - 1; j < count; i = j, j++) {
        double d = closest_dist(hull[i].ab, hull[j].ab);
        if (d < min_dist) { min_dist = d; mini = i; }
      }
      e0 = hull[mini];
      e1 = hull[(mini + 1) % count];
      dir = dperp(dsub(e1.ab, e0.ab));
    }
    const MinkPoint p = mk_support(s1, s2, dir); /* the one support-mapping site */
    if (phase == 0) { v0 = p; dir = dneg(axis); phase = 1; continue; }
    if (phase == 1) { v1 = p; phase = 2; continue; }
    if (phase == 2) {
      if (check_point_greater(p.ab, v0.ab, D2(0, 0)) && check_point_greater(v1.ab, p.ab, D2(0, 0))) {
        /* the origin is inside the simplex: the shapes overlap, continue with EPA on (v0, p, v1) */
        hull[0] = v0; hull[1] = p; hull[2] = v1;
        count = 3;
        phase = 3;
        continue;
      }
      if (check_axis(v0.ab, v1.ab, p.ab, dir)) { e0 = v0; e1 = v1; break; }
      if (closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)) v1 = p; else v0 = p;
      iteration++;
      continue;
    }
    /* phase 3: grow the hull by p, or stop at the closest edge */
    const bool duplicate = (p.id == e0.id || p.id == e1.id);
    if (!duplicate && check_point_greater(e0.ab, e1.ab, p.ab) && eiter < MG_MAX_EPA_ITERATIONS &&
        count + 1 < MG_EPA_HULL_CAP) {
      int count2 = 1;
      hull2[0] = p;
#pragma unroll 1
      for (int i = 0; i < count; i++) {
        int index = (mini + 1 + i) % count;
        d2 h0 = hull2[count2 - 1].ab;
        d2 h1 = hull[index].ab;
        d2 h2 = (i + 1 < count ? hull[(index + 1) % count] : p).ab;
        if (check_point_greater(h0, h2, h1)) { hull2[count2] = hull[index]; count2++; }
      }
      MinkPoint* tmp = hull; hull = hull2; hull2 = tmp;
      count = count2;
      eiter++;
    } else {
      break; /* (e0, e1) is the closest edge */
    }
  }
  return closest_points_new(s1, s2, e0, e1);
}

struct SupEdge {
  d2 a, b;
  unsigned ha, hb;
  double r;
};
MG_HD SupEdge support_edge_poly(const ShapeView& poly, d2 n) {
  int count = poly.nvert;
  int i1 = sv_support_index(poly, n);
  int i0 = (i1 - 1 + count) % count;
  int i2 = (i1 + 1) % count;
  unsigned base = (unsigned)poly.index << 8;
  SupEdge e;
  e.r = poly.radius;
  if (ddot(n, sv_normal(poly, i1)) > ddot(n, sv_normal(poly, i2))) {
    e.a = sv_vert(poly, i0); e.ha = base | (unsigned)i0;
    e.b = sv_vert(poly, i1); e.hb = base | (unsigned)i1;
  } else {
    e.a = sv_vert(poly, i1); e.ha = base | (unsigned)i1;
    e.b = sv_vert(poly, i2); e.hb = base | (unsigned)i2;
  }
  return e;
}
MG_HD SupEdge support_edge_segment(const ShapeView& seg, d2 n) {
  unsigned base = (unsigned)seg.index << 8;
  SupEdge e;
  e.r = seg.radius;
  if (ddot(sv_normal(seg, 0), n) > 0.0) {
    e.a = sv_vert(seg, 0); e.ha = base | 0u;
    e.b = sv_vert(seg, 1); e.hb = base | 1u;
  } else {
    e.a = sv_vert(seg, 1); e.ha = base | 1u;
    e.b = sv_vert(seg, 0); e.hb = base | 0u;
  }
  return e;
}

struct Manifold {
  d2 n;
  double margin; /* separation left between the shapes when they do not touch (<= 0: unknown / touching) */
  int count;
  d2 p1[2], p2[2];
  unsigned hash[2];
};
MG_HD unsigned feature_hash(unsigned a, unsigned b) { return (a << 16 | b) + 1u; }
MG_HD void manifold_push(Manifold& m, d2 p1, d2 p2, unsigned hash) {
  /* count is 0 or 1 here: written branch-free on the index to stay in registers */
  if (m.count == 0) { m.p1[0] = p1; m.p2[0] = p2; m.hash[0] = hash; }
  else { m.p1[1] = p1; m.p2[1] = p2; m.hash[1] = hash; }
  m.count++;
}
MG_HD void contact_points(SupEdge e1, SupEdge e2, const ClosestPts& points, Manifold& m) {
  double mindist = e1.r + e2.r;
  if (points.d <= mindist) {
    d2 n = m.n = points.n;
    double d_e1_a = dcross(e1.a, n);
    double d_e1_b = dcross(e1.b, n);
    double d_e2_a = dcross(e2.a, n);
    double d_e2_b = dcross(e2.b, n);
    double e1_denom = 1.0 / (d_e1_b - d_e1_a + MG_DBL_MIN);
    double e2_denom = 1.0 / (d_e2_b - d_e2_a + MG_DBL_MIN);
    {
      d2 p1 = dadd(dmul(n, e1.r), dlerp(e1.a, e1.b, dclamp01((d_e2_b - d_e1_a) * e1_denom)));
      d2 p2 = dadd(dmul(n, -e2.r), dlerp(e2.a, e2.b, dclamp01((d_e1_a - d_e2_a) * e2_denom)));
      double dist = ddot(dsub(p2, p1), n);
      if (dist <= 0.0) manifold_push(m, p1, p2, feature_hash(e1.ha, e2.hb));
    }
    {
      d2 p1 = dadd(dmul(n, e1.r), dlerp(e1.a, e1.b, dclamp01((d_e2_a - d_e1_a) * e1_denom)));
      d2 p2 = dadd(dmul(n, -e2.r), dlerp(e2.a, e2.b, dclamp01((d_e1_b - d_e2_a) * e2_denom)));
      double dist = ddot(dsub(p2, p1), n);
      if (dist <= 0.0) manifold_push(m, p1, p2, feature_hash(e1.hb, e2.ha));
    }
  }
}

/* Collide shapes a, b whose kinds are already ordered (a.kind <= b.kind).
 * bba/bbb: their exact bounding boxes (needed for the GJK start axis). */
MG_HD void mg_collide(const ShapeView& a, const ShapeView& b, const double* bba, const double* bbb, Manifold& m) {
  m.count = 0;
  m.n = D2(0, 0);
  m.margin = -1.0;
  if (a.kind == 0 && b.kind == 0) {
    double mindist = a.radius + b.radius;
    d2 ca = sv_vert(a, 0), cb = sv_vert(b, 0);
    d2 delta = dsub(cb, ca);
    double distsq = dlengthsq(delta);
    if (distsq < mindist * mindist) {
      double dist = sqrt(distsq);
      d2 n = m.n = (dist ? dmul(delta, 1.0 / dist) : D2(1.0, 0.0));
      manifold_push(m, dadd(ca, dmul(n, a.radius)), dadd(cb, dmul(n, -b.radius)), 0);
    } else {
      m.margin = sqrt(distsq) - mindist;
    }
  } else if (a.kind == 0 && b.kind == 1) {
    d2 seg_a = sv_vert(b, 0), seg_b = sv_vert(b, 1), center = sv_vert(a, 0);
    d2 seg_delta = dsub(seg_b, seg_a);
    double ct = dclamp01(ddot(seg_delta, dsub(center, seg_a)) / dlengthsq(seg_delta));
    d2 closest = dadd(seg_a, dmul(seg_delta, ct));
    double mindist = a.radius + b.radius;
    d2 delta = dsub(closest, center);
    double distsq = dlengthsq(delta);
    if (distsq < mindist * mindist) {
      double dist = sqrt(distsq);
      d2 n = m.n = (dist ? dmul(delta, 1.0 / dist) : sv_normal(b, 0));
      manifold_push(m, dadd(center, dmul(n, a.radius)), dadd(closest, dmul(n, -b.radius)), 0);
    } else {
      m.margin = sqrt(distsq) - mindist;
    }
  } else {
    /* every remaining pair runs GJK/EPA; ONE call site, so that lanes holding different kind pairs
     * (circle-poly, segment-poly, poly-poly) stay converged through the bulk of the work */
    ClosestPts pts = mg_gjk(a, b, bba, bbb);
    if (a.kind == 0) {
      if (pts.d <= a.radius + b.radius) {
        d2 n = m.n = pts.n;
        manifold_push(m, dadd(pts.a, dmul(n, a.radius)), dadd(pts.b, dmul(n, -b.radius)), 0);
      } else {
        m.margin = pts.d - (a.radius + b.radius);
      }
    } else if (pts.d - a.radius - b.radius <= 0.0) {
      SupEdge e1 = (a.kind == 1) ? support_edge_segment(a, pts.n) : support_edge_poly(a, pts.n);
      contact_points(e1, support_edge_poly(b, dneg(pts.n)), pts, m);
    } else {
      m.margin = pts.d - a.radius - b.radius;
    }
  }
}

#endif /* MG_NARROWPHASE_H */
