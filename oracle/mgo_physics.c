/*
 * mgo_physics.c — CPU ORACLE (test infrastructure, never on the product path).
 *
 * Plain-C, fp64, single-environment restatement of the physics the reference
 * delegates to pymunk 5.6 / Chipmunk2D 7.0.x: `pm.Space.step(dt)` as called
 * from `BaseEnv._phys_steps_on_frame` (magical/base_env.py:236-243), plus the
 * Python-side control law `Robot.update` (magical/entities.py:459-479) and
 * `Robot.set_action` (entities.py:439-457).
 *
 * PARITY UNPINNED: Chipmunk2D is a third-party dependency that is not vendored
 * under /root/reference and is not installable here, and the reference's own
 * tests hold no golden vectors for this path (tests/test_rollout_preproc.py:33
 * only asserts trajectory length).  This file therefore restates the
 * *published algorithm* of Chipmunk2D 7.0.3 (cpSpaceStep.c, cpArbiter.c,
 * cpCollision.c, cpPolyShape.c, cp*Joint.c, cpDampedRotarySpring.c,
 * cpSimpleMotor.c — summarised in SURVEY.md Appendix A) for exactly the
 * feature subset MAGICAL uses.  Documented deviations:
 *   - arbiter order is the canonical order of the compiled scene's collision
 *     pair list instead of cpBBTree traversal order (unknowable offline);
 *   - GJK always starts from the bounding-box-centre axis (no cached id);
 *   - sensor shapes (goal regions) are not in the space: they never reach the
 *     solver in Chipmunk either, and are only queried at score time.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.
 */
#include "mgo.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MGO_MAX_ARB 256
#define COLLISION_SLOP 0.01  /* base_env.py:195 */
#define ITERATIONS 10        /* base_env.py:196, benchmarks/__init__.py:404 */
#define PERSISTENCE 3        /* cpSpace default collisionPersistence */
#define DT (1.0 / 8.0 / 10.0) /* base_env.py:238-239 with fps=8 */

/* ------------------------------------------------------------------ vec */
static inline v2 V(double x, double y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline v2 vmult(v2 a, double s) { return V(a.x * s, a.y * s); }
static inline double vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline double vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 vperp(v2 a) { return V(-a.y, a.x); }
static inline v2 vrperp(v2 a) { return V(a.y, -a.x); }
static inline v2 vrotate(v2 a, v2 b) { return V(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
static inline double vlengthsq(v2 a) { return vdot(a, a); }
static inline double vlength(v2 a) { return sqrt(vdot(a, a)); }
static inline v2 vlerp(v2 a, v2 b, double t) { return vadd(vmult(a, 1.0 - t), vmult(b, t)); }
static inline v2 vnormalize(v2 a) { return vmult(a, 1.0 / (vlength(a) + MGO_DBL_MIN)); }
static inline int veql(v2 a, v2 b) { return a.x == b.x && a.y == b.y; }
static inline double fmax_(double a, double b) { return (a > b) ? a : b; }
static inline double fmin_(double a, double b) { return (a < b) ? a : b; }
static inline double fclamp(double f, double lo, double hi) { return fmin_(fmax_(f, lo), hi); }
static inline double fclamp01(double f) { return fmax_(0.0, fmin_(f, 1.0)); }
static inline v2 vclamp(v2 v, double len) {
  return (vdot(v, v) > len * len) ? vmult(vnormalize(v), len) : v;
}

/* sin/cos: libm by default; the deterministic polynomial shared with the CUDA
 * library when mode==1 (used by the bit-exact parity tests). */
#include "../magical_b200/csrc/mg_sincos.h"
static void rot_for_angle(const mgo_env* e, double a, v2* rot) {
  if (e->det_sincos) {
    double s, c;
    mg_det_sincos(a, &s, &c);
    rot->x = c;
    rot->y = s;
  } else {
    rot->x = cos(a);
    rot->y = sin(a);
  }
}

/* ----------------------------------------------------------------- body */
static mgo_body STATIC_BODY_TEMPLATE = {0};

static inline mgo_body* body_of(mgo_env* e, int idx) { return idx < 0 ? &e->static_body : &e->bodies[idx]; }

static inline v2 relative_velocity(const mgo_body* a, const mgo_body* b, v2 r1, v2 r2) {
  v2 v1 = vadd(a->v, vmult(vperp(r1), a->w));
  v2 v2_ = vadd(b->v, vmult(vperp(r2), b->w));
  return vsub(v2_, v1);
}
static inline double normal_relative_velocity(const mgo_body* a, const mgo_body* b, v2 r1, v2 r2, v2 n) {
  return vdot(relative_velocity(a, b, r1, r2), n);
}
static inline void apply_impulse(mgo_body* body, v2 j, v2 r) {
  body->v = vadd(body->v, vmult(j, body->m_inv));
  body->w += body->i_inv * vcross(r, j);
}
static inline void apply_impulses(mgo_body* a, mgo_body* b, v2 r1, v2 r2, v2 j) {
  apply_impulse(a, vneg(j), r1);
  apply_impulse(b, j, r2);
}
static inline void apply_bias_impulse(mgo_body* body, v2 j, v2 r) {
  body->v_bias = vadd(body->v_bias, vmult(j, body->m_inv));
  body->w_bias += body->i_inv * vcross(r, j);
}
static inline void apply_bias_impulses(mgo_body* a, mgo_body* b, v2 r1, v2 r2, v2 j) {
  apply_bias_impulse(a, vneg(j), r1);
  apply_bias_impulse(b, j, r2);
}
static inline double k_scalar_body(const mgo_body* body, v2 r, v2 n) {
  double rcn = vcross(r, n);
  return body->m_inv + body->i_inv * rcn * rcn;
}
static inline double k_scalar(const mgo_body* a, const mgo_body* b, v2 r1, v2 r2, v2 n) {
  return k_scalar_body(a, r1, n) + k_scalar_body(b, r2, n);
}
typedef struct { double a, b, c, d; } mat2;
static mat2 k_tensor(const mgo_body* a, const mgo_body* b, v2 r1, v2 r2) {
  double m_sum = a->m_inv + b->m_inv;
  double k11 = m_sum, k12 = 0.0, k21 = 0.0, k22 = m_sum;
  double a_i_inv = a->i_inv;
  double r1xsq = r1.x * r1.x * a_i_inv;
  double r1ysq = r1.y * r1.y * a_i_inv;
  double r1nxy = -r1.x * r1.y * a_i_inv;
  k11 += r1ysq; k12 += r1nxy; k21 += r1nxy; k22 += r1xsq;
  double b_i_inv = b->i_inv;
  double r2xsq = r2.x * r2.x * b_i_inv;
  double r2ysq = r2.y * r2.y * b_i_inv;
  double r2nxy = -r2.x * r2.y * b_i_inv;
  k11 += r2ysq; k12 += r2nxy; k21 += r2nxy; k22 += r2xsq;
  double det = k11 * k22 - k12 * k21;
  double det_inv = 1.0 / det;
  mat2 m = {k22 * det_inv, -k12 * det_inv, -k21 * det_inv, k11 * det_inv};
  return m;
}
static inline v2 mat2_transform(mat2 m, v2 v) { return V(v.x * m.a + v.y * m.b, v.x * m.c + v.y * m.d); }
static inline double bias_coef(double error_bias, double dt) { return 1.0 - pow(error_bias, dt); }

static inline v2 body_local_to_world_vect(const mgo_body* b, v2 v) {
  /* cpTransformVect with the body's rotation (cog = 0 for every MAGICAL body) */
  return V(b->rot.x * v.x - b->rot.y * v.y, b->rot.y * v.x + b->rot.x * v.y);
}
static inline v2 body_local_to_world_point(const mgo_body* b, v2 v) {
  return V(b->rot.x * v.x - b->rot.y * v.y + b->p.x, b->rot.y * v.x + b->rot.x * v.y + b->p.y);
}

/* --------------------------------------------------------------- shapes */
static void shape_cache(mgo_env* e, mgo_shape* s) {
  const mgo_body* b = body_of(e, s->body);
  double l, r, bt, t;
  if (s->kind == MG_SHAPE_CIRCLE) {
    s->tv[0] = body_local_to_world_point(b, s->lv[0]);
    l = s->tv[0].x - s->radius; r = s->tv[0].x + s->radius;
    bt = s->tv[0].y - s->radius; t = s->tv[0].y + s->radius;
  } else if (s->kind == MG_SHAPE_SEGMENT) {
    s->tv[0] = body_local_to_world_point(b, s->lv[0]);
    s->tv[1] = body_local_to_world_point(b, s->lv[1]);
    s->tn[0] = body_local_to_world_vect(b, s->ln[0]);
    v2 ta = s->tv[0], tb = s->tv[1];
    if (ta.x < tb.x) { l = ta.x; r = tb.x; } else { l = tb.x; r = ta.x; }
    if (ta.y < tb.y) { bt = ta.y; t = tb.y; } else { bt = tb.y; t = ta.y; }
    l -= s->radius; bt -= s->radius; r += s->radius; t += s->radius;
  } else {
    l = INFINITY; r = -INFINITY; bt = INFINITY; t = -INFINITY;
    for (int i = 0; i < s->nvert; i++) {
      v2 v = body_local_to_world_point(b, s->lv[i]);
      s->tv[i] = v;
      s->tn[i] = body_local_to_world_vect(b, s->ln[i]);
      l = fmin_(l, v.x); r = fmax_(r, v.x); bt = fmin_(bt, v.y); t = fmax_(t, v.y);
    }
    l -= s->radius; bt -= s->radius; r += s->radius; t += s->radius;
  }
  s->bb[0] = l; s->bb[1] = bt; s->bb[2] = r; s->bb[3] = t;
}

static inline int bb_intersects(const double* a, const double* b) {
  return a[0] <= b[2] && b[0] <= a[2] && a[1] <= b[3] && b[1] <= a[3];
}

/* ----------------------------------------------------- narrowphase (GJK) */
typedef struct { v2 p; unsigned index; } SupportPoint;
typedef struct { v2 a, b, ab; unsigned id; } MinkowskiPoint;
typedef struct { v2 a, b, n; double d; unsigned id; } ClosestPoints;
typedef struct { const mgo_shape *s1, *s2; } SupportContext;
typedef struct { v2 p; unsigned hash; } EdgePoint;
typedef struct { EdgePoint a, b; double r; v2 n; } Edge;

static int poly_support_index(const mgo_shape* s, v2 n) {
  double max = -INFINITY;
  int index = 0;
  for (int i = 0; i < s->nvert; i++) {
    double d = vdot(s->tv[i], n);
    if (d > max) { max = d; index = i; }
  }
  return index;
}
static SupportPoint support_point(const mgo_shape* s, v2 n) {
  SupportPoint sp;
  if (s->kind == MG_SHAPE_CIRCLE) {
    sp.p = s->tv[0]; sp.index = 0;
  } else if (s->kind == MG_SHAPE_SEGMENT) {
    if (vdot(s->tv[0], n) > vdot(s->tv[1], n)) { sp.p = s->tv[0]; sp.index = 0; }
    else { sp.p = s->tv[1]; sp.index = 1; }
  } else {
    int i = poly_support_index(s, n);
    sp.p = s->tv[i]; sp.index = (unsigned)i;
  }
  return sp;
}
static MinkowskiPoint minkowski_new(SupportPoint a, SupportPoint b) {
  MinkowskiPoint m = {a.p, b.p, vsub(b.p, a.p), (a.index & 0xFF) << 8 | (b.index & 0xFF)};
  return m;
}
static MinkowskiPoint support(const SupportContext* ctx, v2 n) {
  SupportPoint a = support_point(ctx->s1, vneg(n));
  SupportPoint b = support_point(ctx->s2, n);
  return minkowski_new(a, b);
}
static Edge support_edge_for_poly(const mgo_shape* poly, v2 n, int shape_idx) {
  int count = poly->nvert;
  int i1 = poly_support_index(poly, n);
  int i0 = (i1 - 1 + count) % count;
  int i2 = (i1 + 1) % count;
  unsigned base = (unsigned)shape_idx << 8;
  Edge e;
  if (vdot(n, poly->tn[i1]) > vdot(n, poly->tn[i2])) {
    e.a.p = poly->tv[i0]; e.a.hash = base | (unsigned)i0;
    e.b.p = poly->tv[i1]; e.b.hash = base | (unsigned)i1;
    e.r = poly->radius; e.n = poly->tn[i1];
  } else {
    e.a.p = poly->tv[i1]; e.a.hash = base | (unsigned)i1;
    e.b.p = poly->tv[i2]; e.b.hash = base | (unsigned)i2;
    e.r = poly->radius; e.n = poly->tn[i2];
  }
  return e;
}
static Edge support_edge_for_segment(const mgo_shape* seg, v2 n, int shape_idx) {
  unsigned base = (unsigned)shape_idx << 8;
  Edge e;
  if (vdot(seg->tn[0], n) > 0.0) {
    e.a.p = seg->tv[0]; e.a.hash = base | 0u;
    e.b.p = seg->tv[1]; e.b.hash = base | 1u;
    e.r = seg->radius; e.n = seg->tn[0];
  } else {
    e.a.p = seg->tv[1]; e.a.hash = base | 1u;
    e.b.p = seg->tv[0]; e.b.hash = base | 0u;
    e.r = seg->radius; e.n = vneg(seg->tn[0]);
  }
  return e;
}
static inline double closest_t(v2 a, v2 b) {
  v2 delta = vsub(b, a);
  return -fclamp(vdot(delta, vadd(a, b)) / vlengthsq(delta), -1.0, 1.0);
}
static inline v2 lerp_t(v2 a, v2 b, double t) {
  double ht = 0.5 * t;
  return vadd(vmult(a, 0.5 - ht), vmult(b, 0.5 + ht));
}
static inline double closest_dist(v2 v0, v2 v1) { return vlengthsq(lerp_t(v0, v1, closest_t(v0, v1))); }
static inline int check_point_greater(v2 a, v2 b, v2 c) {
  return (b.y - a.y) * (a.x + b.x - 2 * c.x) > (b.x - a.x) * (a.y + b.y - 2 * c.y);
}
static inline int check_axis(v2 v0, v2 v1, v2 p, v2 n) { return vdot(p, n) <= fmax_(vdot(v0, n), vdot(v1, n)); }

static ClosestPoints closest_points_new(MinkowskiPoint v0, MinkowskiPoint v1) {
  double t = closest_t(v0.ab, v1.ab);
  v2 p = lerp_t(v0.ab, v1.ab, t);
  v2 pa = lerp_t(v0.a, v1.a, t);
  v2 pb = lerp_t(v0.b, v1.b, t);
  unsigned id = (v0.id & 0xFFFF) << 16 | (v1.id & 0xFFFF);
  v2 delta = vsub(v1.ab, v0.ab);
  v2 n = vnormalize(vrperp(delta));
  double d = vdot(n, p);
  ClosestPoints pts;
  if (d <= 0.0 || (-1.0 < t && t < 1.0)) {
    pts.a = pa; pts.b = pb; pts.n = n; pts.d = d; pts.id = id;
  } else {
    double d2 = vlength(p);
    v2 n2 = vmult(p, 1.0 / (d2 + MGO_DBL_MIN));
    pts.a = pa; pts.b = pb; pts.n = n2; pts.d = d2; pts.id = id;
  }
  return pts;
}

#define MAX_EPA_ITERATIONS 30
#define MAX_GJK_ITERATIONS 30

static ClosestPoints epa_recurse(const SupportContext* ctx, int count, const MinkowskiPoint* hull, int iteration) {
  int mini = 0;
  double min_dist = INFINITY;
  for (int j = 0, i = count - 1; j < count; i = j, j++) {
    double d = closest_dist(hull[i].ab, hull[j].ab);
    if (d < min_dist) { min_dist = d; mini = i; }
  }
  MinkowskiPoint v0 = hull[mini];
  MinkowskiPoint v1 = hull[(mini + 1) % count];
  MinkowskiPoint p = support(ctx, vperp(vsub(v1.ab, v0.ab)));
  int duplicate = (p.id == v0.id || p.id == v1.id);
  if (!duplicate && check_point_greater(v0.ab, v1.ab, p.ab) && iteration < MAX_EPA_ITERATIONS) {
    MinkowskiPoint hull2[MAX_EPA_ITERATIONS + 8];
    int count2 = 1;
    hull2[0] = p;
    for (int i = 0; i < count; i++) {
      int index = (mini + 1 + i) % count;
      v2 h0 = hull2[count2 - 1].ab;
      v2 h1 = hull[index].ab;
      v2 h2 = (i + 1 < count ? hull[(index + 1) % count] : p).ab;
      if (check_point_greater(h0, h2, h1)) { hull2[count2] = hull[index]; count2++; }
    }
    return epa_recurse(ctx, count2, hull2, iteration + 1);
  }
  return closest_points_new(v0, v1);
}
static ClosestPoints epa(const SupportContext* ctx, MinkowskiPoint v0, MinkowskiPoint v1, MinkowskiPoint v2_) {
  MinkowskiPoint hull[3] = {v0, v1, v2_};
  return epa_recurse(ctx, 3, hull, 1);
}
static ClosestPoints gjk_recurse(const SupportContext* ctx, MinkowskiPoint v0, MinkowskiPoint v1, int iteration) {
  if (iteration > MAX_GJK_ITERATIONS) return closest_points_new(v0, v1);
  if (check_point_greater(v1.ab, v0.ab, V(0, 0))) {
    return gjk_recurse(ctx, v1, v0, iteration);
  } else {
    double t = closest_t(v0.ab, v1.ab);
    v2 n = (-1.0 < t && t < 1.0 ? vperp(vsub(v1.ab, v0.ab)) : vneg(lerp_t(v0.ab, v1.ab, t)));
    MinkowskiPoint p = support(ctx, n);
    if (check_point_greater(p.ab, v0.ab, V(0, 0)) && check_point_greater(v1.ab, p.ab, V(0, 0))) {
      return epa(ctx, v0, p, v1);
    } else {
      if (check_axis(v0.ab, v1.ab, p.ab, n)) {
        return closest_points_new(v0, v1);
      } else {
        if (closest_dist(v0.ab, p.ab) < closest_dist(p.ab, v1.ab)) return gjk_recurse(ctx, v0, p, iteration + 1);
        else return gjk_recurse(ctx, p, v1, iteration + 1);
      }
    }
  }
}
static ClosestPoints gjk(const SupportContext* ctx) {
  const double* b1 = ctx->s1->bb;
  const double* b2 = ctx->s2->bb;
  v2 c1 = vlerp(V(b1[0], b1[1]), V(b1[2], b1[3]), 0.5);
  v2 c2 = vlerp(V(b2[0], b2[1]), V(b2[2], b2[3]), 0.5);
  v2 axis = vperp(vsub(c1, c2));
  MinkowskiPoint v0 = support(ctx, axis);
  MinkowskiPoint v1 = support(ctx, vneg(axis));
  return gjk_recurse(ctx, v0, v1, 1);
}

typedef struct {
  v2 n;
  int count;
  struct { v2 p1, p2; unsigned hash; } arr[2];
} CollisionInfo;

static inline void push_contact(CollisionInfo* info, v2 p1, v2 p2, unsigned hash) {
  info->arr[info->count].p1 = p1;
  info->arr[info->count].p2 = p2;
  info->arr[info->count].hash = hash;
  info->count++;
}
static inline unsigned hash_pair(unsigned a, unsigned b) {
  /* any injective labelling of (feature on a, feature on b) serves: hashes are
   * only compared for equality inside one arbiter. +1 keeps it distinct from
   * the hash 0 used by single-contact (circle) collisions. */
  return (a << 16 | b) + 1u;
}
static void contact_points(Edge e1, Edge e2, ClosestPoints points, CollisionInfo* info) {
  double mindist = e1.r + e2.r;
  if (points.d <= mindist) {
    v2 n = info->n = points.n;
    double d_e1_a = vcross(e1.a.p, n);
    double d_e1_b = vcross(e1.b.p, n);
    double d_e2_a = vcross(e2.a.p, n);
    double d_e2_b = vcross(e2.b.p, n);
    double e1_denom = 1.0 / (d_e1_b - d_e1_a + MGO_DBL_MIN);
    double e2_denom = 1.0 / (d_e2_b - d_e2_a + MGO_DBL_MIN);
    {
      v2 p1 = vadd(vmult(n, e1.r), vlerp(e1.a.p, e1.b.p, fclamp01((d_e2_b - d_e1_a) * e1_denom)));
      v2 p2 = vadd(vmult(n, -e2.r), vlerp(e2.a.p, e2.b.p, fclamp01((d_e1_a - d_e2_a) * e2_denom)));
      double dist = vdot(vsub(p2, p1), n);
      if (dist <= 0.0) push_contact(info, p1, p2, hash_pair(e1.a.hash, e2.b.hash));
    }
    {
      v2 p1 = vadd(vmult(n, e1.r), vlerp(e1.a.p, e1.b.p, fclamp01((d_e2_a - d_e1_a) * e1_denom)));
      v2 p2 = vadd(vmult(n, -e2.r), vlerp(e2.a.p, e2.b.p, fclamp01((d_e1_b - d_e2_a) * e2_denom)));
      double dist = vdot(vsub(p2, p1), n);
      if (dist <= 0.0) push_contact(info, p1, p2, hash_pair(e1.b.hash, e2.a.hash));
    }
  }
}

static void circle_to_circle(const mgo_shape* c1, const mgo_shape* c2, CollisionInfo* info) {
  double mindist = c1->radius + c2->radius;
  v2 delta = vsub(c2->tv[0], c1->tv[0]);
  double distsq = vlengthsq(delta);
  if (distsq < mindist * mindist) {
    double dist = sqrt(distsq);
    v2 n = info->n = (dist ? vmult(delta, 1.0 / dist) : V(1.0, 0.0));
    push_contact(info, vadd(c1->tv[0], vmult(n, c1->radius)), vadd(c2->tv[0], vmult(n, -c2->radius)), 0);
  }
}
static void circle_to_segment(const mgo_shape* circle, const mgo_shape* seg, CollisionInfo* info) {
  v2 seg_a = seg->tv[0], seg_b = seg->tv[1], center = circle->tv[0];
  v2 seg_delta = vsub(seg_b, seg_a);
  double closest_t_ = fclamp01(vdot(seg_delta, vsub(center, seg_a)) / vlengthsq(seg_delta));
  v2 closest = vadd(seg_a, vmult(seg_delta, closest_t_));
  double mindist = circle->radius + seg->radius;
  v2 delta = vsub(closest, center);
  double distsq = vlengthsq(delta);
  if (distsq < mindist * mindist) {
    double dist = sqrt(distsq);
    v2 n = info->n = (dist ? vmult(delta, 1.0 / dist) : seg->tn[0]);
    /* segment end-cap tangents are zero (never set by the reference): always accept */
    push_contact(info, vadd(center, vmult(n, circle->radius)), vadd(closest, vmult(n, -seg->radius)), 0);
  }
}
static void circle_to_poly(const mgo_shape* circle, const mgo_shape* poly, CollisionInfo* info) {
  SupportContext ctx = {circle, poly};
  ClosestPoints points = gjk(&ctx);
  if (points.d <= circle->radius + poly->radius) {
    v2 n = info->n = points.n;
    push_contact(info, vadd(points.a, vmult(n, circle->radius)), vadd(points.b, vmult(n, -poly->radius)), 0);
  }
}
static void segment_to_poly(const mgo_shape* seg, int seg_idx, const mgo_shape* poly, int poly_idx, CollisionInfo* info) {
  SupportContext ctx = {seg, poly};
  ClosestPoints points = gjk(&ctx);
  v2 n = points.n;
  if (points.d - seg->radius - poly->radius <= 0.0) {
    contact_points(support_edge_for_segment(seg, n, seg_idx), support_edge_for_poly(poly, vneg(n), poly_idx), points, info);
  }
}
static void poly_to_poly(const mgo_shape* p1, int i1, const mgo_shape* p2, int i2, CollisionInfo* info) {
  SupportContext ctx = {p1, p2};
  ClosestPoints points = gjk(&ctx);
  if (points.d - p1->radius - p2->radius <= 0.0) {
    contact_points(support_edge_for_poly(p1, points.n, i1), support_edge_for_poly(p2, vneg(points.n), i2), points, info);
  }
}

/* cpCollide: orders the pair by shape type and dispatches. Returns the (a, b) order used. */
void mgo_collide(const mgo_env* e, int ia, int ib, int* out_a, int* out_b, v2* n, int* count, v2 p1[2], v2 p2[2],
                 unsigned hash[2]) {
  if (e->shapes[ia].kind > e->shapes[ib].kind) { int t = ia; ia = ib; ib = t; }
  const mgo_shape* a = &e->shapes[ia];
  const mgo_shape* b = &e->shapes[ib];
  CollisionInfo info;
  info.count = 0;
  info.n = V(0, 0);
  if (a->kind == MG_SHAPE_CIRCLE && b->kind == MG_SHAPE_CIRCLE) circle_to_circle(a, b, &info);
  else if (a->kind == MG_SHAPE_CIRCLE && b->kind == MG_SHAPE_SEGMENT) circle_to_segment(a, b, &info);
  else if (a->kind == MG_SHAPE_CIRCLE && b->kind == MG_SHAPE_POLY) circle_to_poly(a, b, &info);
  else if (a->kind == MG_SHAPE_SEGMENT && b->kind == MG_SHAPE_POLY) segment_to_poly(a, ia, b, ib, &info);
  else if (a->kind == MG_SHAPE_POLY && b->kind == MG_SHAPE_POLY) poly_to_poly(a, ia, b, ib, &info);
  *out_a = ia; *out_b = ib; *n = info.n; *count = info.count;
  for (int i = 0; i < info.count; i++) { p1[i] = info.arr[i].p1; p2[i] = info.arr[i].p2; hash[i] = info.arr[i].hash; }
}

/* -------------------------------------------------------------- arbiters */
static mgo_arbiter* arbiter_find(mgo_env* e, int a, int b, int create) {
  for (int i = 0; i < e->n_cached; i++) {
    if (e->cached[i].a == a && e->cached[i].b == b) return &e->cached[i];
  }
  if (!create) return NULL;
  if (e->n_cached >= MGO_MAX_ARBITERS) { e->overflow = 1; return NULL; }
  mgo_arbiter* arb = &e->cached[e->n_cached++];
  memset(arb, 0, sizeof(*arb));
  arb->a = a; arb->b = b;
  arb->state = MGO_ARB_FIRST;
  return arb;
}

static void collide_shapes(mgo_env* e, int sa, int sb) {
  mgo_shape* A = &e->shapes[sa];
  mgo_shape* B = &e->shapes[sb];
  /* QueryReject: bounding boxes, same body, filter group */
  if (!bb_intersects(A->bb, B->bb)) return;
  if (A->body == B->body) return;
  if (A->group != 0 && A->group == B->group) return;
  e->stat_bb_pass++;
  {
    /* instrumentation only: would a bounding-circle test have culled this pair? */
    double ca[2] = {0.5 * (A->bb[0] + A->bb[2]), 0.5 * (A->bb[1] + A->bb[3])};
    double cb[2] = {0.5 * (B->bb[0] + B->bb[2]), 0.5 * (B->bb[1] + B->bb[3])};
    double ra = 0, rb = 0;
    for (int k = 0; k < A->nvert; k++) { double d = hypot(A->tv[k].x - ca[0], A->tv[k].y - ca[1]); if (d > ra) ra = d; }
    for (int k = 0; k < B->nvert; k++) { double d = hypot(B->tv[k].x - cb[0], B->tv[k].y - cb[1]); if (d > rb) rb = d; }
    if (A->kind != MG_SHAPE_SEGMENT && B->kind != MG_SHAPE_SEGMENT) {
      if (hypot(ca[0] - cb[0], ca[1] - cb[1]) <= ra + rb + A->radius + B->radius) e->stat_circle_pass++;
    } else {
      e->stat_circle_pass++;
    }
  }
  int ia, ib, count;
  v2 n, p1[2], p2[2];
  unsigned hash[2];
  mgo_collide(e, sa, sb, &ia, &ib, &n, &count, p1, p2, hash);
  if (count == 0) return;
  e->stat_hits++;
  e->stat_contacts += count;
  mgo_arbiter* arb = arbiter_find(e, ia, ib, 1);
  if (!arb) return;
  /* cpArbiterUpdate */
  mgo_body* ba = body_of(e, e->shapes[ia].body);
  mgo_body* bb = body_of(e, e->shapes[ib].body);
  mgo_contact fresh[2];
  for (int i = 0; i < count; i++) {
    mgo_contact* con = &fresh[i];
    memset(con, 0, sizeof(*con));
    con->r1 = vsub(p1[i], ba->p);
    con->r2 = vsub(p2[i], bb->p);
    con->hash = hash[i];
    con->jnAcc = con->jtAcc = 0.0;
    for (int j = 0; j < arb->count; j++) {
      if (con->hash == arb->contacts[j].hash) {
        con->jnAcc = arb->contacts[j].jnAcc;
        con->jtAcc = arb->contacts[j].jtAcc;
      }
    }
  }
  for (int i = 0; i < count; i++) arb->contacts[i] = fresh[i];
  arb->count = count;
  arb->n = n;
  arb->u = e->shapes[ia].friction * e->shapes[ib].friction;
  arb->body_a = e->shapes[ia].body;
  arb->body_b = e->shapes[ib].body;
  if (arb->state == MGO_ARB_CACHED) arb->state = MGO_ARB_FIRST;
  /* both-infinite-mass pairs never reach the solver (walls vs the kinematic control body have no
   * shapes here, so this cannot trigger; kept for fidelity) */
  if (!(ba->m_inv == 0.0 && bb->m_inv == 0.0)) {
    if (e->n_active < MGO_MAX_ARBITERS) e->active[e->n_active++] = (int)(arb - e->cached);
  }
  arb->stamp = e->stamp;
}

static void arbiter_prestep(mgo_env* e, mgo_arbiter* arb, double dt, double slop, double bias) {
  mgo_body* a = body_of(e, arb->body_a);
  mgo_body* b = body_of(e, arb->body_b);
  v2 n = arb->n;
  v2 body_delta = vsub(b->p, a->p);
  for (int i = 0; i < arb->count; i++) {
    mgo_contact* con = &arb->contacts[i];
    con->nMass = 1.0 / k_scalar(a, b, con->r1, con->r2, n);
    con->tMass = 1.0 / k_scalar(a, b, con->r1, con->r2, vperp(n));
    double dist = vdot(vadd(vsub(con->r2, con->r1), body_delta), n);
    con->bias = -bias * fmin_(0.0, dist + slop) / dt;
    con->jBias = 0.0;
    con->bounce = normal_relative_velocity(a, b, con->r1, con->r2, n) * 0.0; /* e = 0 */
  }
}
static void arbiter_apply_cached(mgo_env* e, mgo_arbiter* arb, double dt_coef) {
  if (arb->state == MGO_ARB_FIRST) return;
  mgo_body* a = body_of(e, arb->body_a);
  mgo_body* b = body_of(e, arb->body_b);
  v2 n = arb->n;
  for (int i = 0; i < arb->count; i++) {
    mgo_contact* con = &arb->contacts[i];
    v2 j = vrotate(n, V(con->jnAcc, con->jtAcc));
    apply_impulses(a, b, con->r1, con->r2, vmult(j, dt_coef));
  }
}
static void arbiter_apply_impulse(mgo_env* e, mgo_arbiter* arb) {
  mgo_body* a = body_of(e, arb->body_a);
  mgo_body* b = body_of(e, arb->body_b);
  v2 n = arb->n;
  double friction = arb->u;
  for (int i = 0; i < arb->count; i++) {
    mgo_contact* con = &arb->contacts[i];
    double nMass = con->nMass;
    v2 r1 = con->r1, r2 = con->r2;
    v2 vb1 = vadd(a->v_bias, vmult(vperp(r1), a->w_bias));
    v2 vb2 = vadd(b->v_bias, vmult(vperp(r2), b->w_bias));
    v2 vr = relative_velocity(a, b, r1, r2); /* surface_vr = 0 */
    double vbn = vdot(vsub(vb2, vb1), n);
    double vrn = vdot(vr, n);
    double vrt = vdot(vr, vperp(n));
    double jbn = (con->bias - vbn) * nMass;
    double jbnOld = con->jBias;
    con->jBias = fmax_(jbnOld + jbn, 0.0);
    double jn = -(con->bounce + vrn) * nMass;
    double jnOld = con->jnAcc;
    con->jnAcc = fmax_(jnOld + jn, 0.0);
    double jtMax = friction * con->jnAcc;
    double jt = -vrt * con->tMass;
    double jtOld = con->jtAcc;
    con->jtAcc = fclamp(jtOld + jt, -jtMax, jtMax);
    apply_bias_impulses(a, b, r1, r2, vmult(n, con->jBias - jbnOld));
    apply_impulses(a, b, r1, r2, vrotate(n, V(con->jnAcc - jnOld, con->jtAcc - jtOld)));
  }
}

/* ---------------------------------------------------------------- joints */
static void joint_prestep(mgo_env* e, mgo_joint* j, double dt) {
  mgo_body* a = body_of(e, j->a);
  mgo_body* b = body_of(e, j->b);
  switch (j->kind) {
    case MG_JOINT_PIVOT: {
      j->r1 = body_local_to_world_vect(a, j->anchor_a);
      j->r2 = body_local_to_world_vect(b, j->anchor_b);
      mat2 k = k_tensor(a, b, j->r1, j->r2);
      j->k[0] = k.a; j->k[1] = k.b; j->k[2] = k.c; j->k[3] = k.d;
      v2 delta = vsub(vadd(b->p, j->r2), vadd(a->p, j->r1));
      j->bias_v = vclamp(vmult(delta, -bias_coef(j->error_bias, dt) / dt), j->max_bias);
    } break;
    case MG_JOINT_GEAR: {
      double ratio = j->p1, ratio_inv = 1.0 / j->p1;
      j->iSum = 1.0 / (a->i_inv * ratio_inv + ratio * b->i_inv);
      double maxBias = j->max_bias;
      j->bias = fclamp(-bias_coef(j->error_bias, dt) * (b->a * ratio - a->a - j->p0) / dt, -maxBias, maxBias);
    } break;
    case MG_JOINT_ROTARY_SPRING: {
      double moment = a->i_inv + b->i_inv;
      j->iSum = 1.0 / moment;
      j->w_coef = 1.0 - exp(-j->p2 * dt * moment);
      j->target_wrn = 0.0;
      double j_spring = ((a->a - b->a) - j->p0) * j->p1 * dt;
      j->jAcc.x = j_spring;
      a->w -= j_spring * a->i_inv;
      b->w += j_spring * b->i_inv;
    } break;
    case MG_JOINT_PIN: {
      j->r1 = body_local_to_world_vect(a, j->anchor_a);
      j->r2 = body_local_to_world_vect(b, j->anchor_b);
      v2 delta = vsub(vadd(b->p, j->r2), vadd(a->p, j->r1));
      double dist = vlength(delta);
      j->n = vmult(delta, 1.0 / (dist ? dist : (double)INFINITY));
      j->nMass = 1.0 / k_scalar(a, b, j->r1, j->r2, j->n);
      double maxBias = j->max_bias;
      j->bias = fclamp(-bias_coef(j->error_bias, dt) * (dist - j->p0) / dt, -maxBias, maxBias);
    } break;
    case MG_JOINT_ROTARY_LIMIT: {
      double dist = b->a - a->a;
      double pdist = 0.0;
      if (dist > j->p1) pdist = j->p1 - dist;
      else if (dist < j->p0) pdist = j->p0 - dist;
      j->iSum = 1.0 / (a->i_inv + b->i_inv);
      double maxBias = j->max_bias;
      j->bias = fclamp(-bias_coef(j->error_bias, dt) * pdist / dt, -maxBias, maxBias);
      if (!j->bias) j->jAcc.x = 0.0;
    } break;
    case MG_JOINT_MOTOR: {
      j->iSum = 1.0 / (a->i_inv + b->i_inv);
    } break;
  }
}
static void joint_apply_cached(mgo_env* e, mgo_joint* j, double dt_coef) {
  mgo_body* a = body_of(e, j->a);
  mgo_body* b = body_of(e, j->b);
  switch (j->kind) {
    case MG_JOINT_PIVOT:
      apply_impulses(a, b, j->r1, j->r2, vmult(j->jAcc, dt_coef));
      break;
    case MG_JOINT_GEAR: {
      double jj = j->jAcc.x * dt_coef;
      a->w -= jj * a->i_inv * (1.0 / j->p1);
      b->w += jj * b->i_inv;
    } break;
    case MG_JOINT_ROTARY_SPRING:
      break;
    case MG_JOINT_PIN:
      apply_impulses(a, b, j->r1, j->r2, vmult(j->n, j->jAcc.x * dt_coef));
      break;
    case MG_JOINT_ROTARY_LIMIT:
    case MG_JOINT_MOTOR: {
      double jj = j->jAcc.x * dt_coef;
      a->w -= jj * a->i_inv;
      b->w += jj * b->i_inv;
    } break;
  }
}
static void joint_apply_impulse(mgo_env* e, mgo_joint* j, double dt) {
  mgo_body* a = body_of(e, j->a);
  mgo_body* b = body_of(e, j->b);
  switch (j->kind) {
    case MG_JOINT_PIVOT: {
      v2 vr = relative_velocity(a, b, j->r1, j->r2);
      mat2 k = {j->k[0], j->k[1], j->k[2], j->k[3]};
      v2 jj = mat2_transform(k, vsub(j->bias_v, vr));
      v2 jOld = j->jAcc;
      j->jAcc = vclamp(vadd(j->jAcc, jj), j->max_force * dt);
      jj = vsub(j->jAcc, jOld);
      apply_impulses(a, b, j->r1, j->r2, jj);
    } break;
    case MG_JOINT_GEAR: {
      double ratio = j->p1, ratio_inv = 1.0 / j->p1;
      double wr = b->w * ratio - a->w;
      double jMax = j->max_force * dt;
      double jj = (j->bias - wr) * j->iSum;
      double jOld = j->jAcc.x;
      j->jAcc.x = fclamp(jOld + jj, -jMax, jMax);
      jj = j->jAcc.x - jOld;
      a->w -= jj * a->i_inv * ratio_inv;
      b->w += jj * b->i_inv;
    } break;
    case MG_JOINT_ROTARY_SPRING: {
      double wrn = a->w - b->w;
      double w_damp = (j->target_wrn - wrn) * j->w_coef;
      j->target_wrn = wrn + w_damp;
      double j_damp = w_damp * j->iSum;
      j->jAcc.x += j_damp;
      a->w += j_damp * a->i_inv;
      b->w -= j_damp * b->i_inv;
    } break;
    case MG_JOINT_PIN: {
      v2 n = j->n;
      double vrn = normal_relative_velocity(a, b, j->r1, j->r2, n);
      double jnMax = j->max_force * dt;
      double jn = (j->bias - vrn) * j->nMass;
      double jnOld = j->jAcc.x;
      j->jAcc.x = fclamp(jnOld + jn, -jnMax, jnMax);
      jn = j->jAcc.x - jnOld;
      apply_impulses(a, b, j->r1, j->r2, vmult(n, jn));
    } break;
    case MG_JOINT_ROTARY_LIMIT: {
      if (!j->bias) return;
      double wr = b->w - a->w;
      double jMax = j->max_force * dt;
      double jj = -(j->bias + wr) * j->iSum;
      double jOld = j->jAcc.x;
      if (j->bias < 0.0) j->jAcc.x = fclamp(jOld + jj, 0.0, jMax);
      else j->jAcc.x = fclamp(jOld + jj, -jMax, 0.0);
      jj = j->jAcc.x - jOld;
      a->w -= jj * a->i_inv;
      b->w += jj * b->i_inv;
    } break;
    case MG_JOINT_MOTOR: {
      double wr = b->w - a->w + j->rate;
      double jMax = j->max_force * dt;
      double jj = -wr * j->iSum;
      double jOld = j->jAcc.x;
      j->jAcc.x = fclamp(jOld + jj, -jMax, jMax);
      jj = j->jAcc.x - jOld;
      a->w -= jj * a->i_inv;
      b->w += jj * b->i_inv;
    } break;
  }
}

/* ------------------------------------------------------------ space step */
void mgo_get_stats(const mgo_env* e, long out[5]) {
  out[0] = e->stat_substeps; out[1] = e->stat_bb_pass; out[2] = e->stat_circle_pass; out[3] = e->stat_hits;
  out[4] = e->stat_contacts;
}

void mgo_space_step(mgo_env* e, double dt) {
  e->stat_substeps++;
  e->stamp++;
  double prev_dt = e->curr_dt;
  e->curr_dt = dt;

  /* arbiters active last step go back to NORMAL; list is rebuilt */
  for (int i = 0; i < e->n_active; i++) e->cached[e->active[i]].state = MGO_ARB_NORMAL;
  e->n_active = 0;

  /* integrate positions (dynamic AND kinematic bodies: cpBodyUpdatePosition) */
  for (int i = 0; i < e->n_bodies; i++) {
    mgo_body* b = &e->bodies[i];
    b->p = vadd(b->p, vmult(vadd(b->v, b->v_bias), dt));
    b->a = b->a + (b->w + b->w_bias) * dt;
    rot_for_angle(e, b->a, &b->rot);
    b->v_bias = V(0, 0);
    b->w_bias = 0.0;
  }
  /* update shape caches, then collide in the scene's canonical pair order */
  for (int i = 0; i < e->n_shapes; i++) {
    if (e->shapes[i].body >= 0) shape_cache(e, &e->shapes[i]);
  }
  for (int p = 0; p < e->scene.n_bpairs; p++) {
    int pp = p;
    if (e->pair_perm) pp = e->pair_perm[p];
    const mg_cgroup_t* ga = &e->scene.cgroups[e->scene.bpairs[pp][0]];
    const mg_cgroup_t* gb = &e->scene.cgroups[e->scene.bpairs[pp][1]];
    for (int sa = ga->shape0; sa < ga->shape0 + ga->nshape; sa++)
      for (int sb = gb->shape0; sb < gb->shape0 + gb->nshape; sb++) collide_shapes(e, sa, sb);
  }
  /* cpSpaceArbiterSetFilter: age out cached arbiters */
  {
    int w = 0;
    for (int i = 0; i < e->n_cached; i++) {
      mgo_arbiter* arb = &e->cached[i];
      int ticks = e->stamp - arb->stamp;
      if (ticks >= 1 && arb->state != MGO_ARB_CACHED) arb->state = MGO_ARB_CACHED;
      int keep = ticks < PERSISTENCE;
      if (keep) {
        if (w != i) {
          e->cached[w] = *arb;
          for (int k = 0; k < e->n_active; k++) if (e->active[k] == i) e->active[k] = w;
        }
        w++;
      }
    }
    e->n_cached = w;
  }
  /* prestep */
  double slop = COLLISION_SLOP;
  double biasCoef = 1.0 - pow(e->collision_bias, dt);
  for (int i = 0; i < e->n_active; i++) arbiter_prestep(e, &e->cached[e->active[i]], dt, slop, biasCoef);
  for (int i = 0; i < e->n_joints; i++) joint_prestep(e, &e->joints[i], dt);
  /* integrate velocities: damping 1, gravity 0, no forces (kinematic bodies skipped) */
  for (int i = 0; i < e->n_bodies; i++) {
    mgo_body* b = &e->bodies[i];
    if (b->kind == MG_BODY_KINEMATIC) continue;
    double damping = pow(1.0, dt);
    b->v = vadd(vmult(b->v, damping), vmult(vadd(V(0, 0), vmult(V(0, 0), b->m_inv)), dt));
    b->w = b->w * damping + 0.0 * b->i_inv * dt;
  }
  /* warm start */
  double dt_coef = (prev_dt == 0.0 ? 0.0 : dt / prev_dt);
  for (int i = 0; i < e->n_active; i++) arbiter_apply_cached(e, &e->cached[e->active[i]], dt_coef);
  for (int i = 0; i < e->n_joints; i++) joint_apply_cached(e, &e->joints[i], dt_coef);
  /* solver */
  for (int it = 0; it < ITERATIONS; it++) {
    for (int i = 0; i < e->n_active; i++) arbiter_apply_impulse(e, &e->cached[e->active[i]]);
    for (int i = 0; i < e->n_joints; i++) joint_apply_impulse(e, &e->joints[i], dt);
  }
}

/* --------------------------------------------- Robot.set_action / update */
void mgo_set_action(mgo_env* e, int action) {
  /* id = 9*grip + 3*lr + ud (entities.py:162-186) */
  int ud = action % 3, lr = (action / 3) % 3, grip = action / 9;
  double R = e->scene.robot_radius;
  e->rel_turn_angle = 0.0;
  e->target_speed = 0.0;
  if (ud == 1) e->target_speed += 4.0 * R;
  if (ud == 2) e->target_speed -= 3.0 * R;
  if (lr == 1) e->rel_turn_angle += 1.5;
  if (lr == 2) e->rel_turn_angle -= 1.5;
  if (grip == 0) e->target_finger_angle = M_PI / 8; /* OPEN: finger_rot_limit_outer */
  else e->target_finger_angle = -0.0;               /* CLOSE: -finger_rot_limit_inner */
}

void mgo_robot_update(mgo_env* e) {
  const mg_scene_t* s = &e->scene;
  mgo_body* robot = &e->bodies[s->robot_body];
  mgo_body* control = &e->bodies[s->control_body];
  /* control_body.angle = ... goes through cpBodySetAngle: refreshes the rotation */
  control->a = robot->a + e->rel_turn_angle;
  rot_for_angle(e, control->a, &control->rot);
  /* rotation_vector.cpvrotate((0, speed)) */
  v2 xv = V(0.0, e->target_speed);
  control->v = V(robot->rot.x * xv.x - robot->rot.y * xv.y, robot->rot.x * xv.y + robot->rot.y * xv.x);
  for (int f = 0; f < 2; f++) {
    double side = f == 0 ? -1.0 : 1.0;
    double rel_angle = e->bodies[s->finger_body[f]].a - robot->a;
    double angle_error = rel_angle + side * e->target_finger_angle;
    double target_rate = fmax_(-1, fmin_(1, angle_error * 10));
    if (fabs(target_rate) < 1e-4) target_rate = 0.0;
    e->joints[s->motor_joint[f]].rate = target_rate;
  }
}

/* ------------------------------------------------------------- lifecycle */
void mgo_reset(mgo_env* e) {
  const mg_scene_t* s = &e->scene;
  e->n_bodies = s->n_bodies;
  e->n_shapes = s->n_shapes;
  e->n_joints = s->n_joints;
  e->static_body = STATIC_BODY_TEMPLATE;
  e->static_body.rot = V(1.0, 0.0);
  for (int i = 0; i < s->n_bodies; i++) {
    mgo_body* b = &e->bodies[i];
    memset(b, 0, sizeof(*b));
    b->m_inv = s->bodies[i].m_inv;
    b->i_inv = s->bodies[i].i_inv;
    b->kind = s->bodies[i].kind;
    b->p = V(s->bodies[i].p0[0], s->bodies[i].p0[1]);
    b->a = s->bodies[i].a0;
    rot_for_angle(e, b->a, &b->rot);
  }
  for (int i = 0; i < s->n_shapes; i++) {
    mgo_shape* sh = &e->shapes[i];
    const mg_shape_t* src = &s->shapes[i];
    memset(sh, 0, sizeof(*sh));
    sh->kind = src->kind; sh->body = src->body; sh->nvert = src->nvert;
    sh->radius = src->radius; sh->friction = src->friction; sh->group = src->group;
    for (int k = 0; k < src->nvert; k++) sh->lv[k] = V(s->cverts[src->vert0 + k][0], s->cverts[src->vert0 + k][1]);
    if (sh->kind == MG_SHAPE_SEGMENT) {
      sh->ln[0] = vrperp(vnormalize(vsub(sh->lv[1], sh->lv[0])));
    } else if (sh->kind == MG_SHAPE_POLY) {
      /* cpPolyShape SetVerts: plane i = edge (v[i-1] -> v[i]), outward normal */
      for (int k = 0; k < sh->nvert; k++) {
        v2 a = sh->lv[(k - 1 + sh->nvert) % sh->nvert];
        v2 b = sh->lv[k];
        sh->ln[k] = vnormalize(vrperp(vsub(b, a)));
      }
    }
    shape_cache(e, sh);
  }
  for (int i = 0; i < s->n_joints; i++) {
    mgo_joint* j = &e->joints[i];
    const mg_joint_t* src = &s->joints[i];
    memset(j, 0, sizeof(*j));
    j->kind = src->kind; j->a = src->a; j->b = src->b;
    j->anchor_a = V(src->anchor_a[0], src->anchor_a[1]);
    j->anchor_b = V(src->anchor_b[0], src->anchor_b[1]);
    j->p0 = src->p0; j->p1 = src->p1; j->p2 = src->p2;
    j->max_force = src->max_force; j->max_bias = src->max_bias; j->error_bias = src->error_bias;
  }
  e->n_cached = 0;
  e->n_active = 0;
  e->stamp = 0;
  e->curr_dt = 0.0;
  e->episode_steps = 0;
  e->overflow = 0;
  e->rel_turn_angle = 0.0;
  e->target_speed = 0.0;
  e->target_finger_angle = 0.0;
}

mgo_env* mgo_create(const mg_scene_t* scene) {
  mgo_env* e = (mgo_env*)calloc(1, sizeof(mgo_env));
  if (!e) return NULL;
  e->scene = *scene;
  e->collision_bias = pow(1.0 - 0.1, 60.0); /* cpSpace default */
  e->det_sincos = 0;
  e->pair_perm = NULL;
  mgo_reset(e);
  return e;
}
void mgo_destroy(mgo_env* e) {
  if (e) { free(e->pair_perm); free(e); }
}
void mgo_set_det_sincos(mgo_env* e, int on) { e->det_sincos = on; mgo_reset(e); }
void mgo_set_pair_permutation(mgo_env* e, const int32_t* perm) {
  free(e->pair_perm);
  e->pair_perm = NULL;
  if (perm) {
    e->pair_perm = (int32_t*)malloc(sizeof(int32_t) * e->scene.n_bpairs);
    memcpy(e->pair_perm, perm, sizeof(int32_t) * e->scene.n_bpairs);
  }
}

void mgo_phys_steps_on_frame(mgo_env* e) {
  /* base_env.py:236-243 */
  for (int i = 0; i < 10; i++) {
    mgo_robot_update(e);
    mgo_space_step(e, DT);
  }
}

void mgo_step(mgo_env* e, int action, float* reward, uint8_t* done, float* score) {
  /* BaseEnv.step without the render (base_env.py:255-288) */
  mgo_set_action(e, action);
  mgo_phys_steps_on_frame(e);
  e->episode_steps++;
  int d = e->scene.max_steps > 0 && e->episode_steps >= e->scene.max_steps;
  double sc = 0.0;
  if (d) sc = mgo_score(e);
  double rew = 0.0;
  if (e->scene.debug_reward) rew = mgo_debug_reward(e);
  if (reward) *reward = (float)rew;
  if (done) *done = (uint8_t)d;
  if (score) *score = (float)sc;
}

void mgo_set_pose(mgo_env* e, int body, double x, double y, double angle) {
  mgo_body* b = &e->bodies[body];
  b->p = V(x, y);
  b->a = angle;
  rot_for_angle(e, b->a, &b->rot);
  for (int i = 0; i < e->n_shapes; i++) if (e->shapes[i].body == body) shape_cache(e, &e->shapes[i]);
}

void mgo_get_state(const mgo_env* e, mg_state_t* out) {
  memset(out, 0, sizeof(*out));
  out->n_bodies = e->n_bodies;
  out->n_joints = e->n_joints;
  out->episode_steps = e->episode_steps;
  out->overflow = e->overflow;
  for (int i = 0; i < e->n_bodies; i++) {
    out->pos[i][0] = e->bodies[i].p.x; out->pos[i][1] = e->bodies[i].p.y;
    out->angle[i] = e->bodies[i].a;
    out->vel[i][0] = e->bodies[i].v.x; out->vel[i][1] = e->bodies[i].v.y;
    out->angvel[i] = e->bodies[i].w;
  }
  for (int i = 0; i < e->n_joints; i++) {
    out->joint_acc[i][0] = e->joints[i].jAcc.x;
    out->joint_acc[i][1] = e->joints[i].jAcc.y;
  }
  int nc = 0;
  for (int i = 0; i < e->n_active; i++) {
    const mgo_arbiter* arb = &e->cached[e->active[i]];
    for (int k = 0; k < arb->count && nc < 32; k++, nc++) {
      out->contact_shapes[nc][0] = arb->a;
      out->contact_shapes[nc][1] = arb->b;
      out->contact_jn[nc] = arb->contacts[k].jnAcc;
      out->contact_jt[nc] = arb->contacts[k].jtAcc;
    }
  }
  out->n_contacts = nc;
  /* the complete carry-over state (mg_state_t v2): bias velocities and the arbiter cache, this sub-step's
   * arbiters first (collision order), then the older ones */
  out->stamp = e->stamp;
  for (int i = 0; i < e->n_bodies; i++) {
    out->bias_vel[i][0] = e->bodies[i].v_bias.x; out->bias_vel[i][1] = e->bodies[i].v_bias.y;
    out->bias_angvel[i] = e->bodies[i].w_bias;
  }
  int ne = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int n = pass == 0 ? e->n_active : e->n_cached;
    for (int i = 0; i < n; i++) {
      const mgo_arbiter* arb = &e->cached[pass == 0 ? e->active[i] : i];
      if (pass == 1 && arb->stamp == e->stamp) continue; /* already listed */
      for (int k = 0; k < arb->count && ne < MG_STATE_CACHE; k++, ne++) {
        out->cache_shapes[ne][0] = arb->a; out->cache_shapes[ne][1] = arb->b;
        out->cache_hash[ne] = arb->contacts[k].hash;
        out->cache_age[ne] = e->stamp - arb->stamp;
        out->cache_jn[ne] = arb->contacts[k].jnAcc; out->cache_jt[ne] = arb->contacts[k].jtAcc;
      }
    }
  }
  out->n_cache = ne;
}

/* restore a snapshot (mg_set_state's counterpart): entries of one shape pair form one arbiter; an arbiter that
 * collided in the last sub-step (age 0) is NORMAL, older ones are CACHED (what cpSpaceStep leaves behind) */
void mgo_set_state(mgo_env* e, const mg_state_t* in) {
  for (int i = 0; i < e->n_bodies; i++) {
    mgo_body* b = &e->bodies[i];
    b->p = V(in->pos[i][0], in->pos[i][1]);
    b->a = in->angle[i];
    rot_for_angle(e, b->a, &b->rot);
    b->v = V(in->vel[i][0], in->vel[i][1]);
    b->w = in->angvel[i];
    b->v_bias = V(in->bias_vel[i][0], in->bias_vel[i][1]);
    b->w_bias = in->bias_angvel[i];
  }
  for (int i = 0; i < e->n_shapes; i++) if (e->shapes[i].body >= 0) shape_cache(e, &e->shapes[i]);
  for (int i = 0; i < e->n_joints; i++) e->joints[i].jAcc = V(in->joint_acc[i][0], in->joint_acc[i][1]);
  e->stamp = in->stamp;
  /* cpSpace.curr_dt: the warm start scales cached impulses by dt / prev_dt (0 before the first step) */
  e->curr_dt = in->stamp > 0 ? DT : 0.0;
  e->episode_steps = in->episode_steps;
  e->overflow = in->overflow;
  e->n_cached = 0;
  e->n_active = 0;
  for (int k = 0; k < in->n_cache; k++) {
    mgo_arbiter* arb = arbiter_find(e, in->cache_shapes[k][0], in->cache_shapes[k][1], 1);
    if (!arb || arb->count >= 2) continue;
    mgo_contact* con = &arb->contacts[arb->count++];
    memset(con, 0, sizeof(*con));
    con->hash = in->cache_hash[k];
    con->jnAcc = in->cache_jn[k]; con->jtAcc = in->cache_jt[k];
    arb->stamp = in->stamp - in->cache_age[k];
    arb->state = in->cache_age[k] == 0 ? MGO_ARB_NORMAL : MGO_ARB_CACHED;
    arb->body_a = e->shapes[arb->a].body;
    arb->body_b = e->shapes[arb->b].body;
    /* arbiters of the last sub-step are the active list (collision order = snapshot order) */
    if (in->cache_age[k] == 0 && arb->count == 1 && e->n_active < MGO_MAX_ARBITERS)
      e->active[e->n_active++] = (int)(arb - e->cached);
  }
}
