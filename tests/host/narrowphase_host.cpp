// Host build of the PRODUCT's per-lane narrowphase (magical_b200/csrc/mg_narrowphase.h)
// so CPU tests can compare it, bit for bit, with the oracle's recursive restatement and with
// brute-force geometry.  Test infrastructure only.
#include <string.h>

#include "../../magical_b200/csrc/mg_narrowphase.h"
#include "../../magical_b200/csrc/mg_scene_aux.h"
#include "../../magical_b200/csrc/mg_sincos.h"

static ShapeView view_for(const mg_scene_t* sc, const mg_scene_aux_t* aux, const double* poses, int si) {
  const mg_shape_t& sh = sc->shapes[si];
  ShapeView v;
  v.kind = sh.kind; v.nvert = sh.nvert;
  v.lv = &sc->cverts[sh.vert0][0];
  v.ln = &aux->cnorm[sh.vert0][0];
  v.radius = sh.radius; v.index = si;
  if (sh.body >= 0) {
    double s, c;
    mg_det_sincos(poses[3 * sh.body + 2], &s, &c);
    v.rc = c; v.rs = s; v.px = poses[3 * sh.body]; v.py = poses[3 * sh.body + 1];
  } else {
    v.rc = 1.0; v.rs = 0.0; v.px = 0.0; v.py = 0.0;
  }
  return v;
}

extern "C" {

// out: [0]=count [1..2]=n [3..6]=p1[0],p2[0] ... ; hashes in out_hash; returns ordered (a,b) in out_ab
int mgh_collide(const mg_scene_t* sc, const double* poses /* [n_bodies][3] x,y,angle */, int sa, int sb,
                double* out, unsigned* out_hash, int* out_ab) {
  static mg_scene_aux_t aux;  // single-threaded test helper
  const char* why = mg_build_scene_aux(sc, &aux);
  if (why) return -1;
  int ia = sa, ib = sb;
  if (sc->shapes[ia].kind > sc->shapes[ib].kind) { int t = ia; ia = ib; ib = t; }
  ShapeView va = view_for(sc, &aux, poses, ia), vb = view_for(sc, &aux, poses, ib);
  double bba[4], bbb[4];
  sv_bb(va, bba);
  sv_bb(vb, bbb);
  Manifold m;
  m.count = 0;
  m.n = D2(0, 0);
  out_ab[0] = ia; out_ab[1] = ib;
  out[11] = bb_intersects(bba, bbb) ? 1.0 : 0.0;
  mg_collide(va, vb, bba, bbb, m);
  out[0] = m.count; out[1] = m.n.x; out[2] = m.n.y;
  for (int i = 0; i < m.count; i++) {
    out[3 + 4 * i] = m.p1[i].x; out[4 + 4 * i] = m.p1[i].y; out[5 + 4 * i] = m.p2[i].x; out[6 + 4 * i] = m.p2[i].y;
    out_hash[i] = m.hash[i];
  }
  return 0;
}

// signed distance / closest points from the product's GJK+EPA (for brute-force validation)
int mgh_gjk(const mg_scene_t* sc, const double* poses, int sa, int sb, double* out /* d, nx, ny, ax, ay, bx, by */) {
  static mg_scene_aux_t aux;
  if (mg_build_scene_aux(sc, &aux)) return -1;
  ShapeView va = view_for(sc, &aux, poses, sa), vb = view_for(sc, &aux, poses, sb);
  double bba[4], bbb[4];
  sv_bb(va, bba);
  sv_bb(vb, bbb);
  ClosestPts p = mg_gjk(va, vb, bba, bbb);
  out[0] = p.d; out[1] = p.n.x; out[2] = p.n.y; out[3] = p.a.x; out[4] = p.a.y; out[5] = p.b.x; out[6] = p.b.y;
  return 0;
}

int mgh_aux(const mg_scene_t* sc, mg_scene_aux_t* aux) { return mg_build_scene_aux(sc, aux) ? -1 : 0; }
long mgh_sizeof_aux(void) { return (long)sizeof(mg_scene_aux_t); }
}
