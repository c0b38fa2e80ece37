/*
 * mg_scene_aux.h — host-side derivation of per-scene constants the kernels
 * need but the compiled scene does not carry: plane normals, static-shape
 * bounding boxes, per-joint solver constants and the lane/level schedule that
 * lets independent joints run on different lanes while keeping the reference's
 * sequential (Gauss-Seidel) update order bit for bit.
 *
 * Reference semantics being prepared for: Chipmunk2D's constraint preStep /
 * applyImpulse loops inside cpSpaceStep, i.e. `pm.Space.step` at
 * magical/base_env.py:243 with `space.iterations = 10` (base_env.py:196).
 */
#ifndef MG_SCENE_AUX_H
#define MG_SCENE_AUX_H

#include <math.h>
#include <string.h>

#include "../../include/magical_b200.h"

#define MG_DT (1.0 / 8.0 / 10.0) /* base_env.py:238-239 at fps=8 (benchmarks/__init__.py:402) */
#define MG_SUBSTEPS 10           /* base_env.py:237 */
#define MG_ITERATIONS 10         /* benchmarks/__init__.py:404 */
#define MG_COLLISION_SLOP 0.01   /* base_env.py:195 */
#define MG_MAX_LEVELS 32
#define TPE_NO_SLOT_ 15

typedef struct {
  double cnorm[MG_MAX_CVERTS][2]; /* poly: plane i = edge (v[i-1] -> v[i]); segment: normal at vert0 */
  float static_bb[MG_MAX_SHAPES][4]; /* conservative fp32 l,b,r,t for shapes on the static body */
  /* joint solver constants */
  double j_isum[MG_MAX_JOINTS];   /* gear/spring/limit/motor: 1/(sum of inverse moments); pivot: k diagonal */
  double j_wcoef[MG_MAX_JOINTS];  /* spring: 1 - exp(-damping*dt/iSum) */
  double j_bcoef[MG_MAX_JOINTS];  /* bias_coef(error_bias, dt) */
  double j_jmax[MG_MAX_JOINTS];   /* max_force * dt */
  /* schedule: lane l walks sched[0..n_levels)[l] (its component's joints in insertion order, 255 = none) */
  int32_t n_levels;
  int32_t n_springs;
  uint8_t sched[MG_MAX_LEVELS][32]; /* joint index or 255 */
  uint8_t springs[MG_MAX_JOINTS];   /* spring joints in insertion order (their preStep is sequential) */
  double contact_bias_coef;         /* 1 - pow(collision_bias, dt) */
  double body_reach[MG_MAX_BODIES]; /* max distance of any point of the body's shapes from its origin */
  /* per-joint solver constants packed for 128-bit loads: ma, ia, mb, ib (inverse mass / moment of the two
   * bodies, 0 for static and kinematic ones), c0 (pivot k diagonal or iSum), c1 (max_force*dt),
   * c2 (spring w_coef | gear ratio), c3 (gear 1/ratio) */
  double jc[MG_MAX_JOINTS][8];
  uint8_t jkind[MG_MAX_JOINTS], ja[MG_MAX_JOINTS], jb[MG_MAX_JOINTS]; /* body slots; 16 = the static body */
  uint8_t jpin[MG_MAX_JOINTS];  /* pin joints: slot of their per-sub-step frame (r1, r2, n, nMass, bias) */
  uint32_t jpack[MG_MAX_JOINTS]; /* kind | a << 8 | b << 16 | pin slot << 24 */
  int32_t max_per_level;        /* lanes the joint schedule needs (must fit the lanes that serve one environment) */
  int32_t pad2_;
  /* thread-per-environment kernel (mg_physics_tpe.h): the scene must have MAGICAL's canonical structure --
   * one robot whose ten joints appear consecutively in the order of entities.py:255-354, and blocks that each
   * carry exactly a pivot + gear drag pair against the static body (entities.py:703-711) */
  int32_t tpe_ok;
  int32_t tpe_nslots;           /* velocity slots: bodies owning collision shapes + the static dummy (last) */
  int32_t tpe_nblocks;
  int32_t tpe_jr0;              /* first robot joint: jr0+0 pivot, +1 gear, +2,+3 springs, +4..6 / +7..9 pin, limit, motor */
  uint64_t tpe_slotmap;         /* nibble b = slot of body b, 15 = the body has no slot (control, eyes) */
  uint8_t tpe_slot_body[16];    /* slot -> body */
  uint8_t tpe_bj_pivot[MG_MAX_BLOCKS], tpe_bj_gear[MG_MAX_BLOCKS], tpe_bj_slot[MG_MAX_BLOCKS];
  uint8_t tpe_pad_[2];
  int32_t ok;                       /* 0 if the scene uses a feature the kernels do not implement */
  int32_t pad_;
} mg_scene_aux_t;

static inline void mg_aux_norm(double ax, double ay, double bx, double by, double* nx, double* ny) {
  /* cpvnormalize(cpvrperp(b - a)) */
  double dx = bx - ax, dy = by - ay;
  double rx = dy, ry = -dx;
  double inv = 1.0 / (sqrt(rx * rx + ry * ry) + 2.2250738585072014e-308);
  *nx = rx * inv;
  *ny = ry * inv;
}

/* Recognise the canonical MAGICAL structure for the thread-per-environment kernel; sets aux->tpe_ok. */
static inline void mg_build_tpe_aux(const mg_scene_t* s, mg_scene_aux_t* aux) {
  aux->tpe_ok = 0;
  const int robot = s->robot_body, control = s->control_body;
  if (robot < 0 || robot >= s->n_bodies || control < 0 || control >= s->n_bodies) return;
  if (s->bodies[control].kind != MG_BODY_KINEMATIC) return;
  /* bodies that own collision shapes */
  int shaped[MG_MAX_BODIES];
  for (int b = 0; b < MG_MAX_BODIES; b++) shaped[b] = 0;
  for (int i = 0; i < s->n_shapes; i++)
    if (s->shapes[i].body >= 0) shaped[s->shapes[i].body] = 1;
  for (int g = 0; g < s->n_cgroups; g++) { /* a group's shapes all sit on the group's body */
    for (int k = 0; k < s->cgroups[g].nshape; k++)
      if (s->shapes[s->cgroups[g].shape0 + k].body != s->cgroups[g].body) return;
  }
  if (shaped[control] || shaped[s->eye_body[0]] || shaped[s->eye_body[1]]) return;
  if (!shaped[robot] || !shaped[s->finger_body[0]] || !shaped[s->finger_body[1]]) return;
  /* robot chain */
  int jr0 = -1;
  for (int j = 0; j < s->n_joints; j++)
    if (s->joints[j].kind == MG_JOINT_PIVOT && s->joints[j].a == control && s->joints[j].b == robot) { jr0 = j; break; }
  if (jr0 < 0 || jr0 + 10 > s->n_joints) return;
  const int kinds[10] = {MG_JOINT_PIVOT, MG_JOINT_GEAR, MG_JOINT_ROTARY_SPRING, MG_JOINT_ROTARY_SPRING, MG_JOINT_PIN,
                         MG_JOINT_ROTARY_LIMIT, MG_JOINT_MOTOR, MG_JOINT_PIN, MG_JOINT_ROTARY_LIMIT, MG_JOINT_MOTOR};
  const int ja[10] = {control, control, robot, robot, robot, robot, robot, robot, robot, robot};
  const int jb[10] = {robot, robot, s->eye_body[0], s->eye_body[1], s->finger_body[0], s->finger_body[0],
                      s->finger_body[0], s->finger_body[1], s->finger_body[1], s->finger_body[1]};
  for (int k = 0; k < 10; k++) {
    const mg_joint_t* jt = &s->joints[jr0 + k];
    if (jt->kind != kinds[k] || jt->a != ja[k] || jt->b != jb[k]) return;
  }
  if (s->motor_joint[0] != jr0 + 6 || s->motor_joint[1] != jr0 + 9) return;
  if (s->joints[jr0 + 1].p1 == 0.0) return;
  /* every other joint: (pivot, gear) pairs static -> body */
  int nblk = 0;
  int used_body[MG_MAX_BODIES];
  for (int b = 0; b < MG_MAX_BODIES; b++) used_body[b] = 0;
  for (int j = 0; j < s->n_joints;) {
    if (j == jr0) { j += 10; continue; }
    if (j + 1 >= s->n_joints) return;
    const mg_joint_t *p = &s->joints[j], *g = &s->joints[j + 1];
    if (p->kind != MG_JOINT_PIVOT || g->kind != MG_JOINT_GEAR || p->a >= 0 || g->a >= 0 || p->b != g->b) return;
    if (p->b < 0 || p->b >= s->n_bodies || !shaped[p->b] || used_body[p->b] || nblk >= MG_MAX_BLOCKS) return;
    if (p->b == robot || p->b == s->finger_body[0] || p->b == s->finger_body[1]) return;
    if (s->bodies[p->b].kind != MG_BODY_DYNAMIC || g->p1 == 0.0) return;
    if (g->max_bias != 0.0) return; /* entities.py:709: pure angular drag, its bias is exactly zero */
    used_body[p->b] = 1;
    aux->tpe_bj_pivot[nblk] = (uint8_t)j;
    aux->tpe_bj_gear[nblk] = (uint8_t)(j + 1);
    aux->tpe_bj_slot[nblk] = (uint8_t)p->b; /* body for now, mapped to its slot below */
    nblk++;
    j += 2;
  }
  /* slots: shaped bodies in body order; every shaped body must be the robot, a finger or a jointed block */
  int nslots = 0;
  uint64_t map = 0;
  for (int b = 0; b < MG_MAX_BODIES; b++) {
    int slot = TPE_NO_SLOT_;
    if (b < s->n_bodies && shaped[b]) {
      if (!(b == robot || b == s->finger_body[0] || b == s->finger_body[1] || used_body[b])) return;
      if (s->bodies[b].kind != MG_BODY_DYNAMIC) return;
      slot = nslots;
      aux->tpe_slot_body[nslots++] = (uint8_t)b;
    }
    map |= (uint64_t)slot << (4 * b);
  }
  if (nslots + 1 > 15) return;
  for (int k = 0; k < nblk; k++) aux->tpe_bj_slot[k] = (uint8_t)((map >> (4 * aux->tpe_bj_slot[k])) & 15u);
  aux->tpe_slotmap = map;
  aux->tpe_nslots = nslots + 1;
  aux->tpe_nblocks = nblk;
  aux->tpe_jr0 = jr0;
  aux->tpe_ok = 1;
}

static inline const char* mg_build_scene_aux(const mg_scene_t* s, mg_scene_aux_t* aux) {
  memset(aux, 0, sizeof(*aux));
  const double dt = MG_DT;
  if (s->n_bodies < 1 || s->n_bodies > MG_MAX_BODIES) return "bad n_bodies";
  if (s->n_shapes < 0 || s->n_shapes > MG_MAX_SHAPES) return "bad n_shapes";
  if (s->n_joints < 0 || s->n_joints > MG_MAX_JOINTS) return "bad n_joints";
  if (s->n_cgroups < 0 || s->n_cgroups > MG_MAX_CGROUPS) return "bad n_cgroups";
  if (s->n_bpairs < 0 || s->n_bpairs > MG_MAX_BPAIRS) return "bad n_bpairs";
  if (s->n_prims < 0 || s->n_prims > MG_MAX_PRIMS) return "bad n_prims";
  if (s->n_blocks < 0 || s->n_blocks > MG_MAX_BLOCKS) return "bad n_blocks";
  if (s->n_goals < 0 || s->n_goals > MG_MAX_GOALS) return "bad n_goals";
  for (int i = 0; i < s->n_shapes; i++) {
    const mg_shape_t* sh = &s->shapes[i];
    if (sh->vert0 < 0 || sh->vert0 + sh->nvert > MG_MAX_CVERTS) return "shape vertex range";
    if (sh->body >= s->n_bodies) return "shape body index";
    const double(*v)[2] = &s->cverts[sh->vert0];
    if (sh->kind == MG_SHAPE_SEGMENT) {
      /* cpSegmentShape: n = rperp(normalize(b - a)) */
      double dx = v[1][0] - v[0][0], dy = v[1][1] - v[0][1];
      double inv = 1.0 / (sqrt(dx * dx + dy * dy) + 2.2250738585072014e-308);
      double ux = dx * inv, uy = dy * inv;
      aux->cnorm[sh->vert0][0] = uy;
      aux->cnorm[sh->vert0][1] = -ux;
    } else if (sh->kind == MG_SHAPE_POLY) {
      if (sh->nvert < 3 || sh->nvert > 16) return "poly vertex count";
      for (int k = 0; k < sh->nvert; k++) {
        int p = (k - 1 + sh->nvert) % sh->nvert;
        mg_aux_norm(v[p][0], v[p][1], v[k][0], v[k][1], &aux->cnorm[sh->vert0 + k][0], &aux->cnorm[sh->vert0 + k][1]);
      }
    } else if (sh->kind != MG_SHAPE_CIRCLE) {
      return "unknown shape kind";
    }
    if (sh->body < 0) {
      double l = 1e300, r = -1e300, b = 1e300, t = -1e300;
      for (int k = 0; k < sh->nvert; k++) {
        if (v[k][0] < l) l = v[k][0];
        if (v[k][0] > r) r = v[k][0];
        if (v[k][1] < b) b = v[k][1];
        if (v[k][1] > t) t = v[k][1];
      }
      aux->static_bb[i][0] = nextafterf((float)(l - sh->radius), -INFINITY);
      aux->static_bb[i][1] = nextafterf((float)(b - sh->radius), -INFINITY);
      aux->static_bb[i][2] = nextafterf((float)(r + sh->radius), INFINITY);
      aux->static_bb[i][3] = nextafterf((float)(t + sh->radius), INFINITY);
    }
  }
  /* joints */
  int jcomp_a[MG_MAX_JOINTS], jcomp_b[MG_MAX_JOINTS];
  int per_level[MG_MAX_LEVELS];
  for (int i = 0; i < MG_MAX_LEVELS; i++) per_level[i] = 0;
  memset(aux->sched, 255, sizeof(aux->sched));
  for (int j = 0; j < s->n_joints; j++) {
    const mg_joint_t* jt = &s->joints[j];
    if (jt->a >= s->n_bodies || jt->b >= s->n_bodies || jt->b < 0) return "joint body index";
    double ia = jt->a < 0 ? 0.0 : s->bodies[jt->a].i_inv;
    double ib = s->bodies[jt->b].i_inv;
    double ma = jt->a < 0 ? 0.0 : s->bodies[jt->a].m_inv;
    double mb = s->bodies[jt->b].m_inv;
    aux->j_bcoef[j] = 1.0 - pow(jt->error_bias, dt);
    aux->j_jmax[j] = jt->max_force * dt;
    switch (jt->kind) {
      case MG_JOINT_PIVOT: {
        /* the kernels implement the centre-to-centre pivot only (both anchors at the body
         * origins), which is the only form the reference creates (entities.py:255, 703) */
        if (jt->anchor_a[0] != 0.0 || jt->anchor_a[1] != 0.0 || jt->anchor_b[0] != 0.0 || jt->anchor_b[1] != 0.0)
          return "pivot joints with non-zero anchors are not implemented";
        if (jt->max_bias != 0.0) return "pivot joints with positional correction are not implemented";
        double m_sum = ma + mb;
        double det = m_sum * m_sum - 0.0 * 0.0;
        double det_inv = 1.0 / det;
        aux->j_isum[j] = m_sum * det_inv; /* k_tensor diagonal for r1 = r2 = 0 */
      } break;
      case MG_JOINT_GEAR: {
        double ratio = jt->p1, ratio_inv = 1.0 / jt->p1;
        aux->j_isum[j] = 1.0 / (ia * ratio_inv + ratio * ib);
      } break;
      case MG_JOINT_ROTARY_SPRING: {
        double moment = ia + ib;
        aux->j_isum[j] = 1.0 / moment;
        aux->j_wcoef[j] = 1.0 - exp(-jt->p2 * dt * moment);
        aux->springs[aux->n_springs++] = (uint8_t)j;
      } break;
      case MG_JOINT_PIN:
        break;
      case MG_JOINT_ROTARY_LIMIT:
      case MG_JOINT_MOTOR:
        aux->j_isum[j] = 1.0 / (ia + ib);
        break;
      default:
        return "unknown joint kind";
    }
    /* Chains instead of levels: joints are grouped by the connected component (over dynamic
     * bodies) they belong to; one lane walks a component's joints in insertion order, and components
     * share no dynamic body, so no barrier is needed between joints. */
    jcomp_a[j] = (jt->a >= 0 && s->bodies[jt->a].kind == MG_BODY_DYNAMIC) ? jt->a : -1;
    jcomp_b[j] = (s->bodies[jt->b].kind == MG_BODY_DYNAMIC) ? jt->b : -1;
  }
  {
    int parent[MG_MAX_BODIES];
    for (int i = 0; i < MG_MAX_BODIES; i++) parent[i] = i;
    for (int j = 0; j < s->n_joints; j++) {
      if (jcomp_a[j] >= 0 && jcomp_b[j] >= 0) {
        int x = jcomp_a[j], y = jcomp_b[j];
        while (parent[x] != x) x = parent[x];
        while (parent[y] != y) y = parent[y];
        if (x != y) parent[y < x ? x : y] = (y < x ? y : x);
      }
    }
    int comp_lane[MG_MAX_BODIES];
    for (int i = 0; i < MG_MAX_BODIES; i++) comp_lane[i] = -1;
    int n_lanes = 0;
    for (int j = 0; j < s->n_joints; j++) {
      int body = jcomp_b[j] >= 0 ? jcomp_b[j] : jcomp_a[j];
      if (body < 0) return "joint between two non-dynamic bodies";
      while (parent[body] != body) body = parent[body];
      if (comp_lane[body] < 0) {
        if (n_lanes >= 32) return "too many joint components";
        comp_lane[body] = n_lanes++;
      }
      int lane = comp_lane[body];
      int pos = per_level[lane]++; /* per_level[] doubles as the chain length per lane */
      if (pos >= MG_MAX_LEVELS) return "joint chain too deep";
      aux->sched[pos][lane] = (uint8_t)j;
      if (pos + 1 > aux->n_levels) aux->n_levels = pos + 1;
    }
    aux->max_per_level = n_lanes;
  }
  int n_pins = 0;
  for (int j = 0; j < s->n_joints; j++) {
    const mg_joint_t* jt = &s->joints[j];
    aux->jpin[j] = 255;
    if (jt->kind == MG_JOINT_PIN) {
      if (n_pins >= 4) return "more than 4 pin joints";
      aux->jpin[j] = (uint8_t)n_pins++;
    }
    aux->jkind[j] = (uint8_t)jt->kind;
    aux->ja[j] = (uint8_t)(jt->a < 0 ? MG_MAX_BODIES : jt->a);
    aux->jb[j] = (uint8_t)jt->b;
    aux->jpack[j] = (uint32_t)aux->jkind[j] | ((uint32_t)aux->ja[j] << 8) | ((uint32_t)aux->jb[j] << 16) | ((uint32_t)aux->jpin[j] << 24);
    aux->jc[j][0] = jt->a < 0 ? 0.0 : s->bodies[jt->a].m_inv;
    aux->jc[j][1] = jt->a < 0 ? 0.0 : s->bodies[jt->a].i_inv;
    aux->jc[j][2] = s->bodies[jt->b].m_inv;
    aux->jc[j][3] = s->bodies[jt->b].i_inv;
    aux->jc[j][4] = aux->j_isum[j];
    aux->jc[j][5] = aux->j_jmax[j];
    aux->jc[j][6] = jt->kind == MG_JOINT_ROTARY_SPRING ? aux->j_wcoef[j] : (jt->kind == MG_JOINT_GEAR ? jt->p1 : 0.0);
    aux->jc[j][7] = jt->kind == MG_JOINT_GEAR ? 1.0 / jt->p1 : 0.0;
  }
  for (int i = 0; i < s->n_shapes; i++) {
    const mg_shape_t* sh = &s->shapes[i];
    if (sh->body < 0) continue;
    for (int k = 0; k < sh->nvert; k++) {
      double d = sqrt(s->cverts[sh->vert0 + k][0] * s->cverts[sh->vert0 + k][0] +
                      s->cverts[sh->vert0 + k][1] * s->cverts[sh->vert0 + k][1]) + sh->radius;
      d = d * (1.0 + 1e-9) + 1e-12; /* strictly above the rounded rotated vertices: the reach is used as a bound */
      if (d > aux->body_reach[sh->body]) aux->body_reach[sh->body] = d;
    }
  }
  for (int g = 0; g < s->n_cgroups; g++) {
    if (s->cgroups[g].shape0 + s->cgroups[g].nshape > s->n_shapes) return "cgroup shape range";
  }
  for (int p = 0; p < s->n_bpairs; p++) {
    if (s->bpairs[p][0] >= s->n_cgroups || s->bpairs[p][1] >= s->n_cgroups) return "bpair index";
  }
  aux->contact_bias_coef = 1.0 - pow(pow(1.0 - 0.1, 60.0), dt);
  aux->ok = 1;
  mg_build_tpe_aux(s, aux);
  return 0;
}

#endif /* MG_SCENE_AUX_H */
