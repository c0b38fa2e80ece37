/*
 * mg_physics.cu — K1: fused 10-sub-step rigid-body physics for one env-step.
 *
 * Replaces, for a whole batch in one launch, the reference's
 *   Robot.set_action(flags)                         entities.py:439-457
 *   for i in range(10):                             base_env.py:236-243
 *       Robot.update(dt)                            entities.py:459-479
 *       pm.Space.step(dt)  -> Chipmunk cpSpaceStep  (third party)
 *
 * B200 mapping: ONE WARP PER ENVIRONMENT.  The environment's state record
 * (bodies, joint accumulators, arbiter cache; 3.4 KB) is streamed from HBM
 * into shared memory with 128-bit loads once, all 10 sub-steps x 10 solver
 * iterations run out of shared memory / registers, and the record is streamed
 * back once.  Lanes map to bodies (integration), shapes (bounding boxes),
 * candidate pairs (broadphase, narrowphase), contacts and joints (prestep,
 * solve).  The sequential-impulse Gauss-Seidel order of the reference is kept
 * BIT FOR BIT: contacts and joints that share no dynamic body commute, so they
 * are grouped into dependency levels and each level runs across lanes; levels
 * run in order with __syncwarp() between them.  Warp ballots / shuffles do the
 * pair -> contact compaction in canonical order.
 *
 * No tensor cores: there is no dense contraction on this path.
 */
#include "mg_device.cuh"
#include "mg_narrowphase.h"
#include "mg_sincos.h"

#define FULL 0xffffffffu

struct __align__(16) ConSmem {
  double r1x, r1y, r2x, r2y, nx, ny, jn, jt;
  double u;
  uint32_t hash;
  uint8_t ba, bb; /* body slots, 16 = static */
  uint8_t arb, slot;
  uint8_t first;
  uint8_t pad_[7];
};

/* Separation cache: a pair whose shapes were measured `margin` apart cannot touch before the two bodies
 * have moved that far.  `limit` is the value of (path[a] + path[b]) at which the measurement expires. */
struct SepEntry {
  uint8_t a, b;
  uint16_t pad_;
  float limit;
};
#define MG_NSEP 32
#define SLOT_STATIC MG_MAX_BODIES /* velocity slot of the static body: always reads as zero */

struct __align__(16) EnvSmem {
  EnvState st; /* staged copy of the HBM record */
  double4 V[MG_MAX_BODIES + 1];  /* working velocities (vx, vy, w); slot 16 = static body */
  double4 Bv[MG_MAX_BODIES + 1]; /* working bias velocities */
  double2 MI[MG_MAX_BODIES + 1]; /* m_inv, i_inv (0 for static / kinematic) */
  double jdyn[MG_MAX_JOINTS][8]; /* per-sub-step joint data (bias / rate / pin frame) */
  float4 sbb[MG_MAX_SHAPES];     /* conservative fp32 shape boxes (l, b, r, t) */
  float4 gbb[MG_MAX_CGROUPS];    /* collision-group boxes */
  ConSmem con[MG_NCON];
  ArbEntry arb2[MG_NARB];
  SepEntry sep[MG_NSEP];
  float path[MG_MAX_BODIES + 1]; /* upper bound of the distance any point of the body has moved this launch */
  uint8_t cand[MG_NCAND][2];
  uint8_t blevel[MG_MAX_BODIES + 1];
  uint8_t clevel[MG_NCON];
  int32_t max_clevel;
  int32_t n_sep;
  int32_t pad_[2];
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

__device__ __forceinline__ bool f4_overlap(float4 a, float4 b) {
  return a.x <= b.z && b.x <= a.z && a.y <= b.w && b.y <= a.w;
}

__device__ __forceinline__ ShapeView make_view(const EnvSmem& S, const DeviceScene* ds, int si) {
  const mg_shape_t& sh = ds->s.shapes[si];
  ShapeView v;
  v.kind = sh.kind;
  v.nvert = sh.nvert;
  v.lv = &ds->s.cverts[sh.vert0][0];
  v.ln = &ds->aux.cnorm[sh.vert0][0];
  v.radius = sh.radius;
  v.index = si;
  int b = sh.body;
  if (b >= 0) {
    double4 P = S.st.P[b];
    double2 R = S.st.R[b];
    v.rc = R.x; v.rs = R.y; v.px = P.x; v.py = P.y;
  } else {
    v.rc = 1.0; v.rs = 0.0; v.px = 0.0; v.py = 0.0;
  }
  return v;
}

/* apply_impulse(body, j, r): v += j*m_inv; w += i_inv * cross(r, j).  Branch-free: static and kinematic
 * bodies carry zero inverse mass, so the update adds exactly zero to them. */
__device__ __forceinline__ void apply_imp(EnvSmem& S, int b, double m_inv, double i_inv, double jx, double jy,
                                          double rx, double ry) {
  double4 t = S.V[b];
  t.x = t.x + jx * m_inv;
  t.y = t.y + jy * m_inv;
  t.z += i_inv * (rx * jy - ry * jx);
  S.V[b] = t;
}
__device__ __forceinline__ void apply_bias_imp(EnvSmem& S, int b, double m_inv, double i_inv, double jx, double jy,
                                               double rx, double ry) {
  double4 t = S.Bv[b];
  t.x = t.x + jx * m_inv;
  t.y = t.y + jy * m_inv;
  t.z += i_inv * (rx * jy - ry * jx);
  S.Bv[b] = t;
}
__device__ __forceinline__ double ld_angle(const EnvSmem& S, int b) { return (b < MG_MAX_BODIES) ? S.st.P[b].z : 0.0; }

/* ------------------------------------------------------------------ joints */
struct JC {
  double ma, ia, mb, ib, c0, c1, c2, c3;
};
__device__ __forceinline__ JC ld_jc(const DeviceScene* ds, int j) {
  const double2* p = reinterpret_cast<const double2*>(&ds->aux.jc[j][0]);
  double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  JC r;
  r.ma = a.x; r.ia = a.y; r.mb = b.x; r.ib = b.y; r.c0 = c.x; r.c1 = c.y; r.c2 = d.x; r.c3 = d.y;
  return r;
}

__device__ __forceinline__ void joint_prestep(EnvSmem& S, const DeviceScene* ds, int j) {
  const mg_joint_t& J = ds->s.joints[j];
  const double dt = MG_DT;
  const int a = ds->aux.ja[j], b = ds->aux.jb[j];
  switch (ds->aux.jkind[j]) {
    case MG_JOINT_GEAR: {
      double maxBias = J.max_bias;
      double ratio = J.p1;
      S.jdyn[j][0] = dclamp(-ds->aux.j_bcoef[j] * (ld_angle(S, b) * ratio - ld_angle(S, a) - J.p0) / dt, -maxBias, maxBias);
    } break;
    case MG_JOINT_PIN: {
      double2 Ra = (a < MG_MAX_BODIES) ? S.st.R[a] : make_double2(1.0, 0.0);
      double2 Rb = S.st.R[b];
      d2 r1 = D2(Ra.x * J.anchor_a[0] - Ra.y * J.anchor_a[1], Ra.y * J.anchor_a[0] + Ra.x * J.anchor_a[1]);
      d2 r2 = D2(Rb.x * J.anchor_b[0] - Rb.y * J.anchor_b[1], Rb.y * J.anchor_b[0] + Rb.x * J.anchor_b[1]);
      d2 pa = (a < MG_MAX_BODIES) ? D2(S.st.P[a].x, S.st.P[a].y) : D2(0, 0);
      d2 pb = D2(S.st.P[b].x, S.st.P[b].y);
      d2 delta = dsub(dadd(pb, r2), dadd(pa, r1));
      double dist = dlength(delta);
      d2 n = dmul(delta, 1.0 / (dist ? dist : MG_INF));
      double2 ma = S.MI[a], mb = S.MI[b];
      double rcn1 = dcross(r1, n), rcn2 = dcross(r2, n);
      double k = (ma.x + ma.y * rcn1 * rcn1) + (mb.x + mb.y * rcn2 * rcn2);
      double maxBias = J.max_bias;
      S.jdyn[j][0] = r1.x; S.jdyn[j][1] = r1.y; S.jdyn[j][2] = r2.x; S.jdyn[j][3] = r2.y;
      S.jdyn[j][4] = n.x; S.jdyn[j][5] = n.y;
      S.jdyn[j][6] = 1.0 / k;
      S.jdyn[j][7] = dclamp(-ds->aux.j_bcoef[j] * (dist - J.p0) / dt, -maxBias, maxBias);
    } break;
    case MG_JOINT_ROTARY_LIMIT: {
      double dist = ld_angle(S, b) - ld_angle(S, a);
      double pdist = 0.0;
      if (dist > J.p1) pdist = J.p1 - dist;
      else if (dist < J.p0) pdist = J.p0 - dist;
      double maxBias = J.max_bias;
      double bias = dclamp(-ds->aux.j_bcoef[j] * pdist / dt, -maxBias, maxBias);
      S.jdyn[j][0] = bias;
      if (!bias) S.st.jacc[j].x = 0.0;
    } break;
    default:
      break; /* pivot, motor: nothing per step; springs are pre-stepped sequentially */
  }
}

__device__ __forceinline__ void spring_prestep(EnvSmem& S, const DeviceScene* ds, int j) {
  const mg_joint_t& J = ds->s.joints[j];
  const int a = ds->aux.ja[j], b = ds->aux.jb[j];
  S.jdyn[j][0] = 0.0; /* target_wrn */
  double j_spring = ((ld_angle(S, a) - ld_angle(S, b)) - J.p0) * J.p1 * MG_DT;
  S.st.jacc[j].x = j_spring;
  S.V[a].z -= j_spring * S.MI[a].y;
  S.V[b].z += j_spring * S.MI[b].y;
}

__device__ __forceinline__ void joint_warm(EnvSmem& S, const DeviceScene* ds, int j) {
  const int a = ds->aux.ja[j], b = ds->aux.jb[j];
  const JC c = ld_jc(ds, j);
  double2 acc = S.st.jacc[j];
  switch (ds->aux.jkind[j]) {
    case MG_JOINT_PIVOT:
      apply_imp(S, a, c.ma, c.ia, -acc.x, -acc.y, 0.0, 0.0);
      apply_imp(S, b, c.mb, c.ib, acc.x, acc.y, 0.0, 0.0);
      break;
    case MG_JOINT_GEAR:
      S.V[a].z -= acc.x * c.ia * c.c3;
      S.V[b].z += acc.x * c.ib;
      break;
    case MG_JOINT_PIN: {
      double jx = S.jdyn[j][4] * acc.x, jy = S.jdyn[j][5] * acc.x;
      apply_imp(S, a, c.ma, c.ia, -jx, -jy, S.jdyn[j][0], S.jdyn[j][1]);
      apply_imp(S, b, c.mb, c.ib, jx, jy, S.jdyn[j][2], S.jdyn[j][3]);
    } break;
    case MG_JOINT_ROTARY_LIMIT:
    case MG_JOINT_MOTOR:
      S.V[a].z -= acc.x * c.ia;
      S.V[b].z += acc.x * c.ib;
      break;
    default:
      break;
  }
}

__device__ __forceinline__ void joint_apply(EnvSmem& S, const DeviceScene* ds, int j) {
  const int a = ds->aux.ja[j], b = ds->aux.jb[j];
  const JC c = ld_jc(ds, j);
  switch (ds->aux.jkind[j]) {
    case MG_JOINT_PIVOT: {
      double4 va = S.V[a], vb = S.V[b];
      double jx = (0.0 - (vb.x - va.x)) * c.c0;
      double jy = (0.0 - (vb.y - va.y)) * c.c0;
      double2 old = S.st.jacc[j];
      d2 acc = dvclamp(D2(old.x + jx, old.y + jy), c.c1);
      S.st.jacc[j] = make_double2(acc.x, acc.y);
      jx = acc.x - old.x; jy = acc.y - old.y;
      /* anchors are at the body origins: no angular part (adds exactly zero) */
      va.x = va.x + (-jx) * c.ma; va.y = va.y + (-jy) * c.ma;
      vb.x = vb.x + jx * c.mb; vb.y = vb.y + jy * c.mb;
      S.V[a] = va;
      S.V[b] = vb;
    } break;
    case MG_JOINT_GEAR: {
      double wr = S.V[b].z * c.c2 - S.V[a].z;
      double jj = (S.jdyn[j][0] - wr) * c.c0;
      double jOld = S.st.jacc[j].x;
      double jNew = dclamp(jOld + jj, -c.c1, c.c1);
      S.st.jacc[j].x = jNew;
      jj = jNew - jOld;
      S.V[a].z -= jj * c.ia * c.c3;
      S.V[b].z += jj * c.ib;
    } break;
    case MG_JOINT_ROTARY_SPRING: {
      double wrn = S.V[a].z - S.V[b].z;
      double w_damp = (S.jdyn[j][0] - wrn) * c.c2;
      S.jdyn[j][0] = wrn + w_damp;
      double j_damp = w_damp * c.c0;
      S.st.jacc[j].x += j_damp;
      S.V[a].z += j_damp * c.ia;
      S.V[b].z -= j_damp * c.ib;
    } break;
    case MG_JOINT_PIN: {
      d2 r1 = D2(S.jdyn[j][0], S.jdyn[j][1]), r2 = D2(S.jdyn[j][2], S.jdyn[j][3]);
      d2 n = D2(S.jdyn[j][4], S.jdyn[j][5]);
      double4 va = S.V[a], vb = S.V[b];
      d2 v1 = dadd(D2(va.x, va.y), dmul(dperp(r1), va.z));
      d2 v2 = dadd(D2(vb.x, vb.y), dmul(dperp(r2), vb.z));
      double vrn = ddot(dsub(v2, v1), n);
      double jn = (S.jdyn[j][7] - vrn) * S.jdyn[j][6];
      double jnOld = S.st.jacc[j].x;
      double jnNew = dclamp(jnOld + jn, -c.c1, c.c1);
      S.st.jacc[j].x = jnNew;
      jn = jnNew - jnOld;
      double jx = n.x * jn, jy = n.y * jn;
      va.x = va.x + (-jx) * c.ma; va.y = va.y + (-jy) * c.ma;
      va.z += c.ia * (r1.x * (-jy) - r1.y * (-jx));
      vb.x = vb.x + jx * c.mb; vb.y = vb.y + jy * c.mb;
      vb.z += c.ib * (r2.x * jy - r2.y * jx);
      S.V[a] = va;
      S.V[b] = vb;
    } break;
    case MG_JOINT_ROTARY_LIMIT: {
      double bias = S.jdyn[j][0];
      if (!bias) return;
      double wr = S.V[b].z - S.V[a].z;
      double jj = -(bias + wr) * c.c0;
      double jOld = S.st.jacc[j].x;
      double jNew = (bias < 0.0) ? dclamp(jOld + jj, 0.0, c.c1) : dclamp(jOld + jj, -c.c1, 0.0);
      S.st.jacc[j].x = jNew;
      jj = jNew - jOld;
      S.V[a].z -= jj * c.ia;
      S.V[b].z += jj * c.ib;
    } break;
    case MG_JOINT_MOTOR: {
      double wr = S.V[b].z - S.V[a].z + S.jdyn[j][0];
      double jj = -wr * c.c0;
      double jOld = S.st.jacc[j].x;
      double jNew = dclamp(jOld + jj, -c.c1, c.c1);
      S.st.jacc[j].x = jNew;
      jj = jNew - jOld;
      S.V[a].z -= jj * c.ia;
      S.V[b].z += jj * c.ib;
    } break;
  }
}

/* ------------------------------------------------------------------ kernel */
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 4)
k_physics(EnvState* __restrict__ states, const DeviceScene* __restrict__ scenes, const int32_t* __restrict__ actions,
          int batch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * WARPS + warp;
  if (env >= batch) return; /* warp-uniform; no block-wide barriers are used */
  EnvSmem& S = reinterpret_cast<EnvSmem*>(smem_raw)[warp];
  EnvState* G = states + env;

  /* ---- stream the state record in: coalesced 128-bit loads */
  {
    const uint4* src = reinterpret_cast<const uint4*>(G);
    uint4* dst = reinterpret_cast<uint4*>(&S.st);
    constexpr int N16 = sizeof(EnvState) / 16;
#pragma unroll
    for (int i = 0; i < (N16 + 31) / 32; i++) {
      int k = i * 32 + lane;
      if (k < N16) dst[k] = src[k];
    }
  }
  __syncwarp();
  const DeviceScene* ds = scenes + S.st.scene;
  const mg_scene_t& sc = ds->s;
  const int nb = sc.n_bodies, ns = sc.n_shapes, nj = sc.n_joints, ncg = sc.n_cgroups, nbp = sc.n_bpairs;
  if (lane <= MG_MAX_BODIES) {
    S.MI[lane] = (lane < nb) ? make_double2(sc.bodies[lane].m_inv, sc.bodies[lane].i_inv) : make_double2(0.0, 0.0);
    S.V[lane] = (lane < MG_MAX_BODIES) ? S.st.V[lane] : make_double4(0.0, 0.0, 0.0, 0.0);
    S.Bv[lane] = (lane < MG_MAX_BODIES) ? S.st.Bv[lane] : make_double4(0.0, 0.0, 0.0, 0.0);
    S.path[lane] = 0.0f;
  }
  int n_sep = 0;
  const double dt = MG_DT;

  /* ---- Robot.set_action: id = 9*grip + 3*lr + ud (entities.py:162-186, 439-457) */
  int action = actions[env];
  action = action < 0 ? 0 : (action > 17 ? 17 : action);
  const int ud = action % 3, lr = (action / 3) % 3, grip = action / 9;
  const double Rr = sc.robot_radius;
  double target_speed = 0.0, rel_turn = 0.0;
  if (ud == 1) target_speed += 4.0 * Rr;
  if (ud == 2) target_speed -= 3.0 * Rr;
  if (lr == 1) rel_turn += 1.5;
  if (lr == 2) rel_turn -= 1.5;
  const double target_finger = (grip == 0) ? (3.14159265358979323846 / 8) : -0.0;
  const int robot = sc.robot_body, control = sc.control_body;
  int stamp = S.st.stamp;
  int n_arb = S.st.n_arb;
  int overflow = S.st.overflow;
  int ncon = 0;
  __syncwarp();

  for (int sub = 0; sub < MG_SUBSTEPS; ++sub) {
    stamp++;
    /* ---- Robot.update (entities.py:459-479) */
    if (lane == 0) {
      double4 Pc = S.st.P[control];
      Pc.z = S.st.P[robot].z + rel_turn;
      S.st.P[control] = Pc;
      double2 Rb = S.st.R[robot];
      double4 Vc = S.V[control];
      Vc.x = Rb.x * 0.0 - Rb.y * target_speed;
      Vc.y = Rb.x * target_speed + Rb.y * 0.0;
      S.V[control] = Vc;
    } else if (lane <= 2) {
      int f = lane - 1;
      double side = (f == 0) ? -1.0 : 1.0;
      double rel_angle = S.st.P[sc.finger_body[f]].z - S.st.P[robot].z;
      double angle_error = rel_angle + side * target_finger;
      double target_rate = dmaxf(-1, dminf(1, angle_error * 10));
      if (fabs(target_rate) < 1e-4) target_rate = 0.0;
      S.jdyn[sc.motor_joint[f]][0] = target_rate;
    }
    __syncwarp();

    /* ---- integrate positions (cpBodyUpdatePosition; kinematic control body included) */
    if (lane < nb) {
      double4 P = S.st.P[lane], V = S.V[lane], Bv = S.Bv[lane];
      P.x = P.x + (V.x + Bv.x) * dt;
      P.y = P.y + (V.y + Bv.y) * dt;
      P.z = P.z + (V.z + Bv.z) * dt;
      double sn, cs;
      mg_det_sincos(P.z, &sn, &cs);
      S.st.P[lane] = P;
      S.st.R[lane] = make_double2(cs, sn);
      S.Bv[lane] = make_double4(0.0, 0.0, 0.0, 0.0);
      /* how far can any point of this body's shapes have moved: |dp|_1 + reach * |dtheta|, rounded up */
      double moved = (fabs(V.x + Bv.x) + fabs(V.y + Bv.y) + ds->aux.body_reach[lane] * fabs(V.z + Bv.z)) * dt;
      S.path[lane] = __fadd_ru(S.path[lane], __double2float_ru(moved * 1.000001));
    }
    __syncwarp();

    /* ---- shape boxes (conservative fp32; the exact fp64 test is redone per candidate) */
    for (int si = lane; si < ns; si += 32) {
      if (sc.shapes[si].body >= 0) {
        ShapeView v = make_view(S, ds, si);
        double bb[4];
        sv_bb(v, bb);
        S.sbb[si] = make_float4(__double2float_rd(bb[0]), __double2float_rd(bb[1]), __double2float_ru(bb[2]),
                                __double2float_ru(bb[3]));
      } else {
        S.sbb[si] = make_float4(ds->aux.static_bb[si][0], ds->aux.static_bb[si][1], ds->aux.static_bb[si][2],
                                ds->aux.static_bb[si][3]);
      }
    }
    __syncwarp();
    if (lane < ncg) {
      int s0 = sc.cgroups[lane].shape0, n = sc.cgroups[lane].nshape;
      float4 g = S.sbb[s0];
      for (int k = 1; k < n; k++) {
        float4 t = S.sbb[s0 + k];
        g.x = fminf(g.x, t.x); g.y = fminf(g.y, t.y); g.z = fmaxf(g.z, t.z); g.w = fmaxf(g.w, t.w);
      }
      S.gbb[lane] = g;
    }
    __syncwarp();

    /* ---- broadphase: canonical pair list -> candidate shape pairs, order preserved */
    int ncand = 0;
    for (int base = 0; base < nbp; base += 32) {
      int p = base + lane;
      int cnt = 0, sa0 = 0, na = 0, sb0 = 0, nbs = 0;
      if (p < nbp) {
        int ga = sc.bpairs[p][0], gb = sc.bpairs[p][1];
        if (f4_overlap(S.gbb[ga], S.gbb[gb])) {
          sa0 = sc.cgroups[ga].shape0; na = sc.cgroups[ga].nshape;
          sb0 = sc.cgroups[gb].shape0; nbs = sc.cgroups[gb].nshape;
          for (int i = 0; i < na; i++)
            for (int k = 0; k < nbs; k++) cnt += f4_overlap(S.sbb[sa0 + i], S.sbb[sb0 + k]) ? 1 : 0;
        }
      }
      int incl = warp_incl_scan(cnt, lane);
      int total = __shfl_sync(FULL, incl, 31);
      if (cnt) {
        int w = ncand + incl - cnt;
        for (int i = 0; i < na; i++)
          for (int k = 0; k < nbs; k++)
            if (f4_overlap(S.sbb[sa0 + i], S.sbb[sb0 + k])) {
              if (w < MG_NCAND) { S.cand[w][0] = (uint8_t)(sa0 + i); S.cand[w][1] = (uint8_t)(sb0 + k); }
              w++;
            }
      }
      ncand += total;
    }
    if (ncand > MG_NCAND) { ncand = MG_NCAND; overflow |= 1; }
    __syncwarp();

    /* ---- narrowphase + arbiter cache lookup (cpCollide + cpArbiterUpdate) */
    ncon = 0;
    int narb_new = 0;
    for (int base = 0; base < ncand; base += 32) {
      int c = base + lane;
      Manifold m;
      m.count = 0;
      int ia = 0, ib = 0;
      int sep_slot = -1;        /* existing separation-cache entry of this pair */
      float sep_limit = -1.0f;  /* > 0: (re)write the entry with this expiry */
      bool skipped = false;     /* cached separation still valid: narrowphase not run */
      if (c < ncand) {
        ia = S.cand[c][0]; ib = S.cand[c][1];
        if (sc.shapes[ia].kind > sc.shapes[ib].kind) { int t = ia; ia = ib; ib = t; }
        const int sba = sc.shapes[ia].body < 0 ? SLOT_STATIC : sc.shapes[ia].body;
        const int sbb = sc.shapes[ib].body < 0 ? SLOT_STATIC : sc.shapes[ib].body;
        const float travelled = __fadd_ru(S.path[sba], S.path[sbb]);
        for (int k = 0; k < n_sep; k++)
          if (S.sep[k].a == ia && S.sep[k].b == ib) { sep_slot = k; skipped = travelled < S.sep[k].limit; }
        if (!skipped) {
          ShapeView va = make_view(S, ds, ia), vb = make_view(S, ds, ib);
          double bba[4], bbb[4];
          sv_bb(va, bba);
          sv_bb(vb, bbb);
          if (bb_intersects(bba, bbb)) {
            mg_collide(va, vb, bba, bbb, m);
            /* shapes `margin` apart cannot touch until the bodies have travelled that far */
            if (m.count == 0 && m.margin > 1e-6)
              sep_limit = __fadd_rd(travelled, __double2float_rd(m.margin * 0.999999 - 1e-9));
          }
        }
      }
      {
        /* update the separation cache (existing entries in place, new ones appended in lane order) */
        unsigned want_new = __ballot_sync(FULL, sep_limit > 0.0f && sep_slot < 0);
        if (sep_limit > 0.0f) {
          int k = sep_slot >= 0 ? sep_slot : n_sep + __popc(want_new & ((1u << lane) - 1u));
          if (k < MG_NSEP) { S.sep[k].a = (uint8_t)ia; S.sep[k].b = (uint8_t)ib; S.sep[k].pad_ = 0; S.sep[k].limit = sep_limit; }
        } else if (sep_slot >= 0 && !skipped) {
          S.sep[sep_slot].limit = -1.0f; /* measured again: touching, or too close to cache */
        }
        n_sep = min(n_sep + __popc(want_new), MG_NSEP);
        __syncwarp();
      }
      unsigned has = __ballot_sync(FULL, m.count > 0);
      int arb_idx = narb_new + __popc(has & ((1u << lane) - 1u));
      int incl = warp_incl_scan(m.count, lane);
      int con_off = ncon + incl - m.count;
      int total = __shfl_sync(FULL, incl, 31);
      if (m.count > 0) {
        if (arb_idx < MG_NARB && con_off + m.count <= MG_NCON) {
          int ba = sc.shapes[ia].body, bb = sc.shapes[ib].body;
          d2 pa = (ba >= 0) ? D2(S.st.P[ba].x, S.st.P[ba].y) : D2(0, 0);
          d2 pb = (bb >= 0) ? D2(S.st.P[bb].x, S.st.P[bb].y) : D2(0, 0);
          if (ba < 0) ba = SLOT_STATIC;
          if (bb < 0) bb = SLOT_STATIC;
          int found = -1;
          for (int k = 0; k < n_arb; k++)
            if (S.st.arb[k].a == ia && S.st.arb[k].b == ib) found = k;
          bool first = !(found >= 0 && S.st.arb[found].stamp == stamp - 1);
          ArbEntry ne;
          ne.a = (uint8_t)ia; ne.b = (uint8_t)ib; ne.count = (uint8_t)m.count; ne.pad_ = 0; ne.stamp = stamp;
          ne.hash[0] = ne.hash[1] = 0; ne.jn[0] = ne.jn[1] = ne.jt[0] = ne.jt[1] = 0.0;
          double u = sc.shapes[ia].friction * sc.shapes[ib].friction;
#pragma unroll
          for (int i = 0; i < 2; i++) {
            if (i < m.count) {
              double jn = 0.0, jt = 0.0;
              if (found >= 0) {
                int oc = S.st.arb[found].count;
                for (int q = 0; q < oc; q++)
                  if (m.hash[i] == S.st.arb[found].hash[q]) { jn = S.st.arb[found].jn[q]; jt = S.st.arb[found].jt[q]; }
              }
              ConSmem& C = S.con[con_off + i];
              d2 r1 = dsub(m.p1[i], pa), r2 = dsub(m.p2[i], pb);
              C.r1x = r1.x; C.r1y = r1.y; C.r2x = r2.x; C.r2y = r2.y;
              C.nx = m.n.x; C.ny = m.n.y; C.jn = jn; C.jt = jt; C.u = u; C.hash = m.hash[i];
              C.ba = (uint8_t)ba; C.bb = (uint8_t)bb;
              C.arb = (uint8_t)arb_idx; C.slot = (uint8_t)i; C.first = first ? 1 : 0;
              ne.hash[i] = m.hash[i];
            }
          }
          if (found >= 0) S.st.arb[found].pad_ = 1; /* consumed */
          S.arb2[arb_idx] = ne;
        }
      }
      narb_new += __popc(has);
      ncon += total;
    }
    if (narb_new > MG_NARB || ncon > MG_NCON) {
      /* capacity exceeded (never seen on the registered tasks): flag the env and solve this
       * sub-step without contacts rather than with a partial, order-dependent subset */
      overflow |= 2;
      narb_new = 0;
      ncon = 0;
    }
    __syncwarp();
    /* survivors of the old cache (cpSpaceArbiterSetFilter: keep while ticks < persistence) */
    {
      bool alive = false;
      ArbEntry e;
      if (lane < n_arb) {
        e = S.st.arb[lane];
        alive = (e.pad_ == 0) && (stamp - e.stamp) < MG_PERSISTENCE;
      }
      unsigned am = __ballot_sync(FULL, alive);
      int pos = narb_new + __popc(am & ((1u << lane) - 1u));
      if (alive && pos < MG_NARB) S.arb2[pos] = e;
      int tot = narb_new + __popc(am);
      if (tot > MG_NARB) { tot = MG_NARB; overflow |= 4; }
      n_arb = tot;
    }
    __syncwarp();

    /* ---- dependency levels of the contacts (sequential order kept; disjoint contacts share a level) */
    if (lane == 0) {
      for (int b = 0; b <= MG_MAX_BODIES; b++) S.blevel[b] = 0;
      int mx = 0;
      for (int c = 0; c < ncon; c++) {
        int ba = S.con[c].ba, bb = S.con[c].bb;
        /* bodies without inverse mass (static, kinematic) are never written: no dependency through them */
        int la = (S.MI[ba].x != 0.0 || S.MI[ba].y != 0.0) ? S.blevel[ba] : 0;
        int lb = (S.MI[bb].x != 0.0 || S.MI[bb].y != 0.0) ? S.blevel[bb] : 0;
        int lv = (la > lb ? la : lb) + 1;
        S.blevel[ba] = (uint8_t)lv;
        S.blevel[bb] = (uint8_t)lv;
        S.clevel[c] = (uint8_t)lv;
        if (lv > mx) mx = lv;
      }
      S.max_clevel = mx;
    }
    __syncwarp();
    const int max_clevel = S.max_clevel;

    /* ---- contact prestep: everything a contact needs for the solve lives in this lane's registers */
    double c_r1x = 0, c_r1y = 0, c_r2x = 0, c_r2y = 0, c_nx = 0, c_ny = 0, c_nMass = 0, c_tMass = 0, c_bias = 0;
    double c_jn = 0, c_jt = 0, c_jb = 0, c_u = 0, c_ma = 0, c_ia = 0, c_mb = 0, c_ib = 0;
    int c_ba = SLOT_STATIC, c_bb = SLOT_STATIC, c_level = 0, c_first = 1;
    if (lane < ncon) {
      const ConSmem& C = S.con[lane];
      c_r1x = C.r1x; c_r1y = C.r1y; c_r2x = C.r2x; c_r2y = C.r2y; c_nx = C.nx; c_ny = C.ny;
      c_jn = C.jn; c_jt = C.jt; c_u = C.u; c_first = C.first;
      c_ba = C.ba; c_bb = C.bb;
      c_level = S.clevel[lane];
      double2 ma = S.MI[c_ba], mb = S.MI[c_bb];
      c_ma = ma.x; c_ia = ma.y; c_mb = mb.x; c_ib = mb.y;
      d2 r1 = D2(c_r1x, c_r1y), r2 = D2(c_r2x, c_r2y), n = D2(c_nx, c_ny), t = dperp(n);
      double rcn1 = dcross(r1, n), rcn2 = dcross(r2, n);
      c_nMass = 1.0 / ((c_ma + c_ia * rcn1 * rcn1) + (c_mb + c_ib * rcn2 * rcn2));
      double rct1 = dcross(r1, t), rct2 = dcross(r2, t);
      c_tMass = 1.0 / ((c_ma + c_ia * rct1 * rct1) + (c_mb + c_ib * rct2 * rct2));
      d2 pa = (c_ba < MG_MAX_BODIES) ? D2(S.st.P[c_ba].x, S.st.P[c_ba].y) : D2(0, 0);
      d2 pb = (c_bb < MG_MAX_BODIES) ? D2(S.st.P[c_bb].x, S.st.P[c_bb].y) : D2(0, 0);
      d2 body_delta = dsub(pb, pa);
      double dist = ddot(dadd(dsub(r2, r1), body_delta), n);
      c_bias = -ds->aux.contact_bias_coef * dminf(0.0, dist + MG_COLLISION_SLOP) / dt;
      c_jb = 0.0;
    }
    /* ---- joint prestep: parallel except the rotary springs, whose preStep applies an impulse */
    for (int j = lane; j < nj; j += 32) joint_prestep(S, ds, j);
    __syncwarp();
    if (lane == 0) {
      for (int k = 0; k < ds->aux.n_springs; k++) spring_prestep(S, ds, ds->aux.springs[k]);
    }
    __syncwarp();

    /* ---- warm start (cpArbiterApplyCachedImpulse, then the joints' applyCachedImpulse; dt_coef = 1) */
    for (int L = 1; L <= max_clevel; L++) {
      if (lane < ncon && c_level == L && !c_first) {
        double jx = c_nx * c_jn - c_ny * c_jt, jy = c_nx * c_jt + c_ny * c_jn;
        apply_imp(S, c_ba, c_ma, c_ia, -jx, -jy, c_r1x, c_r1y);
        apply_imp(S, c_bb, c_mb, c_ib, jx, jy, c_r2x, c_r2y);
      }
      __syncwarp();
    }
    const int n_levels = ds->aux.n_levels;
    for (int L = 0; L < n_levels; L++) {
      int j = ds->aux.sched[L][lane];
      if (j != 255) joint_warm(S, ds, j);
      __syncwarp();
    }

    /* ---- solver iterations (cpArbiterApplyImpulse for every arbiter, then every joint) */
    for (int it = 0; it < MG_ITERATIONS; ++it) {
      for (int L = 1; L <= max_clevel; L++) {
        if (lane < ncon && c_level == L) {
          double4 va = S.V[c_ba], vb = S.V[c_bb];
          double4 ba_ = S.Bv[c_ba], bb_ = S.Bv[c_bb];
          /* vb1 = a.v_bias + perp(r1)*a.w_bias, etc. */
          double vb1x = ba_.x + (-c_r1y) * ba_.z, vb1y = ba_.y + c_r1x * ba_.z;
          double vb2x = bb_.x + (-c_r2y) * bb_.z, vb2y = bb_.y + c_r2x * bb_.z;
          double v1x = va.x + (-c_r1y) * va.z, v1y = va.y + c_r1x * va.z;
          double v2x = vb.x + (-c_r2y) * vb.z, v2y = vb.y + c_r2x * vb.z;
          double vrx = v2x - v1x, vry = v2y - v1y;
          double vbn = (vb2x - vb1x) * c_nx + (vb2y - vb1y) * c_ny;
          double vrn = vrx * c_nx + vry * c_ny;
          double vrt = vrx * (-c_ny) + vry * c_nx;
          double jbn = (c_bias - vbn) * c_nMass;
          double jbnOld = c_jb;
          c_jb = dmaxf(jbnOld + jbn, 0.0);
          double jn = -(0.0 + vrn) * c_nMass;
          double jnOld = c_jn;
          c_jn = dmaxf(jnOld + jn, 0.0);
          double jtMax = c_u * c_jn;
          double jt = -vrt * c_tMass;
          double jtOld = c_jt;
          c_jt = dclamp(jtOld + jt, -jtMax, jtMax);
          double bjx = c_nx * (c_jb - jbnOld), bjy = c_ny * (c_jb - jbnOld);
          apply_bias_imp(S, c_ba, c_ma, c_ia, -bjx, -bjy, c_r1x, c_r1y);
          apply_bias_imp(S, c_bb, c_mb, c_ib, bjx, bjy, c_r2x, c_r2y);
          double dn = c_jn - jnOld, dtt = c_jt - jtOld;
          double jx = c_nx * dn - c_ny * dtt, jy = c_nx * dtt + c_ny * dn;
          apply_imp(S, c_ba, c_ma, c_ia, -jx, -jy, c_r1x, c_r1y);
          apply_imp(S, c_bb, c_mb, c_ib, jx, jy, c_r2x, c_r2y);
        }
        __syncwarp();
      }
      for (int L = 0; L < n_levels; L++) {
        int j = ds->aux.sched[L][lane];
        if (j != 255) joint_apply(S, ds, j);
        __syncwarp();
      }
    }

    /* ---- persist the contact accumulators in the arbiter cache, make it current */
    if (lane < ncon) {
      const ConSmem& C = S.con[lane];
      S.arb2[C.arb].jn[C.slot] = c_jn;
      S.arb2[C.arb].jt[C.slot] = c_jt;
    }
    __syncwarp();
    {
      const uint4* src = reinterpret_cast<const uint4*>(S.arb2);
      uint4* dst = reinterpret_cast<uint4*>(S.st.arb);
      constexpr int N16 = sizeof(ArbEntry) * MG_NARB / 16;
      for (int k = lane; k < N16; k += 32) dst[k] = src[k];
    }
    __syncwarp();
  }

  /* ---- stream the record back */
  if (lane < MG_MAX_BODIES) {
    S.st.V[lane] = S.V[lane];
    S.st.Bv[lane] = S.Bv[lane];
  }
  if (lane == 0) {
    S.st.stamp = stamp;
    S.st.n_arb = n_arb;
    S.st.overflow = overflow;
    S.st.last_contacts = ncon;
  }
  __syncwarp();
  {
    uint4* dst = reinterpret_cast<uint4*>(G);
    const uint4* src = reinterpret_cast<const uint4*>(&S.st);
    constexpr int N16 = sizeof(EnvState) / 16;
#pragma unroll
    for (int i = 0; i < (N16 + 31) / 32; i++) {
      int k = i * 32 + lane;
      if (k < N16) dst[k] = src[k];
    }
  }
}

size_t mg_physics_smem_bytes(int warps) { return sizeof(EnvSmem) * (size_t)warps; }

cudaError_t mg_launch_physics(EnvState* states, const DeviceScene* scenes, const int32_t* actions, int batch,
                              cudaStream_t stream) {
  constexpr int WARPS = 4;
  static bool configured = false;
  size_t smem = mg_physics_smem_bytes(WARPS);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_physics<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_physics<WARPS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int grid = (batch + WARPS - 1) / WARPS;
  k_physics<WARPS><<<grid, WARPS * 32, smem, stream>>>(states, scenes, actions, batch);
  return cudaGetLastError();
}
