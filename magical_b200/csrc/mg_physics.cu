/*
 * mg_physics.cu — K1: fused 10-sub-step rigid-body physics for one env-step.
 *
 * Replaces, for a whole batch in one launch, the reference's
 *   Robot.set_action(flags)                         entities.py:439-457
 *   for i in range(10):                             base_env.py:236-243
 *       Robot.update(dt)                            entities.py:459-479
 *       pm.Space.step(dt)  -> Chipmunk cpSpaceStep  (third party)
 *
 * B200 mapping: ONE ENVIRONMENT PER (HALF-)WARP.  G lanes serve one env (G = 32,
 * or 16 so that two envs share a warp and the per-env register cost halves,
 * which doubles the environments in flight per SM -- the kernel is bound by
 * dependent latency, not by issue slots or HBM).  The env's state record
 * (3.1 KB) is streamed from HBM into shared memory with 128-bit loads once,
 * all 10 sub-steps x 10 solver iterations run out of shared memory /
 * registers, and the record is streamed back once.  Lanes map to bodies
 * (integration), shapes (bounding boxes), candidate pairs (broadphase,
 * narrowphase), contacts and joints (prestep, solve).  The sequential-impulse
 * Gauss-Seidel order of the reference is kept BIT FOR BIT: contacts and joints
 * that share no dynamic body commute, so they are grouped into dependency
 * levels and each level runs across lanes; levels run in order with a
 * (sub-)warp barrier between them.  Ballots / shuffles do the pair -> contact
 * compaction in canonical order.
 *
 * No tensor cores: there is no dense contraction on this path.
 */
#include "mg_device.cuh"
#include "mg_narrowphase.h"
#include "mg_sincos.h"

#define SLOT_STATIC MG_MAX_BODIES /* velocity slot of the static body: always reads as zero */
#define MG_NSEP 32
#ifndef MG_PHASES
#define MG_PHASES 4 /* block barriers per sub-step in the phase-aligned variant */
#endif

struct __align__(8) ConSmem {
  double r1x, r1y, r2x, r2y, nx, ny, jn, jt;
  double u;
  uint32_t hash;
  uint8_t ba, bb; /* body slots, 16 = static */
  uint8_t sa, sb; /* shape indices (cache key) */
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

struct BoxSmem {
  float4 sbb[MG_MAX_SHAPES];  /* conservative fp32 shape boxes (l, b, r, t) */
  float4 gbb[MG_MAX_CGROUPS]; /* collision-group boxes */
};

struct __align__(16) EnvSmem {
  double4 V[MG_MAX_BODIES + 1];  /* velocities (vx, vy, w); slot 16 = static body */
  double4 Bv[MG_MAX_BODIES + 1]; /* bias velocities */
  double4 P[MG_MAX_BODIES];      /* x, y, angle */
  double2 R[MG_MAX_BODIES];      /* cos, sin */
  double2 jacc[MG_MAX_JOINTS];
  CEntry cache[MG_NCACHE];
  double2 MI[MG_MAX_BODIES + 1]; /* m_inv, i_inv (0 for static / kinematic) */
  double jd[MG_MAX_JOINTS];      /* per-sub-step scalar of a joint: bias / motor rate / spring target */
  double pin[MG_MAX_PINS][8];    /* pin joints: r1, r2, n, nMass, bias */
  union {                        /* boxes are dead once the candidates exist; contacts are born after */
    BoxSmem bb;
    ConSmem con[MG_NCON];
  } u;
  SepEntry sep[MG_NSEP];
  float path[MG_MAX_BODIES + 4]; /* upper bound of the distance any point of the body has moved this launch */
  uint8_t cand[MG_NCAND][2];
  uint8_t blevel[MG_MAX_BODIES + 4];
  uint8_t clevel[MG_NCON];
  int32_t max_clevel;
  int32_t pad_[3];
};

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
    double4 P = S.P[b];
    double2 R = S.R[b];
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
__device__ __forceinline__ double ld_angle(const EnvSmem& S, int b) { return (b < MG_MAX_BODIES) ? S.P[b].z : 0.0; }

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
      S.jd[j] = dclamp(-ds->aux.j_bcoef[j] * (ld_angle(S, b) * ratio - ld_angle(S, a) - J.p0) / dt, -maxBias, maxBias);
    } break;
    case MG_JOINT_PIN: {
      double* pn = S.pin[ds->aux.jpin[j]];
      double2 Ra = (a < MG_MAX_BODIES) ? S.R[a] : make_double2(1.0, 0.0);
      double2 Rb = S.R[b];
      d2 r1 = D2(Ra.x * J.anchor_a[0] - Ra.y * J.anchor_a[1], Ra.y * J.anchor_a[0] + Ra.x * J.anchor_a[1]);
      d2 r2 = D2(Rb.x * J.anchor_b[0] - Rb.y * J.anchor_b[1], Rb.y * J.anchor_b[0] + Rb.x * J.anchor_b[1]);
      d2 pa = (a < MG_MAX_BODIES) ? D2(S.P[a].x, S.P[a].y) : D2(0, 0);
      d2 pb = D2(S.P[b].x, S.P[b].y);
      d2 delta = dsub(dadd(pb, r2), dadd(pa, r1));
      double dist = dlength(delta);
      d2 n = dmul(delta, 1.0 / (dist ? dist : MG_INF));
      double2 ma = S.MI[a], mb = S.MI[b];
      double rcn1 = dcross(r1, n), rcn2 = dcross(r2, n);
      double k = (ma.x + ma.y * rcn1 * rcn1) + (mb.x + mb.y * rcn2 * rcn2);
      double maxBias = J.max_bias;
      pn[0] = r1.x; pn[1] = r1.y; pn[2] = r2.x; pn[3] = r2.y;
      pn[4] = n.x; pn[5] = n.y;
      pn[6] = 1.0 / k;
      pn[7] = dclamp(-ds->aux.j_bcoef[j] * (dist - J.p0) / dt, -maxBias, maxBias);
    } break;
    case MG_JOINT_ROTARY_LIMIT: {
      double dist = ld_angle(S, b) - ld_angle(S, a);
      double pdist = 0.0;
      if (dist > J.p1) pdist = J.p1 - dist;
      else if (dist < J.p0) pdist = J.p0 - dist;
      double maxBias = J.max_bias;
      double bias = dclamp(-ds->aux.j_bcoef[j] * pdist / dt, -maxBias, maxBias);
      S.jd[j] = bias;
      if (!bias) S.jacc[j].x = 0.0;
    } break;
    default:
      break; /* pivot, motor: nothing per step; springs are pre-stepped sequentially */
  }
}

__device__ __forceinline__ void spring_prestep(EnvSmem& S, const DeviceScene* ds, int j) {
  const mg_joint_t& J = ds->s.joints[j];
  const int a = ds->aux.ja[j], b = ds->aux.jb[j];
  S.jd[j] = 0.0; /* target_wrn */
  double j_spring = ((ld_angle(S, a) - ld_angle(S, b)) - J.p0) * J.p1 * MG_DT;
  S.jacc[j].x = j_spring;
  S.V[a].z -= j_spring * S.MI[a].y;
  S.V[b].z += j_spring * S.MI[b].y;
}

__device__ __forceinline__ void joint_warm(EnvSmem& S, const DeviceScene* ds, int j) {
  const uint32_t hd = __ldg(&ds->aux.jpack[j]);
  const int kind = hd & 0xFF, a = (hd >> 8) & 0xFF, b = (hd >> 16) & 0xFF;
  const JC c = ld_jc(ds, j);
  double2 acc = S.jacc[j];
  if (kind == MG_JOINT_PIVOT) {
    apply_imp(S, a, c.ma, c.ia, -acc.x, -acc.y, 0.0, 0.0);
    apply_imp(S, b, c.mb, c.ib, acc.x, acc.y, 0.0, 0.0);
  } else if (kind == MG_JOINT_PIN) {
    const double* pn = S.pin[hd >> 24];
    double jx = pn[4] * acc.x, jy = pn[5] * acc.x;
    apply_imp(S, a, c.ma, c.ia, -jx, -jy, pn[0], pn[1]);
    apply_imp(S, b, c.mb, c.ib, jx, jy, pn[2], pn[3]);
  } else if (kind != MG_JOINT_ROTARY_SPRING) {
    /* gear / rotary limit / motor: a.w -= j*ia(*ratio_inv); b.w += j*ib */
    double da = acc.x * c.ia;
    if (kind == MG_JOINT_GEAR) da = da * c.c3;
    S.V[a].z -= da;
    S.V[b].z += acc.x * c.ib;
  }
}

/* One sequential-impulse update of joint j (the constraint's applyImpulse in Chipmunk).  The three purely
 * angular kinds share one branch-free path: every alternative is evaluated exactly as its reference
 * formula and selected, so the arithmetic of each kind is unchanged. */
__device__ __forceinline__ void joint_apply(EnvSmem& S, const DeviceScene* ds, int j) {
  const uint32_t hd = __ldg(&ds->aux.jpack[j]);
  const int kind = hd & 0xFF, a = (hd >> 8) & 0xFF, b = (hd >> 16) & 0xFF;
  const JC c = ld_jc(ds, j);
  if (kind == MG_JOINT_GEAR || kind == MG_JOINT_ROTARY_LIMIT || kind == MG_JOINT_MOTOR) {
    const bool gear = kind == MG_JOINT_GEAR, limit = kind == MG_JOINT_ROTARY_LIMIT;
    const double x = S.jd[j]; /* gear: bias; limit: bias; motor: rate */
    if (limit && !x) return;  /* joint not at a limit */
    const double wa = S.V[a].z, wb = S.V[b].z;
    double wr = gear ? (wb * c.c2 - wa) : (wb - wa);
    if (!gear && !limit) wr = wr + x;                 /* motor: b.w - a.w + rate */
    double jj = gear ? (x - wr) * c.c0 : (limit ? -(x + wr) * c.c0 : -wr * c.c0);
    const double lo = (limit && x < 0.0) ? 0.0 : -c.c1;
    const double hi = (limit && !(x < 0.0)) ? 0.0 : c.c1;
    const double jOld = S.jacc[j].x;
    const double jNew = dclamp(jOld + jj, lo, hi);
    S.jacc[j].x = jNew;
    jj = jNew - jOld;
    double da = jj * c.ia;
    if (gear) da = da * c.c3;
    S.V[a].z = wa - da;
    S.V[b].z = wb + jj * c.ib;
  } else if (kind == MG_JOINT_PIVOT) {
    double4 va = S.V[a], vb = S.V[b];
    double jx = (0.0 - (vb.x - va.x)) * c.c0;
    double jy = (0.0 - (vb.y - va.y)) * c.c0;
    double2 old = S.jacc[j];
    d2 acc = dvclamp(D2(old.x + jx, old.y + jy), c.c1);
    S.jacc[j] = make_double2(acc.x, acc.y);
    jx = acc.x - old.x; jy = acc.y - old.y;
    /* anchors are at the body origins: no angular part (adds exactly zero) */
    va.x = va.x + (-jx) * c.ma; va.y = va.y + (-jy) * c.ma;
    vb.x = vb.x + jx * c.mb; vb.y = vb.y + jy * c.mb;
    S.V[a] = va;
    S.V[b] = vb;
  } else if (kind == MG_JOINT_ROTARY_SPRING) {
    double wa = S.V[a].z, wb = S.V[b].z;
    double wrn = wa - wb;
    double w_damp = (S.jd[j] - wrn) * c.c2;
    S.jd[j] = wrn + w_damp;
    double j_damp = w_damp * c.c0;
    S.jacc[j].x += j_damp;
    S.V[a].z = wa + j_damp * c.ia;
    S.V[b].z = wb - j_damp * c.ib;
  } else { /* MG_JOINT_PIN */
    const double* pn = S.pin[hd >> 24];
    d2 r1 = D2(pn[0], pn[1]), r2 = D2(pn[2], pn[3]);
    d2 n = D2(pn[4], pn[5]);
    double4 va = S.V[a], vb = S.V[b];
    d2 v1 = dadd(D2(va.x, va.y), dmul(dperp(r1), va.z));
    d2 v2 = dadd(D2(vb.x, vb.y), dmul(dperp(r2), vb.z));
    double vrn = ddot(dsub(v2, v1), n);
    double jn = (pn[7] - vrn) * pn[6];
    double jnOld = S.jacc[j].x;
    double jnNew = dclamp(jnOld + jn, -c.c1, c.c1);
    S.jacc[j].x = jnNew;
    jn = jnNew - jnOld;
    double jx = n.x * jn, jy = n.y * jn;
    va.x = va.x + (-jx) * c.ma; va.y = va.y + (-jy) * c.ma;
    va.z += c.ia * (r1.x * (-jy) - r1.y * (-jx));
    vb.x = vb.x + jx * c.mb; vb.y = vb.y + jy * c.mb;
    vb.z += c.ib * (r2.x * jy - r2.y * jx);
    S.V[a] = va;
    S.V[b] = vb;
  }
}

/* Exact narrowphase of one candidate pair.  Deliberately OUT OF LINE: with the separation cache it runs
 * rarely, and keeping its ~9k instructions out of the solver loop's neighbourhood is what keeps the hot
 * code inside the 32 KB instruction cache. */
__device__ __noinline__ void narrow_pair(const EnvSmem* Sp, const DeviceScene* ds, int ia, int ib, Manifold* out) {
  const EnvSmem& S = *Sp;
  ShapeView va = make_view(S, ds, ia), vb = make_view(S, ds, ib);
  double bba[4], bbb[4];
  sv_bb(va, bba);
  sv_bb(vb, bbb);
  Manifold m;
  m.count = 0;
  m.margin = -1.0;
  m.n = D2(0, 0);
  if (bb_intersects(bba, bbb)) mg_collide(va, vb, bba, bbb, m);
  *out = m;
}

/* cooperative 16-byte copy by the G lanes of one environment */
template <int G>
__device__ __forceinline__ void gcopy16(void* dst, const void* src, int nbytes, int gl) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int k = gl; k < nbytes / 16; k += G) d[k] = s[k];
}

/* ------------------------------------------------------------------ kernel
 * G lanes per environment (32 or 16); THREADS / G environments per block.  With THREADS = 512 one block
 * fills an SM and two block barriers per sub-step keep all its environments in the same phase, so the
 * instruction working set at any moment is one phase (fits the 32 KB I-cache) instead of the whole
 * 60 KB sub-step body. */
template <int G, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 512 ? 1 : 4))
k_physics(EnvState* __restrict__ states, const DeviceScene* __restrict__ scenes, const int32_t* __restrict__ actions,
          int batch) {
  constexpr int NCONG = (G < MG_NCON) ? G : MG_NCON; /* contacts one group of lanes can own */
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);        /* lane within the environment's group */
  const int gbase = lane & ~(G - 1);    /* first lane of the group inside the warp */
  const unsigned gbits = (G == 32) ? 0xffffffffu : 0xffffu;
  const unsigned gmask = gbits << gbase;
  const int slot = threadIdx.x / G;
  const int env = blockIdx.x * (THREADS / G) + slot;
  constexpr bool PHASED = THREADS == 512;
  /* block barrier that tolerates the two half-warps of a warp arriving at different times
   * (__syncthreads() is the aligned form and would require the warp to be converged) */
#define BLOCK_PHASE_SYNC() asm volatile("barrier.sync 0, %0;" ::"r"(THREADS) : "memory")
  if (env >= batch) { /* group-uniform */
    if (PHASED)
      for (int sub = 0; sub < MG_SUBSTEPS; ++sub) {
        for (int k = 0; k < MG_PHASES; ++k) BLOCK_PHASE_SYNC();
      }
    return;
  }
  EnvSmem& S = reinterpret_cast<EnvSmem*>(smem_raw)[slot];
  EnvState* Gs = states + env;
#define GSYNC() __syncwarp(gmask)
  auto gballot = [&](bool p) -> unsigned { return (__ballot_sync(gmask, p) >> gbase) & gbits; };
  auto gscan = [&](int v) -> int { /* inclusive prefix sum over the group */
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
      int t = __shfl_up_sync(gmask, v, d, G);
      if (gl >= d) v += t;
    }
    return v;
  };
  auto glast = [&](int v) -> int { return __shfl_sync(gmask, v, G - 1, G); };

  /* ---- stream the state record in: coalesced 128-bit loads */
  const int scene_idx = Gs->scene;
  int stamp = Gs->stamp;
  int n_cache = Gs->n_cache;
  int overflow = Gs->overflow;
  gcopy16<G>(S.V, Gs->V, sizeof(Gs->V), gl);
  gcopy16<G>(S.Bv, Gs->Bv, sizeof(Gs->Bv), gl);
  gcopy16<G>(S.P, Gs->P, sizeof(Gs->P), gl);
  gcopy16<G>(S.R, Gs->R, sizeof(Gs->R), gl);
  gcopy16<G>(S.jacc, Gs->jacc, sizeof(Gs->jacc), gl);
  gcopy16<G>(S.cache, Gs->cache, sizeof(Gs->cache), gl);
  const DeviceScene* ds = scenes + scene_idx;
  const mg_scene_t& sc = ds->s;
  const int nb = sc.n_bodies, ns = sc.n_shapes, nj = sc.n_joints, ncg = sc.n_cgroups, nbp = sc.n_bpairs;
  for (int b = gl; b <= MG_MAX_BODIES; b += G) {
    S.MI[b] = (b < nb) ? make_double2(sc.bodies[b].m_inv, sc.bodies[b].i_inv) : make_double2(0.0, 0.0);
    S.path[b] = 0.0f;
    if (b == MG_MAX_BODIES) {
      S.V[b] = make_double4(0.0, 0.0, 0.0, 0.0);
      S.Bv[b] = make_double4(0.0, 0.0, 0.0, 0.0);
    }
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
  int ncon = 0;
  GSYNC();

  for (int sub = 0; sub < MG_SUBSTEPS; ++sub) {
    stamp++;
    /* ---- Robot.update (entities.py:459-479) */
    if (gl == 0) {
      double4 Pc = S.P[control];
      Pc.z = S.P[robot].z + rel_turn;
      S.P[control] = Pc;
      double2 Rb = S.R[robot];
      double4 Vc = S.V[control];
      Vc.x = Rb.x * 0.0 - Rb.y * target_speed;
      Vc.y = Rb.x * target_speed + Rb.y * 0.0;
      S.V[control] = Vc;
    } else if (gl <= 2) {
      int f = gl - 1;
      double side = (f == 0) ? -1.0 : 1.0;
      double rel_angle = S.P[sc.finger_body[f]].z - S.P[robot].z;
      double angle_error = rel_angle + side * target_finger;
      double target_rate = dmaxf(-1, dminf(1, angle_error * 10));
      if (fabs(target_rate) < 1e-4) target_rate = 0.0;
      S.jd[sc.motor_joint[f]] = target_rate;
    }
    GSYNC();

    /* ---- integrate positions (cpBodyUpdatePosition; kinematic control body included) */
    if (gl < nb) {
      double4 P = S.P[gl], V = S.V[gl], Bv = S.Bv[gl];
      P.x = P.x + (V.x + Bv.x) * dt;
      P.y = P.y + (V.y + Bv.y) * dt;
      P.z = P.z + (V.z + Bv.z) * dt;
      double sn, cs;
      mg_det_sincos(P.z, &sn, &cs);
      S.P[gl] = P;
      S.R[gl] = make_double2(cs, sn);
      S.Bv[gl] = make_double4(0.0, 0.0, 0.0, 0.0);
      /* how far can any point of this body's shapes have moved: |dp|_1 + reach * |dtheta|, rounded up */
      double moved = (fabs(V.x + Bv.x) + fabs(V.y + Bv.y) + ds->aux.body_reach[gl] * fabs(V.z + Bv.z)) * dt;
      S.path[gl] = __fadd_ru(S.path[gl], __double2float_ru(moved * 1.000001));
    }
    GSYNC();

    /* ---- shape boxes (conservative fp32; the exact fp64 test is redone per candidate) */
    for (int si = gl; si < ns; si += G) {
      if (sc.shapes[si].body >= 0) {
        ShapeView v = make_view(S, ds, si);
        double bb[4];
        sv_bb(v, bb);
        S.u.bb.sbb[si] = make_float4(__double2float_rd(bb[0]), __double2float_rd(bb[1]), __double2float_ru(bb[2]),
                                     __double2float_ru(bb[3]));
      } else {
        S.u.bb.sbb[si] = make_float4(ds->aux.static_bb[si][0], ds->aux.static_bb[si][1], ds->aux.static_bb[si][2],
                                     ds->aux.static_bb[si][3]);
      }
    }
    GSYNC();
    for (int g = gl; g < ncg; g += G) {
      int s0 = sc.cgroups[g].shape0, n = sc.cgroups[g].nshape;
      float4 bx = S.u.bb.sbb[s0];
      for (int k = 1; k < n; k++) {
        float4 t = S.u.bb.sbb[s0 + k];
        bx.x = fminf(bx.x, t.x); bx.y = fminf(bx.y, t.y); bx.z = fmaxf(bx.z, t.z); bx.w = fmaxf(bx.w, t.w);
      }
      S.u.bb.gbb[g] = bx;
    }
    GSYNC();

    /* ---- broadphase: canonical pair list -> candidate shape pairs, order preserved */
    int ncand = 0;
    for (int base = 0; base < nbp; base += G) {
      int p = base + gl;
      int cnt = 0, sa0 = 0, na = 0, sb0 = 0, nbs = 0;
      if (p < nbp) {
        int ga = sc.bpairs[p][0], gb = sc.bpairs[p][1];
        if (f4_overlap(S.u.bb.gbb[ga], S.u.bb.gbb[gb])) {
          sa0 = sc.cgroups[ga].shape0; na = sc.cgroups[ga].nshape;
          sb0 = sc.cgroups[gb].shape0; nbs = sc.cgroups[gb].nshape;
          for (int i = 0; i < na; i++)
            for (int k = 0; k < nbs; k++) cnt += f4_overlap(S.u.bb.sbb[sa0 + i], S.u.bb.sbb[sb0 + k]) ? 1 : 0;
        }
      }
      int incl = gscan(cnt);
      int total = glast(incl);
      if (cnt) {
        int w = ncand + incl - cnt;
        for (int i = 0; i < na; i++)
          for (int k = 0; k < nbs; k++)
            if (f4_overlap(S.u.bb.sbb[sa0 + i], S.u.bb.sbb[sb0 + k])) {
              if (w < MG_NCAND) { S.cand[w][0] = (uint8_t)(sa0 + i); S.cand[w][1] = (uint8_t)(sb0 + k); }
              w++;
            }
      }
      ncand += total;
    }
    if (ncand > MG_NCAND) { ncand = MG_NCAND; overflow |= 1; }
    GSYNC(); /* boxes are dead from here on: the union now holds contacts */
    if (PHASED && MG_PHASES >= 4) BLOCK_PHASE_SYNC();

    /* ---- narrowphase + contact cache lookup (cpCollide + cpArbiterUpdate) */
    ncon = 0;
    for (int base = 0; base < ncand; base += G) {
      int c = base + gl;
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
          narrow_pair(&S, ds, ia, ib, &m);
          /* shapes `margin` apart cannot touch until the bodies have travelled that far */
          if (m.count == 0 && m.margin > 1e-6)
            sep_limit = __fadd_rd(travelled, __double2float_rd(m.margin * 0.999999 - 1e-9));
        }
      }
      {
        /* update the separation cache (existing entries in place, new ones appended in lane order) */
        unsigned want_new = gballot(sep_limit > 0.0f && sep_slot < 0);
        if (sep_limit > 0.0f) {
          int k = sep_slot >= 0 ? sep_slot : n_sep + __popc(want_new & ((1u << gl) - 1u));
          if (k < MG_NSEP) { S.sep[k].a = (uint8_t)ia; S.sep[k].b = (uint8_t)ib; S.sep[k].pad_ = 0; S.sep[k].limit = sep_limit; }
        } else if (sep_slot >= 0 && !skipped) {
          S.sep[sep_slot].limit = -1.0f; /* measured again: touching, or too close to cache */
        }
        n_sep = min(n_sep + __popc(want_new), MG_NSEP);
      }
      int incl = gscan(m.count);
      int con_off = ncon + incl - m.count;
      int total = glast(incl);
      if (m.count > 0 && con_off + m.count <= NCONG) {
        int ba = sc.shapes[ia].body, bb = sc.shapes[ib].body;
        d2 pa = (ba >= 0) ? D2(S.P[ba].x, S.P[ba].y) : D2(0, 0);
        d2 pb = (bb >= 0) ? D2(S.P[bb].x, S.P[bb].y) : D2(0, 0);
        if (ba < 0) ba = SLOT_STATIC;
        if (bb < 0) bb = SLOT_STATIC;
        /* the pair's cached contacts all stem from its last collision; they are warm-started only if
         * that was the previous sub-step (cpArbiterApplyCachedImpulse skips first-contact arbiters) */
        bool first = true;
        double jn[2] = {0.0, 0.0}, jt[2] = {0.0, 0.0};
        for (int k = 0; k < n_cache; k++) {
          CEntry e = S.cache[k];
          if (e.a == ia && e.b == ib) {
            if (e.stamp == stamp - 1) first = false;
            if (m.hash[0] == e.hash) { jn[0] = e.jn; jt[0] = e.jt; }
            if (m.count > 1 && m.hash[1] == e.hash) { jn[1] = e.jn; jt[1] = e.jt; }
            S.cache[k].used = 1;
          }
        }
        double u = sc.shapes[ia].friction * sc.shapes[ib].friction;
#pragma unroll
        for (int i = 0; i < 2; i++) {
          if (i < m.count) {
            ConSmem& C = S.u.con[con_off + i];
            d2 r1 = dsub(m.p1[i], pa), r2 = dsub(m.p2[i], pb);
            C.r1x = r1.x; C.r1y = r1.y; C.r2x = r2.x; C.r2y = r2.y;
            C.nx = m.n.x; C.ny = m.n.y; C.jn = jn[i]; C.jt = jt[i]; C.u = u; C.hash = m.hash[i];
            C.ba = (uint8_t)ba; C.bb = (uint8_t)bb; C.sa = (uint8_t)ia; C.sb = (uint8_t)ib;
            C.first = first ? 1 : 0;
          }
        }
      }
      ncon += total;
      GSYNC();
    }
    if (ncon > NCONG) {
      /* capacity exceeded (never seen on the registered tasks): flag the env and solve this
       * sub-step without contacts rather than with a partial, order-dependent subset */
      overflow |= 2;
      ncon = 0;
    }
    GSYNC();

    if (PHASED && MG_PHASES >= 4) BLOCK_PHASE_SYNC();
    /* ---- dependency levels of the contacts (sequential order kept; disjoint contacts share a level) */
    if (gl == 0) {
      for (int b = 0; b <= MG_MAX_BODIES; b++) S.blevel[b] = 0;
      int mx = 0;
      for (int c = 0; c < ncon; c++) {
        int ba = S.u.con[c].ba, bb = S.u.con[c].bb;
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
    GSYNC();
    const int max_clevel = S.max_clevel;

    /* ---- contact prestep: everything a contact needs for the solve lives in this lane's registers */
    double c_r1x = 0, c_r1y = 0, c_r2x = 0, c_r2y = 0, c_nx = 0, c_ny = 0, c_nMass = 0, c_tMass = 0, c_bias = 0;
    double c_jn = 0, c_jt = 0, c_jb = 0, c_u = 0, c_ma = 0, c_ia = 0, c_mb = 0, c_ib = 0;
    int c_ba = SLOT_STATIC, c_bb = SLOT_STATIC, c_level = 0, c_first = 1, c_sa = 0, c_sb = 0;
    uint32_t c_hash = 0;
    if (gl < ncon) {
      const ConSmem& C = S.u.con[gl];
      c_r1x = C.r1x; c_r1y = C.r1y; c_r2x = C.r2x; c_r2y = C.r2y; c_nx = C.nx; c_ny = C.ny;
      c_jn = C.jn; c_jt = C.jt; c_u = C.u; c_first = C.first;
      c_ba = C.ba; c_bb = C.bb; c_sa = C.sa; c_sb = C.sb; c_hash = C.hash;
      c_level = S.clevel[gl];
      double2 ma = S.MI[c_ba], mb = S.MI[c_bb];
      c_ma = ma.x; c_ia = ma.y; c_mb = mb.x; c_ib = mb.y;
      d2 r1 = D2(c_r1x, c_r1y), r2 = D2(c_r2x, c_r2y), n = D2(c_nx, c_ny), t = dperp(n);
      double rcn1 = dcross(r1, n), rcn2 = dcross(r2, n);
      c_nMass = 1.0 / ((c_ma + c_ia * rcn1 * rcn1) + (c_mb + c_ib * rcn2 * rcn2));
      double rct1 = dcross(r1, t), rct2 = dcross(r2, t);
      c_tMass = 1.0 / ((c_ma + c_ia * rct1 * rct1) + (c_mb + c_ib * rct2 * rct2));
      d2 pa = (c_ba < MG_MAX_BODIES) ? D2(S.P[c_ba].x, S.P[c_ba].y) : D2(0, 0);
      d2 pb = (c_bb < MG_MAX_BODIES) ? D2(S.P[c_bb].x, S.P[c_bb].y) : D2(0, 0);
      d2 body_delta = dsub(pb, pa);
      double dist = ddot(dadd(dsub(r2, r1), body_delta), n);
      c_bias = -ds->aux.contact_bias_coef * dminf(0.0, dist + MG_COLLISION_SLOP) / dt;
      c_jb = 0.0;
    }
    /* ---- joint prestep: parallel except the rotary springs, whose preStep applies an impulse */
    for (int j = gl; j < nj; j += G) joint_prestep(S, ds, j);
    GSYNC();
    if (gl == 0) {
      for (int k = 0; k < ds->aux.n_springs; k++) spring_prestep(S, ds, ds->aux.springs[k]);
    }
    GSYNC();

    if (PHASED) BLOCK_PHASE_SYNC(); /* everyone enters the solver together */

    /* ---- warm start (cpArbiterApplyCachedImpulse, then the joints' applyCachedImpulse; dt_coef = 1) */
    for (int L = 1; L <= max_clevel; L++) {
      if (gl < ncon && c_level == L && !c_first) {
        double jx = c_nx * c_jn - c_ny * c_jt, jy = c_nx * c_jt + c_ny * c_jn;
        apply_imp(S, c_ba, c_ma, c_ia, -jx, -jy, c_r1x, c_r1y);
        apply_imp(S, c_bb, c_mb, c_ib, jx, jy, c_r2x, c_r2y);
      }
      GSYNC();
    }
    /* joints: lane l walks the joints of connected component l in insertion order; components share
     * no dynamic body, so no barrier is needed until the contacts run again */
    const int n_levels = ds->aux.n_levels;
    for (int L = 0; L < n_levels; L++) {
      int j = ds->aux.sched[L][gl];
      if (j == 255) break;
      joint_warm(S, ds, j);
    }
    GSYNC();

    /* ---- solver iterations (cpArbiterApplyImpulse for every arbiter, then every joint) */
    for (int it = 0; it < MG_ITERATIONS; ++it) {
      for (int L = 1; L <= max_clevel; L++) {
        if (gl < ncon && c_level == L) {
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
        GSYNC();
      }
      for (int L = 0; L < n_levels; L++) {
        int j = ds->aux.sched[L][gl];
        if (j == 255) break;
        joint_apply(S, ds, j);
      }
      GSYNC();
    }

    /* ---- rebuild the contact cache: this sub-step's contacts first (canonical order), then the
     * entries of pairs that did not collide now and are younger than the persistence window
     * (cpSpaceArbiterSetFilter).  Old entries are read into registers before anything is written. */
    {
      constexpr int R = (MG_NCACHE + G - 1) / G;
      CEntry keep[R];
      bool alive[R];
      int pos[R];
      int nsurv = 0;
#pragma unroll
      for (int r = 0; r < R; r++) {
        int k = r * G + gl;
        alive[r] = false;
        if (k < n_cache) {
          keep[r] = S.cache[k];
          alive[r] = (keep[r].used == 0) && (stamp - keep[r].stamp) < MG_PERSISTENCE;
        }
        unsigned am = gballot(alive[r]);
        pos[r] = ncon + nsurv + __popc(am & ((1u << gl) - 1u));
        nsurv += __popc(am);
      }
      GSYNC();
      if (gl < ncon) {
        CEntry e;
        e.a = (uint8_t)c_sa; e.b = (uint8_t)c_sb; e.used = 0; e.pad_ = 0; e.hash = c_hash; e.stamp = stamp; e.pad2_ = 0;
        e.jn = c_jn; e.jt = c_jt;
        S.cache[gl] = e;
      }
#pragma unroll
      for (int r = 0; r < R; r++)
        if (alive[r] && pos[r] < MG_NCACHE) S.cache[pos[r]] = keep[r];
      int tot = ncon + nsurv;
      if (tot > MG_NCACHE) { tot = MG_NCACHE; overflow |= 4; }
      n_cache = tot;
    }
    GSYNC();
    if (PHASED) BLOCK_PHASE_SYNC(); /* everyone leaves the solver together */
  }

  /* ---- stream the record back */
  gcopy16<G>(Gs->V, S.V, sizeof(Gs->V), gl);
  gcopy16<G>(Gs->Bv, S.Bv, sizeof(Gs->Bv), gl);
  gcopy16<G>(Gs->P, S.P, sizeof(Gs->P), gl);
  gcopy16<G>(Gs->R, S.R, sizeof(Gs->R), gl);
  gcopy16<G>(Gs->jacc, S.jacc, sizeof(Gs->jacc), gl);
  gcopy16<G>(Gs->cache, S.cache, sizeof(Gs->cache), gl);
  if (gl == 0) {
    Gs->stamp = stamp;
    Gs->n_cache = n_cache;
    Gs->overflow = overflow;
    Gs->last_contacts = ncon;
  }
#undef GSYNC
#undef BLOCK_PHASE_SYNC
}

size_t mg_physics_smem_bytes(int lanes_per_env, int threads) { return sizeof(EnvSmem) * (size_t)(threads / lanes_per_env); }

template <int G, int THREADS>
static cudaError_t launch_g(EnvState* states, const DeviceScene* scenes, const int32_t* actions, int batch,
                            cudaStream_t stream) {
  static bool configured_on[64] = {false}; /* per device: the attribute does not carry over to other GPUs */
  int dev = 0;
  cudaError_t de = cudaGetDevice(&dev);
  if (de != cudaSuccess) return de;
  bool& configured = configured_on[dev & 63];
  size_t smem = mg_physics_smem_bytes(G, THREADS);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_physics<G, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_physics<G, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int epb = THREADS / G;
  k_physics<G, THREADS><<<(batch + epb - 1) / epb, THREADS, smem, stream>>>(states, scenes, actions, batch);
  return cudaGetLastError();
}

cudaError_t mg_launch_physics(EnvState* states, const DeviceScene* scenes, const int32_t* actions, int batch,
                              int lanes_per_env, int block_threads, cudaStream_t stream) {
  if (lanes_per_env == 16) {
    if (block_threads == 512) return launch_g<16, 512>(states, scenes, actions, batch, stream);
    return launch_g<16, 128>(states, scenes, actions, batch, stream);
  }
  if (block_threads == 512) return launch_g<32, 512>(states, scenes, actions, batch, stream);
  return launch_g<32, 128>(states, scenes, actions, batch, stream);
}
