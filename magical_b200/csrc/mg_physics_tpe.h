/*
 * mg_physics_tpe.h — K1 body, "thread per environment" form: one env-step (Robot.set_action + 10 x
 * (Robot.update + Chipmunk-equivalent space step)) of ONE environment executed by ONE thread.
 *
 * Replaces the same reference calls as mg_physics.cu (entities.py:439-479, base_env.py:236-243 and
 * pymunk's Space.step).  Why this shape on B200: the sequential-impulse solver is a dependent chain
 * (10 sub-steps x 10 iterations x (contacts, then the robot's 10 joints which all touch the robot
 * body)), so lanes that cooperate on one environment mostly wait for each other (ncu, round 1: 7 of 32
 * lanes active, 25 % of stall samples at barriers).  Here every lane of a warp owns a DIFFERENT
 * environment, all lanes walk the same instruction stream (environments of a batch share a scene
 * structure), and nothing is ever exchanged between lanes, so there are no barriers at all.
 *
 * Memory plan per environment
 *   - private 8-byte words in shared memory, laid out [word][lane] (conflict-free, dynamic indexing is
 *     free): velocities + bias velocities + pose/rotation of every body that owns collision shapes,
 *     the blocks' drag-joint accumulators, the motion-bound table of the separation cache and the
 *     sub-step's solver contacts (sharing their words with the broadphase group boxes, which are dead
 *     by the time contacts are born);
 *   - registers: everything that belongs to the robot's joint chain (control / eye bodies, the ten
 *     joint accumulators, per-sub-step pin frames, biases, motor rates) -- the chain is the same in every
 *     MAGICAL scene (entities.py:238-354), so it is unrolled with static indexing;
 *   - the environment's record in HBM/L2 is touched at the start and the end of the launch, plus the
 *     small contact cache once per sub-step.
 *
 * The arithmetic of every formula is the literal operation order of mg_physics.cu / oracle/mgo_physics.c
 * (bit-exact parity); only the *schedule* differs.  The Gauss-Seidel order is the reference's: contacts in
 * canonical arbiter order, then joints; joints of different bodies' drag pairs and the robot chain share
 * no dynamic body, so their relative order is immaterial (they commute exactly).
 *
 * The file compiles for the host as well (tests/host/tpe_host.cpp), which is how the CPU test-suite
 * checks this very source against the oracle without a GPU.
 */
#ifndef MG_PHYSICS_TPE_H
#define MG_PHYSICS_TPE_H

#include "mg_device.cuh"
#include "mg_narrowphase.h"
#include "mg_sincos.h"

#define TPE_CON_WORDS 14
#define TPE_MAX_CONTACTS 32 /* solver contacts per sub-step (private words first, then the spill area) */
#define TPE_NO_SLOT 15

#if defined(__CUDACC__)
#define MG_HDM __host__ __device__ __forceinline__
#define TPE_NOINLINE static __host__ __device__ __noinline__
#else
#define MG_HDM inline
#define TPE_NOINLINE static
#endif
#if defined(__CUDA_ARCH__)
#define TPE_LDG(p) __ldg(p)
#define TPE_CTZ(x) (__ffs((int)(x)) - 1)
#else
#define TPE_LDG(p) (*(p))
#define TPE_CTZ(x) __builtin_ctz(x)
#endif

/* bits [cell(lo), cell(hi)] of a 16-cell axis over [-1.28, 1.28); cell() is monotone and clamps */
MG_HD uint32_t tpe_cell_mask(float lo, float hi) {
  const float a = fminf(fmaxf((lo + 1.28f) * 6.25f, 0.0f), 15.0f);
  const float b = fminf(fmaxf((hi + 1.28f) * 6.25f, 0.0f), 15.0f);
  const int ia = (int)a, ib = (int)b; /* truncation of a non-negative value: monotone */
  return ((2u << ib) - 1u) & ~((1u << ia) - 1u);
}

MG_HD float tpe_d2f_ru(double x) {
#if defined(__CUDA_ARCH__)
  return __double2float_ru(x);
#else
  float f = (float)x;
  if ((double)f < x) f = nextafterf(f, INFINITY);
  return f;
#endif
}
MG_HD float tpe_d2f_rd(double x) {
#if defined(__CUDA_ARCH__)
  return __double2float_rd(x);
#else
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -INFINITY);
  return f;
#endif
}
MG_HD float tpe_fadd_ru(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_ru(a, b);
#else
  return tpe_d2f_ru((double)a + (double)b); /* the double sum of two floats is exact */
#endif
}
MG_HD float tpe_fadd_rd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rd(a, b);
#else
  return tpe_d2f_rd((double)a + (double)b);
#endif
}

/* Private-word layout, uniform for a launch (sized for the largest scene of the handle). */
struct TpeLayout {
  int nslots;   /* velocity slots incl. the static dummy */
  int nblocks;
  int off_bv, off_pr, off_bj, off_path, off_con;
  int kcon;     /* contacts that fit in the private words */
  int off_it, nitems; /* narrowphase work items of the sub-step (32 bits each) */
  int off_sep;        /* separation cache: one 16-bit truncated float per candidate group pair */
  int scratch_global; /* 1: items + separation cache live in a per-environment scratch record in HBM/L2
                         (touched a few times per sub-step) instead of the private words */
  int scratch_u32;    /* 32-bit words of one environment's scratch record */
  int words;    /* 8-byte words per environment */
};

static inline TpeLayout tpe_make_layout(int nslots, int nblocks, int ncgroups, int nbpairs, int kcon, int nitems,
                                        int scratch_global) {
  TpeLayout L;
  L.nslots = nslots;
  L.nblocks = nblocks;
  L.off_bv = nslots * 3;
  L.off_pr = L.off_bv + nslots * 3;
  L.off_bj = L.off_pr + (nslots - 1) * 5;
  L.off_path = L.off_bj + nblocks * 3;
  L.off_con = L.off_path + (nslots + 1) / 2;
  int con_words = kcon * TPE_CON_WORDS;
  if (con_words < ncgroups * 3) con_words = ncgroups * 3; /* group boxes + info share the contact words */
  L.kcon = con_words / TPE_CON_WORDS;
  if (L.kcon > TPE_MAX_CONTACTS) L.kcon = TPE_MAX_CONTACTS;
  L.off_it = L.off_con + con_words;
  L.nitems = nitems < 2 ? 2 : (nitems > 64 ? 64 : (nitems + 1) / 2 * 2); /* item indices are 6 bits */
  L.off_sep = L.off_it + L.nitems / 2;
  L.words = L.off_sep + (nbpairs + 3) / 4;
  L.scratch_global = scratch_global;
  L.scratch_u32 = (L.nitems + (nbpairs + 1) / 2 + 3) / 4 * 4;
  if (scratch_global) L.words = L.off_it;
  return L;
}

/* accessors over one environment's private words; S = distance (in words) between consecutive words */
template <int S>
struct Tpe {
  double* wd;   /* the three views of the private words, each already offset by the lane */
  float* wf;
  uint16_t* wh;
  uint32_t* itp;  /* work items: item k at itp[k * it_st]; the next lane's array starts it_lane elements on */
  uint16_t* sepp; /* separation cache, same stride */
  int it_st, it_lane;
  TpeLayout L;
  double* spill; /* this lane's column of the block's [TPE_MAX_CONTACTS - kcon][TPE_CON_WORDS][S] spill record, or null */
  uint64_t slotmap;
  int static_slot;
  MG_HDM double& V(int s, int k) const { return wd[(s * 3 + k) * S]; }
  MG_HDM double& Bv(int s, int k) const { return wd[(L.off_bv + s * 3 + k) * S]; }
  MG_HDM double& PR(int s, int k) const { return wd[(L.off_pr + s * 5 + k) * S]; } /* x y angle cos sin */
  MG_HDM double& BJ(int b, int k) const { return wd[(L.off_bj + b * 3 + k) * S]; } /* pivot x,y  gear */
  MG_HDM float& path(int s) const { return wf[(L.off_path * 2 + s) * S]; }
  /* group box l, b, r, t (k = 0..3) and, at k = 4, shape0 | nshape << 8 | slot << 16 of the group */
  MG_HDM float& gbb(int g, int k) const { return wf[(L.off_con * 2 + g * 6 + k) * S]; }
  MG_HDM uint32_t& ginfo(int g) const { return reinterpret_cast<uint32_t*>(wf)[(L.off_con * 2 + g * 6 + 4) * S]; }
  MG_HDM uint32_t& gcell(int g) const { return reinterpret_cast<uint32_t*>(wf)[(L.off_con * 2 + g * 6 + 5) * S]; }
  MG_HDM uint32_t& IT(int k) const { return itp[k * it_st]; }
  MG_HDM uint16_t& SEP(int p) const { return sepp[p * it_st]; }
  /* bind the item / separation-cache views: private words, or `scratch` = this lane's record in HBM */
  MG_HDM void bind_scratch(uint32_t* scratch) {
    if (L.scratch_global) {
      itp = scratch; sepp = reinterpret_cast<uint16_t*>(scratch + L.nitems); it_st = 1; it_lane = L.scratch_u32;
    } else {
      itp = reinterpret_cast<uint32_t*>(wf) + (size_t)L.off_it * 2 * S;
      sepp = wh + (size_t)L.off_sep * 4 * S;
      it_st = S; it_lane = 1;
    }
  }
  MG_HDM int slot(int body) const { /* body index or <0 / MG_MAX_BODIES for the static body */
    return (body < 0 || body >= MG_MAX_BODIES) ? static_slot : (int)((slotmap >> (4 * body)) & 15u);
  }
};

/* reference to the words of contact c (private words or spill area).  The spill area uses the same
 * [word][lane] interleaving as the private words (one record per 32-environment block), so word k of a contact
 * is always S elements after word k - 1 and the accesses compile to immediate offsets. */
template <int S>
struct TpeCon {
  double* p;
  MG_HDM double& operator[](int k) const { return p[k * S]; }
};
template <int S>
MG_HD TpeCon<S> tpe_con(const Tpe<S>& T, int c) {
  TpeCon<S> r;
  if (c < T.L.kcon) r.p = &T.wd[(T.L.off_con + c * TPE_CON_WORDS) * S];
  else r.p = T.spill + (size_t)((c - T.L.kcon) * TPE_CON_WORDS) * S;
  return r;
}
/* contact words: 0,1 r1 | 2,3 r2 | 4,5 n | 6 nMass | 7 tMass | 8 bias | 9 jn | 10 jt | 11 jb | 12 u | 13 ids */
MG_HD double tpe_pack_ids(unsigned hash, int sa, int sb, int ba, int bb, int first) {
  unsigned long long v = (unsigned long long)hash | ((unsigned long long)(sa & 0xFF) << 32) |
                         ((unsigned long long)(sb & 0xFF) << 40) | ((unsigned long long)(ba & 0x1F) << 48) |
                         ((unsigned long long)(bb & 0x1F) << 53) | ((unsigned long long)(first & 1) << 58);
  double d;
  memcpy(&d, &v, 8);
  return d;
}
MG_HD unsigned long long tpe_unpack_ids(double d) {
  unsigned long long v;
  memcpy(&v, &d, 8);
  return v;
}


#if defined(TPE_STATS) && !defined(__CUDACC__)
extern long tpe_stats[8]; /* host instrumentation: 0 sub-steps, 1 items, 2 items with contacts, 3 box-overlapping group pairs, 4 sep-skipped,
                             5 blocks, 6 blocks walked by the solver */
#define TPE_STAT(i) (tpe_stats[i]++)
#else
#define TPE_STAT(i) ((void)0)
#endif
#define TPE_IT_LAST (1u << 24)
#define TPE_IT_CONT (1u << 25)
#define TPE_NL(S) ((S) > 1 ? 32 : 1) /* lanes that cooperate: the warp on the device, one lane on the host */

MG_HD bool tpe_f4_overlap(float4 a, float4 b) { return a.x <= b.z && b.x <= a.z && a.y <= b.w && b.y <= a.w; }
/* gap between two disjoint boxes (exact in double): a lower bound of the distance of their contents */
MG_HD double tpe_f4_gap(float4 a, float4 b) {
  return dmaxf(dmaxf((double)b.x - (double)a.z, (double)a.x - (double)b.z),
               dmaxf((double)b.y - (double)a.w, (double)a.y - (double)b.w));
}

/* Separation cache entry of candidate pair p: the value of path[a] + path[b] at which the measured
 * separation expires, kept as the upper 16 bits of the float (truncation rounds a positive limit DOWN,
 * i.e. towards re-measuring early); 0 = no valid measurement. */
template <int S> struct Tpe;
template <int S> MG_HD float tpe_sep_get(const Tpe<S>& T, int p) {
  const uint32_t bits = (uint32_t)T.SEP(p) << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
template <int S> MG_HD void tpe_sep_set(const Tpe<S>& T, int p, float limit) {
  uint32_t bits;
  memcpy(&bits, &limit, 4);
  T.SEP(p) = (limit > 0.0f) ? (uint16_t)(bits >> 16) : (uint16_t)0;
}

#ifndef TPE_MAX_SURV
#define TPE_MAX_SURV 10 /* survivors per environment and sub-step that go through the cooperative GJK stage */
#endif

/* 6-bit "level" of a positive separation margin: the largest value 2^e * {1, 1.25, 1.5} (e >= -20) not
 * above it, so that level -> value is a LOWER bound (exact in binary).  Level 1 = 2^-20, below the
 * caching threshold; level 0 is reserved for "needs / produced a narrowphase result". */
MG_HD int tpe_level(double m) {
  if (!(m > 9.5367431640625e-07)) return 1; /* 2^-20 */
  unsigned long long bits;
  memcpy(&bits, &m, 8);
  const int e = (int)((bits >> 52) & 0x7FFull) - 1023;
  const int top2 = (int)((bits >> 50) & 3ull); /* mantissa >= .25 / .5 / .75 */
  const int k = top2 >= 2 ? 2 : top2;
  const int level = 3 * (e + 20) + k + 1;
  return level > 63 ? 63 : level;
}
MG_HD double tpe_level_value(int level) {
  const int L = level - 1;
  const int e = L / 3 - 20, k = L % 3;
  const unsigned long long bits = ((unsigned long long)(e + 1023) << 52) | ((unsigned long long)k << 50);
  double v;
  memcpy(&v, &bits, 8);
  return v;
}

/* warp cooperation primitives; on the host (S == 1) the warp is a single lane */
template <int S> MG_HD int tpe_lane() {
#if defined(__CUDA_ARCH__)
  return S > 1 ? (int)(threadIdx.x & 31) : 0;
#else
  return 0;
#endif
}
template <int S> MG_HD int tpe_shfl(int v, int src) {
#if defined(__CUDA_ARCH__)
  return S > 1 ? __shfl_sync(0xffffffffu, v, src) : v;
#else
  (void)src; return v;
#endif
}
template <int S> MG_HD double tpe_shfld(double v, int src) {
#if defined(__CUDA_ARCH__)
  return S > 1 ? __shfl_sync(0xffffffffu, v, src) : v;
#else
  (void)src; return v;
#endif
}
template <int S> MG_HD uint64_t tpe_shfl64(uint64_t v, int src) {
#if defined(__CUDA_ARCH__)
  return S > 1 ? (uint64_t)__shfl_sync(0xffffffffu, (unsigned long long)v, src) : v;
#else
  (void)src; return v;
#endif
}
template <int S> MG_HD int tpe_scan_incl(int v) {
#if defined(__CUDA_ARCH__)
  if (S > 1) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
  }
#endif
  return v;
}
template <int S> MG_HD int tpe_max(int v) {
#if defined(__CUDA_ARCH__)
  if (S > 1) return __reduce_max_sync(0xffffffffu, v);
#endif
  return v;
}
template <int S> MG_HD void tpe_sync() {
#if defined(__CUDA_ARCH__)
  if (S > 1) __syncwarp();
#endif
}
/* owner of queue entry j, given every lane's exclusive prefix of its entry count: the last lane whose
 * prefix is <= j (all lanes must call this together) */
template <int S> MG_HD int tpe_find_owner(int excl, int j) {
  int lo = 0, hi = TPE_NL(S) - 1;
#pragma unroll
  for (int step = 0; step < 5; step++) {
    const int mid = (lo + hi + 1) >> 1;
    const int e = tpe_shfl<S>(excl, mid);
    if (lo < hi) { if (e <= j) lo = mid; else hi = mid - 1; }
  }
  return lo;
}

template <int S>
MG_HD ShapeView tpe_view(const Tpe<S>& T, const DeviceScene* ds, int si) {
  const mg_shape_t& sh = ds->s.shapes[si];
  ShapeView v;
  v.kind = sh.kind;
  v.nvert = sh.nvert;
  v.lv = &ds->s.cverts[sh.vert0][0];
  v.ln = &ds->aux.cnorm[sh.vert0][0];
  v.radius = sh.radius;
  v.index = si;
  const int b = sh.body;
  if (b >= 0) {
    const int s = T.slot(b);
    v.px = T.PR(s, 0); v.py = T.PR(s, 1); v.rc = T.PR(s, 3); v.rs = T.PR(s, 4);
  } else {
    v.rc = 1.0; v.rs = 0.0; v.px = 0.0; v.py = 0.0;
  }
  return v;
}

/* conservative fp32 box of one shape (outward rounding of the exact fp64 box) */
template <int S>
MG_HD float4 tpe_shape_box(const Tpe<S>& T, const DeviceScene* ds, int si) {
  if (ds->s.shapes[si].body >= 0) {
    ShapeView v = tpe_view(T, ds, si);
    double bb[4];
    sv_bb(v, bb);
    return make_float4(tpe_d2f_rd(bb[0]), tpe_d2f_rd(bb[1]), tpe_d2f_ru(bb[2]), tpe_d2f_ru(bb[3]));
  }
  return make_float4(ds->aux.static_bb[si][0], ds->aux.static_bb[si][1], ds->aux.static_bb[si][2],
                     ds->aux.static_bb[si][3]);
}

/* gap between the exact boxes of two shapes: > 0 = disjoint (a lower bound of the shapes' distance) */
template <int S>
MG_HD double tpe_pair_gap(const Tpe<S>& T, const DeviceScene* ds, int ia, int ib) {
  ShapeView va = tpe_view(T, ds, ia), vb = tpe_view(T, ds, ib);
  double bba[4], bbb[4];
  sv_bb(va, bba);
  sv_bb(vb, bbb);
  if (bb_intersects(bba, bbb)) return -1.0;
  return dmaxf(dmaxf(bbb[0] - bba[2], bba[0] - bbb[2]), dmaxf(bbb[1] - bba[3], bba[1] - bbb[3]));
}

/* Exact narrowphase of one shape pair (kinds already ordered).  When the exact boxes are disjoint the
 * gap between them is a valid lower bound of the shapes' distance and is reported as the margin. */
template <int S>
TPE_NOINLINE void tpe_narrow_pair(const Tpe<S>* Tp, const DeviceScene* ds, int ia, int ib, Manifold* out) {
  const Tpe<S>& T = *Tp;
  ShapeView va = tpe_view(T, ds, ia), vb = tpe_view(T, ds, ib);
  double bba[4], bbb[4];
  sv_bb(va, bba);
  sv_bb(vb, bbb);
  Manifold m;
  m.count = 0;
  m.margin = -1.0;
  m.n = D2(0, 0);
  if (bb_intersects(bba, bbb)) {
    mg_collide(va, vb, bba, bbb, m);
  } else {
    m.margin = dmaxf(dmaxf(bbb[0] - bba[2], bba[0] - bbb[2]), dmaxf(bbb[1] - bba[3], bba[1] - bbb[3]));
  }
  *out = m;
}

/* v += j*m_inv; w += i_inv * cross(r, j) on a velocity slot (static dummy: zero inverse mass) */
#define TPE_APPLY(ARR, s, m_inv, i_inv, jx, jy, rx, ry)           \
  do {                                                            \
    double tx_ = T.ARR(s, 0), ty_ = T.ARR(s, 1), tz_ = T.ARR(s, 2); \
    tx_ = tx_ + (jx) * (m_inv);                                   \
    ty_ = ty_ + (jy) * (m_inv);                                   \
    tz_ += (i_inv) * ((rx) * (jy) - (ry) * (jx));                 \
    T.ARR(s, 0) = tx_; T.ARR(s, 1) = ty_; T.ARR(s, 2) = tz_;      \
  } while (0)

struct TpeJC {
  double ma, ia, mb, ib, c0, c1, c2, c3;
};
MG_HD TpeJC tpe_jc(const DeviceScene* ds, int j) {
  const double* p = &ds->aux.jc[j][0];
  TpeJC r;
  r.ma = TPE_LDG(p + 0); r.ia = TPE_LDG(p + 1); r.mb = TPE_LDG(p + 2); r.ib = TPE_LDG(p + 3);
  r.c0 = TPE_LDG(p + 4); r.c1 = TPE_LDG(p + 5); r.c2 = TPE_LDG(p + 6); r.c3 = TPE_LDG(p + 7);
  return r;
}

/* the three purely angular joint kinds of the chain on explicit angular velocities */
MG_HD void tpe_gear(const TpeJC& c, double bias, double& acc, double& wa, double& wb) {
  double wr = wb * c.c2 - wa;
  double jj = (bias - wr) * c.c0;
  const double jOld = acc;
  const double jNew = dclamp(jOld + jj, -c.c1, c.c1);
  acc = jNew;
  jj = jNew - jOld;
  double da = jj * c.ia;
  da = da * c.c3;
  wa = wa - da;
  wb = wb + jj * c.ib;
}
MG_HD void tpe_limit(const TpeJC& c, double bias, double& acc, double& wa, double& wb) {
  if (!bias) return;
  double wr = wb - wa;
  double jj = -(bias + wr) * c.c0;
  const double lo = (bias < 0.0) ? 0.0 : -c.c1;
  const double hi = (!(bias < 0.0)) ? 0.0 : c.c1;
  const double jOld = acc;
  const double jNew = dclamp(jOld + jj, lo, hi);
  acc = jNew;
  jj = jNew - jOld;
  double da = jj * c.ia;
  wa = wa - da;
  wb = wb + jj * c.ib;
}
MG_HD void tpe_motor(const TpeJC& c, double rate, double& acc, double& wa, double& wb) {
  double wr = wb - wa;
  wr = wr + rate;
  double jj = -wr * c.c0;
  const double jOld = acc;
  const double jNew = dclamp(jOld + jj, -c.c1, c.c1);
  acc = jNew;
  jj = jNew - jOld;
  double da = jj * c.ia;
  wa = wa - da;
  wb = wb + jj * c.ib;
}
MG_HD void tpe_spring(const TpeJC& c, double& target, double& acc, double& wa, double& wb) {
  double wrn = wa - wb;
  double w_damp = (target - wrn) * c.c2;
  target = wrn + w_damp;
  double j_damp = w_damp * c.c0;
  acc += j_damp;
  wa = wa + j_damp * c.ia;
  wb = wb - j_damp * c.ib;
}
struct TpePin {
  double r1x, r1y, r2x, r2y, nx, ny, nMass, bias;
};
MG_HD void tpe_pin(const TpeJC& c, const TpePin& pn, double& acc, double& vax, double& vay, double& wa, double& vbx,
                   double& vby, double& wb) {
  d2 r1 = D2(pn.r1x, pn.r1y), r2 = D2(pn.r2x, pn.r2y);
  d2 n = D2(pn.nx, pn.ny);
  d2 v1 = dadd(D2(vax, vay), dmul(dperp(r1), wa));
  d2 v2 = dadd(D2(vbx, vby), dmul(dperp(r2), wb));
  double vrn = ddot(dsub(v2, v1), n);
  double jn = (pn.bias - vrn) * pn.nMass;
  double jnOld = acc;
  double jnNew = dclamp(jnOld + jn, -c.c1, c.c1);
  acc = jnNew;
  jn = jnNew - jnOld;
  double jx = n.x * jn, jy = n.y * jn;
  vax = vax + (-jx) * c.ma; vay = vay + (-jy) * c.ma;
  wa += c.ia * (r1.x * (-jy) - r1.y * (-jx));
  vbx = vbx + jx * c.mb; vby = vby + jy * c.mb;
  wb += c.ib * (r2.x * jy - r2.y * jx);
}

/* One env-step of one environment. */
template <int S>
MG_HD void tpe_env_step(Tpe<S> T, EnvState* __restrict__ G, const DeviceScene* __restrict__ ds, int action,
                         const bool live) {
  const mg_scene_t& sc = ds->s;
  const mg_scene_aux_t& ax = ds->aux;
  const double dt = MG_DT;
  /* lanes beyond the batch (`live` == false) stay in the warp for the cooperative narrowphase: they read a
   * valid record, do no per-environment work (all trip counts zero) and never store to global memory */
  const int nblk = live ? ax.tpe_nblocks : 0;
  const int nslots = live ? ax.tpe_nslots : 1;
  T.slotmap = ax.tpe_slotmap;
  T.static_slot = nslots - 1;
  const int robot = sc.robot_body, control = sc.control_body;
  const int eye0 = sc.eye_body[0], eye1 = sc.eye_body[1];
  const int fb0 = sc.finger_body[0], fb1 = sc.finger_body[1];
  const int s_robot = T.slot(robot), s_f0 = T.slot(fb0), s_f1 = T.slot(fb1);
  const int jr0 = ax.tpe_jr0;
  const int nbp = live ? sc.n_bpairs : 0, ncg = live ? sc.n_cgroups : 0;

  int stamp = G->stamp, n_cache = live ? G->n_cache : 0, overflow = G->overflow;

  /* ---- load: shaped bodies into the private words, chain bodies / accumulators into registers */
  for (int s = 0; s < nslots - 1; s++) {
    const int b = ax.tpe_slot_body[s];
    const double4 v = G->V[b], bv = G->Bv[b], p = G->P[b];
    const double2 r = G->R[b];
    T.V(s, 0) = v.x; T.V(s, 1) = v.y; T.V(s, 2) = v.z;
    T.Bv(s, 0) = bv.x; T.Bv(s, 1) = bv.y; T.Bv(s, 2) = bv.z;
    T.PR(s, 0) = p.x; T.PR(s, 1) = p.y; T.PR(s, 2) = p.z; T.PR(s, 3) = r.x; T.PR(s, 4) = r.y;
    T.path(s) = 0.0f;
  }
  {
    const int s = nslots - 1;
    T.V(s, 0) = 0.0; T.V(s, 1) = 0.0; T.V(s, 2) = 0.0;
    T.Bv(s, 0) = 0.0; T.Bv(s, 1) = 0.0; T.Bv(s, 2) = 0.0;
    T.path(s) = 0.0f;
  }
  for (int k = 0; k < nblk; k++) {
    const double2 ap = G->jacc[ax.tpe_bj_pivot[k]];
    T.BJ(k, 0) = ap.x; T.BJ(k, 1) = ap.y;
    T.BJ(k, 2) = G->jacc[ax.tpe_bj_gear[k]].x;
  }
  for (int p = 0; p < nbp; p++) T.SEP(p) = 0;
  /* control body (kinematic) and the two eye bodies (no shapes) */
  double4 Pc = G->P[control], Vc = G->V[control], Bc = G->Bv[control];
  double4 Pe0 = G->P[eye0], Ve0 = G->V[eye0], Be0 = G->Bv[eye0];
  double4 Pe1 = G->P[eye1], Ve1 = G->V[eye1], Be1 = G->Bv[eye1];
  double2 Rc = G->R[control], Re0 = G->R[eye0], Re1 = G->R[eye1];
  /* chain accumulators */
  double a_pivx, a_pivy, a_gear, a_spr0, a_spr1, a_pin0, a_lim0, a_mot0, a_pin1, a_lim1, a_mot1;
  { double2 t = G->jacc[jr0]; a_pivx = t.x; a_pivy = t.y; }
  a_gear = G->jacc[jr0 + 1].x; a_spr0 = G->jacc[jr0 + 2].x; a_spr1 = G->jacc[jr0 + 3].x;
  a_pin0 = G->jacc[jr0 + 4].x; a_lim0 = G->jacc[jr0 + 5].x; a_mot0 = G->jacc[jr0 + 6].x;
  a_pin1 = G->jacc[jr0 + 7].x; a_lim1 = G->jacc[jr0 + 8].x; a_mot1 = G->jacc[jr0 + 9].x;

  /* ---- Robot.set_action: id = 9*grip + 3*lr + ud (entities.py:162-186, 439-457) */
  action = action < 0 ? 0 : (action > 17 ? 17 : action);
  const int ud = action % 3, lr = (action / 3) % 3, grip = action / 9;
  const double Rr = sc.robot_radius;
  double target_speed = 0.0, rel_turn = 0.0;
  if (ud == 1) target_speed += 4.0 * Rr;
  if (ud == 2) target_speed -= 3.0 * Rr;
  if (lr == 1) rel_turn += 1.5;
  if (lr == 2) rel_turn -= 1.5;
  const double target_finger = (grip == 0) ? (3.14159265358979323846 / 8) : -0.0;
  int ncon = 0;

  for (int sub = 0; sub < MG_SUBSTEPS; ++sub) {
    stamp++;
    TPE_STAT(0);
    /* warp fence: the previous sub-step's cooperative stages read other lanes' pose words, and the group boxes
     * written below share their words with other lanes' solver contacts of the previous sub-step */
    tpe_sync<S>();
    /* ---- Robot.update (entities.py:459-479) */
    double rate0, rate1;
    {
      Pc.z = T.PR(s_robot, 2) + rel_turn;
      const double rc = T.PR(s_robot, 3), rs = T.PR(s_robot, 4);
      Vc.x = rc * 0.0 - rs * target_speed;
      Vc.y = rc * target_speed + rs * 0.0;
      const double ra = T.PR(s_robot, 2);
      {
        double angle_error = (T.PR(s_f0, 2) - ra) + (-1.0) * target_finger;
        double r = dmaxf(-1, dminf(1, angle_error * 10));
        if (fabs(r) < 1e-4) r = 0.0;
        rate0 = r;
      }
      {
        double angle_error = (T.PR(s_f1, 2) - ra) + (1.0) * target_finger;
        double r = dmaxf(-1, dminf(1, angle_error * 10));
        if (fabs(r) < 1e-4) r = 0.0;
        rate1 = r;
      }
    }

    /* ---- integrate positions (cpBodyUpdatePosition; kinematic control body included) */
    for (int s = 0; s < nslots - 1; s++) {
      const double vx = T.V(s, 0) + T.Bv(s, 0), vy = T.V(s, 1) + T.Bv(s, 1), vw = T.V(s, 2) + T.Bv(s, 2);
      const double x = T.PR(s, 0) + vx * dt;
      const double y = T.PR(s, 1) + vy * dt;
      const double a = T.PR(s, 2) + vw * dt;
      double sn, cs;
      mg_det_sincos(a, &sn, &cs);
      T.PR(s, 0) = x; T.PR(s, 1) = y; T.PR(s, 2) = a; T.PR(s, 3) = cs; T.PR(s, 4) = sn;
      T.Bv(s, 0) = 0.0; T.Bv(s, 1) = 0.0; T.Bv(s, 2) = 0.0;
      /* how far can any point of this body's shapes have moved: |dp|_1 + reach * |dtheta|, rounded up */
      const double moved = (fabs(vx) + fabs(vy) + TPE_LDG(&ax.body_reach[ax.tpe_slot_body[s]]) * fabs(vw)) * dt;
      T.path(s) = tpe_fadd_ru(T.path(s), tpe_d2f_ru(moved * 1.000001));
    }
#define TPE_INTEGRATE_REG(P, V, B, R)                   \
  do {                                                  \
    P.x = P.x + (V.x + B.x) * dt;                       \
    P.y = P.y + (V.y + B.y) * dt;                       \
    P.z = P.z + (V.z + B.z) * dt;                       \
    double sn_, cs_;                                    \
    mg_det_sincos(P.z, &sn_, &cs_);                     \
    R = make_double2(cs_, sn_);                         \
    B = make_double4(0.0, 0.0, 0.0, 0.0);               \
  } while (0)
    TPE_INTEGRATE_REG(Pc, Vc, Bc, Rc);
    TPE_INTEGRATE_REG(Pe0, Ve0, Be0, Re0);
    TPE_INTEGRATE_REG(Pe1, Ve1, Be1, Re1);

    /* ---- conservative fp32 boxes of the collision groups: body position +- reach (the largest distance of
     * any point of the body's shapes from its origin), which needs no vertex transforms; the exact boxes
     * are only computed for the few pairs that get past this filter and the separation cache */
    for (int g = 0; g < ncg; g++) {
      const int gbody = sc.cgroups[g].body;
      float l, b, r, t;
      if (gbody >= 0) {
        const int s = T.slot(gbody);
        const double reach = TPE_LDG(&ax.body_reach[gbody]);
        const double x = T.PR(s, 0), y = T.PR(s, 1);
        l = tpe_d2f_rd(x - reach); b = tpe_d2f_rd(y - reach); r = tpe_d2f_ru(x + reach); t = tpe_d2f_ru(y + reach);
      } else {
        const int s0 = sc.cgroups[g].shape0, n = sc.cgroups[g].nshape;
        l = INFINITY; b = INFINITY; r = -INFINITY; t = -INFINITY;
        for (int k = 0; k < n; k++) {
          const float4 bx = tpe_shape_box(T, ds, s0 + k);
          l = fminf(l, bx.x); b = fminf(b, bx.y); r = fmaxf(r, bx.z); t = fmaxf(t, bx.w);
        }
      }
      T.gbb(g, 0) = l; T.gbb(g, 1) = b; T.gbb(g, 2) = r; T.gbb(g, 3) = t;
      T.ginfo(g) = (uint32_t)sc.cgroups[g].shape0 | ((uint32_t)sc.cgroups[g].nshape << 8) |
                   ((uint32_t)T.slot(gbody) << 16);
      /* the 16 x 16 grid cells the box touches, as two 16-bit masks (x low, y high): boxes that overlap
       * share a cell in both axes (the cell index is monotone in the coordinate), so one AND rejects most
       * of the pair list before any float compare */
      T.gcell(g) = tpe_cell_mask(l, r) | (tpe_cell_mask(b, t) << 16);
    }

    /* ---- broadphase (own environment): canonical pair list -> work items for the exact narrowphase.
     * Only trivial per-lane work happens here (lanes diverge): group boxes, the cached separation of the
     * group pair, then one 32-bit item per shape pair.  The exact per-shape boxes are tested by the lanes
     * that execute the items, converged, inside tpe_narrow_pair. */
    int n_items = 0;
    bool truncated = false;
    int res_p = 0, res_i = 0, res_k = 0;
    /* pass 1 (all lanes in step): box test of every candidate pair -> bit mask in registers */
    uint32_t pmask[(MG_MAX_BPAIRS + 31) / 32];
#pragma unroll
    for (int w = 0; w < (MG_MAX_BPAIRS + 31) / 32; w++) pmask[w] = 0u;
    for (int w0 = 0; w0 * 32 < nbp; w0++) {
      uint32_t m = 0u;
      const int pe = (nbp - w0 * 32 < 32) ? nbp - w0 * 32 : 32;
      for (int b = 0; b < pe; b++) {
        const int p = w0 * 32 + b;
        const unsigned pr = TPE_LDG(reinterpret_cast<const unsigned short*>(&sc.bpairs[p][0]));
        const int ga = (int)(pr & 0xFFu), gb = (int)(pr >> 8);
        const uint32_t shared = T.gcell(ga) & T.gcell(gb);
        const bool hit = (shared & 0xFFFFu) != 0u && (shared >> 16) != 0u;
        m |= (hit ? 1u : 0u) << b;
      }
#pragma unroll
      for (int w = 0; w < (MG_MAX_BPAIRS + 31) / 32; w++)
        if (w0 == w) pmask[w] = m;
    }
    /* pass 2 (lanes diverge, but only over their own few hits): separation cache, then the items */
#pragma unroll
    for (int w = 0; w < (MG_MAX_BPAIRS + 31) / 32; w++) {
      uint32_t bits = pmask[w];
      while (bits && !truncated) {
        const int p = w * 32 + TPE_CTZ(bits);
        bits &= bits - 1u;
        TPE_STAT(3);
        const int ga = sc.bpairs[p][0], gb = sc.bpairs[p][1];
        if (!(T.gbb(ga, 0) <= T.gbb(gb, 2) && T.gbb(gb, 0) <= T.gbb(ga, 2) &&
              T.gbb(ga, 1) <= T.gbb(gb, 3) && T.gbb(gb, 1) <= T.gbb(ga, 3))) continue;
        const uint32_t ia_ = T.ginfo(ga), ib_ = T.ginfo(gb);
        const float travelled = tpe_fadd_ru(T.path((int)(ia_ >> 16)), T.path((int)(ib_ >> 16)));
        if (travelled < tpe_sep_get(T, p)) { TPE_STAT(4); continue; }
        const int sa0 = (int)(ia_ & 0xFFu), na = (int)((ia_ >> 8) & 0xFFu);
        const int sb0 = (int)(ib_ & 0xFFu), nbs = (int)((ib_ >> 8) & 0xFFu);
        const int first_item = n_items;
        for (int i = 0; i < na && !truncated; i++)
          for (int k = 0; k < nbs; k++) {
            if (n_items == T.L.nitems) { truncated = true; res_p = p; res_i = i; res_k = k; break; }
            /* (the executing lane orders the two shapes by kind, as cpCollide does, in stage A) */
            T.IT(n_items++) = (uint32_t)((sa0 + i) | ((sb0 + k) << 8) | (p << 16));
            TPE_STAT(1);
          }
        /* LAST = settle the pair's separation after this item; CONT = the pair continues in the serial tail */
        if (n_items > first_item) T.IT(n_items - 1) |= truncated ? TPE_IT_CONT : TPE_IT_LAST;
      }
    }

    /* ---- narrowphase + contact cache lookup (cpCollide + cpArbiterUpdate).
     * The items of all the warp's environments form one queue; every lane runs the exact narrowphase of
     * one item (of whichever environment) on the owner's private words, and the owners then pull their
     * results, in canonical order, with shuffles.  On the host build the "warp" is one lane. */
    ncon = 0;
    bool too_many = false;
    uint64_t cache_used = 0ull;
    const int kcap = T.spill ? TPE_MAX_CONTACTS : T.L.kcon;
    /* turn one pair's manifold into solver contacts of this environment */
    auto take_manifold = [&](int ia, int ib, const Manifold& m) {
      if (m.count == 0) return;
      TPE_STAT(2);
      if (ncon + m.count > kcap) { too_many = true; ncon += m.count; return; }
      int ba = sc.shapes[ia].body, bb = sc.shapes[ib].body;
      d2 pa = D2(0, 0), pb = D2(0, 0);
      if (ba >= 0) { const int s = T.slot(ba); pa = D2(T.PR(s, 0), T.PR(s, 1)); } else ba = MG_MAX_BODIES;
      if (bb >= 0) { const int s = T.slot(bb); pb = D2(T.PR(s, 0), T.PR(s, 1)); } else bb = MG_MAX_BODIES;
      /* the pair's cached contacts all stem from its last collision; they are warm-started only if
       * that was the previous sub-step (cpArbiterApplyCachedImpulse skips first-contact arbiters) */
      bool first = true;
      double jn[2] = {0.0, 0.0}, jt[2] = {0.0, 0.0};
      for (int q = 0; q < n_cache; q++) {
        const CEntry e = G->cache[q];
        if (e.a == ia && e.b == ib) {
          if (e.stamp == stamp - 1) first = false;
          if (m.hash[0] == e.hash) { jn[0] = e.jn; jt[0] = e.jt; }
          if (m.count > 1 && m.hash[1] == e.hash) { jn[1] = e.jn; jt[1] = e.jt; }
          cache_used |= 1ull << q;
        }
      }
      const double u = sc.shapes[ia].friction * sc.shapes[ib].friction;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q < m.count) {
          TpeCon<S> C = tpe_con(T, ncon + q);
          d2 r1 = dsub(m.p1[q], pa), r2 = dsub(m.p2[q], pb);
          C[0] = r1.x; C[1] = r1.y; C[2] = r2.x; C[3] = r2.y; C[4] = m.n.x; C[5] = m.n.y;
          C[9] = jn[q]; C[10] = jt[q]; C[12] = u;
          C[13] = tpe_pack_ids(m.hash[q], ia, ib, ba, bb, first ? 1 : 0);
        }
      }
      ncon += m.count;
    };
    /* shapes `margin` apart cannot touch until the bodies have travelled that far */
    auto settle_sep = [&](int p, double gm) {
      const int sla = T.slot(sc.cgroups[sc.bpairs[p][0]].body), slb = T.slot(sc.cgroups[sc.bpairs[p][1]].body);
      const float travelled = tpe_fadd_ru(T.path(sla), T.path(slb));
      tpe_sep_set(T, p, (gm > 1e-6 && gm < MG_INF) ? tpe_fadd_rd(travelled, tpe_d2f_rd(gm * 0.999999 - 1e-9)) : -1.0f);
    };
    {
      const int lane = tpe_lane<S>();
      /* warp fence: every lane's poses and work items are written, and nobody reads a group box any more
       * (contacts born below reuse those words), before lanes start working on each other's environments */
      tpe_sync<S>();
      /* ---- stage A: exact boxes of every item, one item per lane.  A pair whose boxes are disjoint is
       * settled here: the gap is a lower bound of the shapes' distance and goes into the item's level. */
      {
        const int incl = tpe_scan_incl<S>(n_items);
        const int excl = incl - n_items;
        const int n_all = tpe_shfl<S>(incl, TPE_NL(S) - 1);
        for (int base = 0; base < n_all; base += TPE_NL(S)) {
          const int j = base + lane;
          const int owner = tpe_find_owner<S>(excl, j);
          const int oexcl = tpe_shfl<S>(excl, owner);
          Tpe<S> To = T;
          To.wd = T.wd - lane + owner; To.wf = T.wf - lane + owner; To.wh = T.wh - lane + owner;
          To.itp = T.itp + (owner - lane) * T.it_lane;
          To.slotmap = tpe_shfl64<S>(T.slotmap, owner);
          To.static_slot = tpe_shfl<S>(T.static_slot, owner);
          const DeviceScene* dso = (const DeviceScene*)tpe_shfl64<S>((uint64_t)ds, owner);
          if (j < n_all) {
            uint32_t it = To.IT(j - oexcl);
            int ia = (int)(it & 0xFFu), ib = (int)((it >> 8) & 0xFFu);
            if (dso->s.shapes[ia].kind > dso->s.shapes[ib].kind) { /* Chipmunk's type order */
              const int t = ia; ia = ib; ib = t;
              it = (it & 0xFFFF0000u) | (uint32_t)ia | ((uint32_t)ib << 8);
            }
            const double gap = tpe_pair_gap(To, dso, ia, ib);
            if (gap > 0.0) it |= (uint32_t)tpe_level(gap) << 26;
            To.IT(j - oexcl) = it;
          }
        }
        tpe_sync<S>();
      }
      /* ---- the items that need GJK ("survivors"), as a packed list of item indices (6 bits each) */
      uint64_t surv = 0;
      int n_surv = 0;
      for (int k = 0; k < n_items; k++)
        if ((T.IT(k) >> 26) == 0u) {
          if (n_surv < TPE_MAX_SURV) surv |= (uint64_t)k << (6 * n_surv);
          n_surv++;
        }
      const int n_coop = n_surv < TPE_MAX_SURV ? n_surv : TPE_MAX_SURV;
      /* ---- stage B: GJK / EPA / clipping of the survivors, one per lane; the owners then pull their
       * results, in canonical order, with shuffles */
      {
        const int incl = tpe_scan_incl<S>(n_coop);
        const int excl = incl - n_coop;
        const int n_all = tpe_shfl<S>(incl, TPE_NL(S) - 1);
        int consumed = 0;
        for (int base = 0; base < n_all; base += TPE_NL(S)) {
          const int j = base + lane;
          const int owner = tpe_find_owner<S>(excl, j);
          const int oexcl = tpe_shfl<S>(excl, owner);
          const uint64_t osurv = tpe_shfl64<S>(surv, owner);
          Tpe<S> To = T;
          To.wd = T.wd - lane + owner; To.wf = T.wf - lane + owner; To.wh = T.wh - lane + owner;
          To.itp = T.itp + (owner - lane) * T.it_lane;
          To.slotmap = tpe_shfl64<S>(T.slotmap, owner);
          To.static_slot = tpe_shfl<S>(T.static_slot, owner);
          const DeviceScene* dso = (const DeviceScene*)tpe_shfl64<S>((uint64_t)ds, owner);
          Manifold m;
          m.count = 0; m.margin = -1.0; m.n = D2(0, 0);
          m.p1[0] = m.p1[1] = m.p2[0] = m.p2[1] = D2(0, 0);
          m.hash[0] = m.hash[1] = 0u;
          if (j < n_all) {
            const uint32_t it = To.IT((int)((osurv >> (6 * (j - oexcl))) & 63u));
            tpe_narrow_pair<S>(&To, dso, (int)(it & 0xFFu), (int)((it >> 8) & 0xFFu), &m);
          }
          const int a = (excl > base ? excl : base) - base;
          const int b = (incl < base + TPE_NL(S) ? incl : base + TPE_NL(S)) - base;
          const int cnt = b > a ? b - a : 0;
          const int most = tpe_max<S>(cnt);
          for (int t = 0; t < most; t++) {
            const int src = t < cnt ? a + t : lane;
            Manifold r;
            r.count = tpe_shfl<S>(m.count, src);
            r.margin = tpe_shfld<S>(m.margin, src);
            r.n.x = tpe_shfld<S>(m.n.x, src); r.n.y = tpe_shfld<S>(m.n.y, src);
            r.p1[0].x = tpe_shfld<S>(m.p1[0].x, src); r.p1[0].y = tpe_shfld<S>(m.p1[0].y, src);
            r.p2[0].x = tpe_shfld<S>(m.p2[0].x, src); r.p2[0].y = tpe_shfld<S>(m.p2[0].y, src);
            r.p1[1].x = tpe_shfld<S>(m.p1[1].x, src); r.p1[1].y = tpe_shfld<S>(m.p1[1].y, src);
            r.p2[1].x = tpe_shfld<S>(m.p2[1].x, src); r.p2[1].y = tpe_shfld<S>(m.p2[1].y, src);
            r.hash[0] = (unsigned)tpe_shfl<S>((int)m.hash[0], src);
            r.hash[1] = (unsigned)tpe_shfl<S>((int)m.hash[1], src);
            if (t < cnt) {
              const int k = (int)((surv >> (6 * consumed)) & 63u);
              consumed++;
              const uint32_t it = T.IT(k);
              take_manifold((int)(it & 0xFFu), (int)((it >> 8) & 0xFFu), r);
              if (r.count == 0 && r.margin > 0.0) T.IT(k) = it | ((uint32_t)tpe_level(r.margin) << 26);
            }
          }
        }
      }
      tpe_sync<S>();
      /* survivors beyond the packed list (rare): by the owner itself, still in canonical order */
      if (n_surv > TPE_MAX_SURV) {
        const int last_coop = (int)((surv >> (6 * (TPE_MAX_SURV - 1))) & 63u);
        for (int k = last_coop + 1; k < n_items; k++) {
          const uint32_t it = T.IT(k);
          if ((it >> 26) != 0u) continue;
          Manifold m;
          tpe_narrow_pair<S>(&T, ds, (int)(it & 0xFFu), (int)((it >> 8) & 0xFFu), &m);
          take_manifold((int)(it & 0xFFu), (int)((it >> 8) & 0xFFu), m);
          if (m.count == 0 && m.margin > 0.0) T.IT(k) = it | ((uint32_t)tpe_level(m.margin) << 26);
        }
      }
    }
    /* ---- separation cache: per candidate pair, the smallest margin over its items (level 0 = touching
     * or unknown -> not cached) */
    double gm_run = MG_INF;
    for (int k = 0; k < n_items; k++) {
      const uint32_t it = T.IT(k);
      const uint32_t lvl = it >> 26;
      gm_run = dminf(gm_run, lvl ? tpe_level_value((int)lvl) : 0.0);
      if (it & TPE_IT_LAST) { settle_sep((int)((it >> 16) & 0xFFu), gm_run); gm_run = MG_INF; }
      /* CONT: gm_run is carried into the serial tail */
    }
    if (truncated) {
      /* serial tail (rare: more candidate pairs than item words): the remaining pairs in canonical order;
       * tpe_narrow_pair tests the exact boxes itself */
      for (int p = res_p; p < nbp; p++) {
        const int ga = sc.bpairs[p][0], gb = sc.bpairs[p][1];
        const int sa0 = sc.cgroups[ga].shape0, na = sc.cgroups[ga].nshape;
        const int sb0 = sc.cgroups[gb].shape0, nbs = sc.cgroups[gb].nshape;
        if (p != res_p) {
          const int sla = T.slot(sc.cgroups[ga].body), slb = T.slot(sc.cgroups[gb].body);
          if (tpe_fadd_ru(T.path(sla), T.path(slb)) < tpe_sep_get(T, p)) continue;
          gm_run = MG_INF;
        }
        for (int i = (p == res_p ? res_i : 0); i < na; i++)
          for (int k = (p == res_p && i == res_i ? res_k : 0); k < nbs; k++) {
            int ia = sa0 + i, ib = sb0 + k;
            if (sc.shapes[ia].kind > sc.shapes[ib].kind) { int t = ia; ia = ib; ib = t; }
            Manifold m;
            tpe_narrow_pair<S>(&T, ds, ia, ib, &m);
            take_manifold(ia, ib, m);
            gm_run = dminf(gm_run, (m.count == 0 && m.margin > 0.0) ? m.margin : 0.0);
          }
        settle_sep(p, gm_run);
      }
    }
    if (too_many) {
      /* capacity exceeded: flag the env and solve this sub-step without contacts rather than with a
       * partial, order-dependent subset */
      overflow |= 2;
      ncon = 0;
    }

    /* ---- contact prestep */
    for (int c = 0; c < ncon; c++) {
      TpeCon<S> C = tpe_con(T, c);
      const unsigned long long ids = tpe_unpack_ids(C[13]);
      const int ba = (int)((ids >> 48) & 0x1F), bb = (int)((ids >> 53) & 0x1F);
      const double ma = ba < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[ba].m_inv) : 0.0;
      const double ia_ = ba < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[ba].i_inv) : 0.0;
      const double mb = bb < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[bb].m_inv) : 0.0;
      const double ib_ = bb < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[bb].i_inv) : 0.0;
      d2 r1 = D2(C[0], C[1]), r2 = D2(C[2], C[3]), n = D2(C[4], C[5]), t = dperp(n);
      double rcn1 = dcross(r1, n), rcn2 = dcross(r2, n);
      C[6] = 1.0 / ((ma + ia_ * rcn1 * rcn1) + (mb + ib_ * rcn2 * rcn2));
      double rct1 = dcross(r1, t), rct2 = dcross(r2, t);
      C[7] = 1.0 / ((ma + ia_ * rct1 * rct1) + (mb + ib_ * rct2 * rct2));
      d2 pa = D2(0, 0), pb = D2(0, 0);
      if (ba < MG_MAX_BODIES) { const int s = T.slot(ba); pa = D2(T.PR(s, 0), T.PR(s, 1)); }
      if (bb < MG_MAX_BODIES) { const int s = T.slot(bb); pb = D2(T.PR(s, 0), T.PR(s, 1)); }
      d2 body_delta = dsub(pb, pa);
      double dist = ddot(dadd(dsub(r2, r1), body_delta), n);
      C[8] = -ax.contact_bias_coef * dminf(0.0, dist + MG_COLLISION_SLOP) / dt;
      C[11] = 0.0;
    }

    /* ---- joint prestep: chain */
    const double ang_r = T.PR(s_robot, 2);
    double b_gear;
    {
      const mg_joint_t& J = sc.joints[jr0 + 1];
      const double maxBias = J.max_bias;
      b_gear = dclamp(-ax.j_bcoef[jr0 + 1] * (ang_r * J.p1 - Pc.z - J.p0) / dt, -maxBias, maxBias);
    }
    TpePin pin0, pin1;
#define TPE_PIN_PRESTEP(PN, J_, SF)                                                                       \
  do {                                                                                                    \
    const mg_joint_t& J = sc.joints[J_];                                                                  \
    const double rac = T.PR(s_robot, 3), ras = T.PR(s_robot, 4), rbc = T.PR(SF, 3), rbs = T.PR(SF, 4);    \
    d2 r1 = D2(rac * J.anchor_a[0] - ras * J.anchor_a[1], ras * J.anchor_a[0] + rac * J.anchor_a[1]);     \
    d2 r2 = D2(rbc * J.anchor_b[0] - rbs * J.anchor_b[1], rbs * J.anchor_b[0] + rbc * J.anchor_b[1]);     \
    d2 pa = D2(T.PR(s_robot, 0), T.PR(s_robot, 1)), pb = D2(T.PR(SF, 0), T.PR(SF, 1));                    \
    d2 delta = dsub(dadd(pb, r2), dadd(pa, r1));                                                          \
    double dist = dlength(delta);                                                                         \
    d2 n = dmul(delta, 1.0 / (dist ? dist : MG_INF));                                                     \
    const TpeJC c = tpe_jc(ds, J_);                                                                       \
    double rcn1 = dcross(r1, n), rcn2 = dcross(r2, n);                                                    \
    double k = (c.ma + c.ia * rcn1 * rcn1) + (c.mb + c.ib * rcn2 * rcn2);                                 \
    const double maxBias = J.max_bias;                                                                    \
    PN.r1x = r1.x; PN.r1y = r1.y; PN.r2x = r2.x; PN.r2y = r2.y; PN.nx = n.x; PN.ny = n.y;                 \
    PN.nMass = 1.0 / k;                                                                                   \
    PN.bias = dclamp(-ax.j_bcoef[J_] * (dist - J.p0) / dt, -maxBias, maxBias);                            \
  } while (0)
    TPE_PIN_PRESTEP(pin0, jr0 + 4, s_f0);
    TPE_PIN_PRESTEP(pin1, jr0 + 7, s_f1);
    double b_lim0, b_lim1;
#define TPE_LIMIT_PRESTEP(B_, ACC, J_, SF)                                     \
  do {                                                                         \
    const mg_joint_t& J = sc.joints[J_];                                       \
    double dist = T.PR(SF, 2) - ang_r;                                         \
    double pdist = 0.0;                                                        \
    if (dist > J.p1) pdist = J.p1 - dist;                                      \
    else if (dist < J.p0) pdist = J.p0 - dist;                                 \
    const double maxBias = J.max_bias;                                         \
    B_ = dclamp(-ax.j_bcoef[J_] * pdist / dt, -maxBias, maxBias);              \
    if (!B_) ACC = 0.0;                                                        \
  } while (0)
    TPE_LIMIT_PRESTEP(b_lim0, a_lim0, jr0 + 5, s_f0);
    TPE_LIMIT_PRESTEP(b_lim1, a_lim1, jr0 + 8, s_f1);
    /* (the blocks' drag gears have max_bias = 0, checked by mg_build_tpe_aux: their bias is exactly +0.0) */

    /* chain velocities live in registers while joints run */
    double rvx = T.V(s_robot, 0), rvy = T.V(s_robot, 1), rw = T.V(s_robot, 2);
    double f0x = T.V(s_f0, 0), f0y = T.V(s_f0, 1), f0w = T.V(s_f0, 2);
    double f1x = T.V(s_f1, 0), f1y = T.V(s_f1, 1), f1w = T.V(s_f1, 2);
    double e0w = Ve0.z, e1w = Ve1.z;
    const TpeJC c_piv = tpe_jc(ds, jr0), c_gear = tpe_jc(ds, jr0 + 1), c_sp0 = tpe_jc(ds, jr0 + 2),
                c_sp1 = tpe_jc(ds, jr0 + 3), c_pin0 = tpe_jc(ds, jr0 + 4), c_lim0 = tpe_jc(ds, jr0 + 5),
                c_mot0 = tpe_jc(ds, jr0 + 6), c_pin1 = tpe_jc(ds, jr0 + 7), c_lim1 = tpe_jc(ds, jr0 + 8),
                c_mot1 = tpe_jc(ds, jr0 + 9);
    /* rotary springs: their preStep applies the spring impulse, sequentially in insertion order */
    double t_spr0 = 0.0, t_spr1 = 0.0;
    {
      const mg_joint_t& J = sc.joints[jr0 + 2];
      double j_spring = ((ang_r - Pe0.z) - J.p0) * J.p1 * MG_DT;
      a_spr0 = j_spring;
      rw -= j_spring * c_sp0.ia;
      e0w += j_spring * c_sp0.ib;
    }
    {
      const mg_joint_t& J = sc.joints[jr0 + 3];
      double j_spring = ((ang_r - Pe1.z) - J.p0) * J.p1 * MG_DT;
      a_spr1 = j_spring;
      rw -= j_spring * c_sp1.ia;
      e1w += j_spring * c_sp1.ib;
    }
    T.V(s_robot, 2) = rw;

    /* ---- warm start (cpArbiterApplyCachedImpulse, then the joints' applyCachedImpulse; dt_coef = 1) */
    uint32_t touched = 0; /* slots that carry a contact in this sub-step */
    for (int c = 0; c < ncon; c++) {
      TpeCon<S> C = tpe_con(T, c);
      const unsigned long long ids = tpe_unpack_ids(C[13]);
      const int ba = (int)((ids >> 48) & 0x1F), bb = (int)((ids >> 53) & 0x1F);
      const int sa_ = T.slot(ba), sb_ = T.slot(bb);
      touched |= (1u << sa_) | (1u << sb_);
      if ((ids >> 58) & 1ull) continue; /* first contact of the pair */
      const double ma = ba < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[ba].m_inv) : 0.0;
      const double ia_ = ba < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[ba].i_inv) : 0.0;
      const double mb = bb < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[bb].m_inv) : 0.0;
      const double ib_ = bb < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[bb].i_inv) : 0.0;
      const double nx = C[4], ny = C[5], cjn = C[9], cjt = C[10];
      const double jx = nx * cjn - ny * cjt, jy = nx * cjt + ny * cjn;
      TPE_APPLY(V, sa_, ma, ia_, -jx, -jy, C[0], C[1]);
      TPE_APPLY(V, sb_, mb, ib_, jx, jy, C[2], C[3]);
    }
    rvx = T.V(s_robot, 0); rvy = T.V(s_robot, 1); rw = T.V(s_robot, 2);
    f0x = T.V(s_f0, 0); f0y = T.V(s_f0, 1); f0w = T.V(s_f0, 2);
    f1x = T.V(s_f1, 0); f1y = T.V(s_f1, 1); f1w = T.V(s_f1, 2);
    {
      /* pivot (control -> robot): anchors at the body origins, no angular part */
      rvx = rvx + a_pivx * c_piv.mb; rvy = rvy + a_pivy * c_piv.mb;
      /* gear (control -> robot) */
      rw += a_gear * c_gear.ib;
      /* springs: nothing cached; pins, limits, motors of the two fingers */
#define TPE_PIN_WARM(PN, C_, ACC, FX, FY, FW)                      \
  do {                                                             \
    double jx = PN.nx * ACC, jy = PN.ny * ACC;                     \
    rvx = rvx + (-jx) * C_.ma; rvy = rvy + (-jy) * C_.ma;          \
    rw += C_.ia * (PN.r1x * (-jy) - PN.r1y * (-jx));               \
    FX = FX + jx * C_.mb; FY = FY + jy * C_.mb;                    \
    FW += C_.ib * (PN.r2x * jy - PN.r2y * jx);                     \
  } while (0)
#define TPE_ANG_WARM(C_, ACC, FW)   \
  do {                              \
    double da = ACC * C_.ia;        \
    rw -= da;                       \
    FW += ACC * C_.ib;              \
  } while (0)
      TPE_PIN_WARM(pin0, c_pin0, a_pin0, f0x, f0y, f0w);
      TPE_ANG_WARM(c_lim0, a_lim0, f0w);
      TPE_ANG_WARM(c_mot0, a_mot0, f0w);
      TPE_PIN_WARM(pin1, c_pin1, a_pin1, f1x, f1y, f1w);
      TPE_ANG_WARM(c_lim1, a_lim1, f1w);
      TPE_ANG_WARM(c_mot1, a_mot1, f1w);
    }
    /* A block that is at rest after the warm start (zero velocity, zero accumulated drag impulses) and
     * carries no contact stays exactly so through all iterations: its pivot and gear compute j = 0 and
     * add 0 (only the sign of a zero can differ, which nothing downstream reads).  Each lane therefore
     * walks only its own moving / touched blocks below, and the warp makes as many trips as its busiest
     * lane needs instead of one per block. */
    uint32_t bact = 0;
    for (int k = 0; k < nblk; k++) {
      const int s = ax.tpe_bj_slot[k];
      const TpeJC cp = tpe_jc(ds, ax.tpe_bj_pivot[k]), cg = tpe_jc(ds, ax.tpe_bj_gear[k]);
      const double j0 = T.BJ(k, 0), j1 = T.BJ(k, 1), j2 = T.BJ(k, 2);
      const double v0 = T.V(s, 0) + j0 * cp.mb, v1 = T.V(s, 1) + j1 * cp.mb;
      double v2 = T.V(s, 2);
      v2 += j2 * cg.ib;
      T.V(s, 0) = v0; T.V(s, 1) = v1; T.V(s, 2) = v2;
      const bool moving = v0 != 0.0 || v1 != 0.0 || v2 != 0.0 || j0 != 0.0 || j1 != 0.0 || j2 != 0.0 ||
                          ((touched >> s) & 1u) != 0u;
      bact |= (moving ? 1u : 0u) << k;
    }
#if defined(TPE_STATS) && !defined(__CUDACC__)
    tpe_stats[5] += nblk; /* block joints seen / walked */
    tpe_stats[6] += __builtin_popcount(bact);
#endif
    T.V(s_robot, 0) = rvx; T.V(s_robot, 1) = rvy; T.V(s_robot, 2) = rw;
    T.V(s_f0, 0) = f0x; T.V(s_f0, 1) = f0y; T.V(s_f0, 2) = f0w;
    T.V(s_f1, 0) = f1x; T.V(s_f1, 1) = f1y; T.V(s_f1, 2) = f1w;

    /* ---- solver iterations (cpArbiterApplyImpulse for every arbiter, then every joint) */
    for (int it = 0; it < MG_ITERATIONS; ++it) {
      for (int c = 0; c < ncon; c++) {
        TpeCon<S> C = tpe_con(T, c);
        const unsigned long long ids = tpe_unpack_ids(C[13]);
        const int ba = (int)((ids >> 48) & 0x1F), bb = (int)((ids >> 53) & 0x1F);
        const int sa_ = T.slot(ba), sb_ = T.slot(bb);
        const double c_ma = ba < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[ba].m_inv) : 0.0;
        const double c_ia = ba < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[ba].i_inv) : 0.0;
        const double c_mb = bb < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[bb].m_inv) : 0.0;
        const double c_ib = bb < MG_MAX_BODIES ? TPE_LDG(&sc.bodies[bb].i_inv) : 0.0;
        const double c_r1x = C[0], c_r1y = C[1], c_r2x = C[2], c_r2y = C[3], c_nx = C[4], c_ny = C[5];
        const double c_nMass = C[6], c_tMass = C[7], c_bias = C[8], c_u = C[12];
        double c_jn = C[9], c_jt = C[10], c_jb = C[11];
        const double vax = T.V(sa_, 0), vay = T.V(sa_, 1), vaz = T.V(sa_, 2);
        const double vbx = T.V(sb_, 0), vby = T.V(sb_, 1), vbz = T.V(sb_, 2);
        const double bax = T.Bv(sa_, 0), bay = T.Bv(sa_, 1), baz = T.Bv(sa_, 2);
        const double bbx = T.Bv(sb_, 0), bby = T.Bv(sb_, 1), bbz = T.Bv(sb_, 2);
        /* vb1 = a.v_bias + perp(r1)*a.w_bias, etc. */
        double vb1x = bax + (-c_r1y) * baz, vb1y = bay + c_r1x * baz;
        double vb2x = bbx + (-c_r2y) * bbz, vb2y = bby + c_r2x * bbz;
        double v1x = vax + (-c_r1y) * vaz, v1y = vay + c_r1x * vaz;
        double v2x = vbx + (-c_r2y) * vbz, v2y = vby + c_r2x * vbz;
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
        TPE_APPLY(Bv, sa_, c_ma, c_ia, -bjx, -bjy, c_r1x, c_r1y);
        TPE_APPLY(Bv, sb_, c_mb, c_ib, bjx, bjy, c_r2x, c_r2y);
        double dn = c_jn - jnOld, dtt = c_jt - jtOld;
        double jx = c_nx * dn - c_ny * dtt, jy = c_nx * dtt + c_ny * dn;
        TPE_APPLY(V, sa_, c_ma, c_ia, -jx, -jy, c_r1x, c_r1y);
        TPE_APPLY(V, sb_, c_mb, c_ib, jx, jy, c_r2x, c_r2y);
        C[9] = c_jn; C[10] = c_jt; C[11] = c_jb;
      }
      /* blocks: force-capped pivot + gear against the static body (entities.py:703-711); every lane walks
       * its own live blocks, one per trip (two per trip for instruction-level parallelism measured slower:
       * the trip count is set by the busiest lane and most lanes have at most one live block) */
      for (uint32_t m = bact; m != 0u; m &= m - 1u) {
        const int k = TPE_CTZ(m);
        const int sA = ax.tpe_bj_slot[k];
        const TpeJC cpA = tpe_jc(ds, ax.tpe_bj_pivot[k]), cgA = tpe_jc(ds, ax.tpe_bj_gear[k]);
        double vxA = T.V(sA, 0), vyA = T.V(sA, 1), wA = T.V(sA, 2);
        const double oxA = T.BJ(k, 0), oyA = T.BJ(k, 1), biasA = 0.0;
        double gaA = T.BJ(k, 2);
        double jxA = (0.0 - (vxA - 0.0)) * cpA.c0, jyA = (0.0 - (vyA - 0.0)) * cpA.c0;
        const d2 accA = dvclamp(D2(oxA + jxA, oyA + jyA), cpA.c1);
        jxA = accA.x - oxA; jyA = accA.y - oyA;
        vxA = vxA + jxA * cpA.mb; vyA = vyA + jyA * cpA.mb;
        double waA = 0.0;
        tpe_gear(cgA, biasA, gaA, waA, wA);
        T.BJ(k, 0) = accA.x; T.BJ(k, 1) = accA.y; T.BJ(k, 2) = gaA;
        T.V(sA, 0) = vxA; T.V(sA, 1) = vyA; T.V(sA, 2) = wA;
      }
      /* the robot chain, in insertion order (entities.py:255-354) */
      rvx = T.V(s_robot, 0); rvy = T.V(s_robot, 1); rw = T.V(s_robot, 2);
      f0x = T.V(s_f0, 0); f0y = T.V(s_f0, 1); f0w = T.V(s_f0, 2);
      f1x = T.V(s_f1, 0); f1y = T.V(s_f1, 1); f1w = T.V(s_f1, 2);
      {
        double jx = (0.0 - (rvx - Vc.x)) * c_piv.c0;
        double jy = (0.0 - (rvy - Vc.y)) * c_piv.c0;
        const double ox = a_pivx, oy = a_pivy;
        d2 acc = dvclamp(D2(ox + jx, oy + jy), c_piv.c1);
        a_pivx = acc.x; a_pivy = acc.y;
        jx = acc.x - ox; jy = acc.y - oy;
        rvx = rvx + jx * c_piv.mb; rvy = rvy + jy * c_piv.mb;
      }
      {
        double wa = Vc.z;
        tpe_gear(c_gear, b_gear, a_gear, wa, rw);
      }
      tpe_spring(c_sp0, t_spr0, a_spr0, rw, e0w);
      tpe_spring(c_sp1, t_spr1, a_spr1, rw, e1w);
      tpe_pin(c_pin0, pin0, a_pin0, rvx, rvy, rw, f0x, f0y, f0w);
      tpe_limit(c_lim0, b_lim0, a_lim0, rw, f0w);
      tpe_motor(c_mot0, rate0, a_mot0, rw, f0w);
      tpe_pin(c_pin1, pin1, a_pin1, rvx, rvy, rw, f1x, f1y, f1w);
      tpe_limit(c_lim1, b_lim1, a_lim1, rw, f1w);
      tpe_motor(c_mot1, rate1, a_mot1, rw, f1w);
      T.V(s_robot, 0) = rvx; T.V(s_robot, 1) = rvy; T.V(s_robot, 2) = rw;
      T.V(s_f0, 0) = f0x; T.V(s_f0, 1) = f0y; T.V(s_f0, 2) = f0w;
      T.V(s_f1, 0) = f1x; T.V(s_f1, 1) = f1y; T.V(s_f1, 2) = f1w;
    }
    Ve0.z = e0w; Ve1.z = e1w;

    /* ---- rebuild the contact cache: this sub-step's contacts first (canonical order), then the entries
     * of pairs that did not collide now and are younger than the persistence window
     * (cpSpaceArbiterSetFilter), in their old order */
    if (n_cache > 0 || ncon > 0) {
      int nsurv = 0;
      for (int q = 0; q < n_cache; q++) { /* compact the survivors in place (their rank never exceeds q) */
        const CEntry e = G->cache[q];
        if (((cache_used >> q) & 1ull) == 0ull && (stamp - e.stamp) < MG_PERSISTENCE) {
          if (nsurv != q) G->cache[nsurv] = e;
          nsurv++;
        }
      }
      int tot = ncon + nsurv;
      if (tot > MG_NCACHE) { tot = MG_NCACHE; overflow |= 4; }
      if (ncon > 0)
        for (int q = tot - ncon - 1; q >= 0; q--) G->cache[q + ncon] = G->cache[q]; /* shift up, back to front */
      for (int c = 0; c < ncon; c++) {
        TpeCon<S> C = tpe_con(T, c);
        const unsigned long long ids = tpe_unpack_ids(C[13]);
        CEntry e;
        e.a = (uint8_t)((ids >> 32) & 0xFF); e.b = (uint8_t)((ids >> 40) & 0xFF); e.used = 0; e.pad_ = 0;
        e.hash = (uint32_t)(ids & 0xFFFFFFFFull); e.stamp = stamp; e.pad2_ = 0;
        e.jn = C[9]; e.jt = C[10];
        G->cache[c] = e;
      }
      n_cache = tot;
    }
  }

  /* ---- store the record */
  if (!live) return;
  for (int s = 0; s < nslots - 1; s++) {
    const int b = ax.tpe_slot_body[s];
    G->V[b] = make_double4(T.V(s, 0), T.V(s, 1), T.V(s, 2), 0.0);
    G->Bv[b] = make_double4(T.Bv(s, 0), T.Bv(s, 1), T.Bv(s, 2), 0.0);
    G->P[b] = make_double4(T.PR(s, 0), T.PR(s, 1), T.PR(s, 2), 0.0);
    G->R[b] = make_double2(T.PR(s, 3), T.PR(s, 4));
  }
  G->P[control] = Pc; G->V[control] = Vc; G->Bv[control] = Bc; G->R[control] = Rc;
  G->P[eye0] = Pe0; G->V[eye0] = Ve0; G->Bv[eye0] = Be0; G->R[eye0] = Re0;
  G->P[eye1] = Pe1; G->V[eye1] = Ve1; G->Bv[eye1] = Be1; G->R[eye1] = Re1;
  for (int k = 0; k < nblk; k++) {
    G->jacc[ax.tpe_bj_pivot[k]] = make_double2(T.BJ(k, 0), T.BJ(k, 1));
    G->jacc[ax.tpe_bj_gear[k]].x = T.BJ(k, 2);
  }
  G->jacc[jr0] = make_double2(a_pivx, a_pivy);
  G->jacc[jr0 + 1].x = a_gear; G->jacc[jr0 + 2].x = a_spr0; G->jacc[jr0 + 3].x = a_spr1;
  G->jacc[jr0 + 4].x = a_pin0; G->jacc[jr0 + 5].x = a_lim0; G->jacc[jr0 + 6].x = a_mot0;
  G->jacc[jr0 + 7].x = a_pin1; G->jacc[jr0 + 8].x = a_lim1; G->jacc[jr0 + 9].x = a_mot1;
  G->stamp = stamp;
  G->n_cache = n_cache;
  G->overflow = overflow;
  G->last_contacts = ncon;
}

#endif /* MG_PHYSICS_TPE_H */
