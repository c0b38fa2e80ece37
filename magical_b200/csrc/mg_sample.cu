/*
 * mg_sample.cu — K5 `k_sample_layouts`: reset-time randomisation of goal sizes and entity poses ON THE DEVICE
 * (SURVEY 8(f) N1), one warp per resetting environment.
 *
 * Replaces, at reset time, the reference's rejection sampler
 *     randomise_hw              magical/geom.py:344-359     one uniform (h, w) draw per goal region
 *     pm_randomise_all_poses    magical/geom.py:281-341     entities placed one after the other, <= 10 retries
 *     pm_randomise_pose         magical/geom.py:116-264     <= 10 000 uniform (x, y, angle) tries per entity, a try
 *                                                           is rejected when any of the entity's shapes touches
 *                                                           (space.shape_query) a wall, a fixed entity, or an
 *                                                           entity placed before it; goal sensors count
 *     pm_shift_bodies           magical/geom.py:362-384     the entity's bodies move rigidly with the main body
 * The structure of the scene (shape types, colours, counts, dynamics) comes from a host-built TEMPLATE; the
 * kernel copies the template into the environment's own scene slot, draws sizes and poses, patches the slot
 * (body reset poses, goal sensors, the goals' world-space draw rectangles) and resets the environment from it.
 *
 * The 32 lanes evaluate 32 consecutive tries of an entity at once and the FIRST successful try wins, which is
 * exactly what trying them one after the other yields.  Random numbers are Philox4x32-10 keyed by
 * (reset_seed, env) with the counter (reset count | retry, entity, try, draw): the reference's distribution,
 * not numpy's MT19937 stream ("match distributions, not streams").
 * The overlap predicate is the product's own exact narrowphase (mg_collide: contact count > 0), the same code
 * GoalRegion.get_overlapping_ents uses in mg_finish.cu.
 */
#include "mg_device.cuh"
#include "mg_narrowphase.h"
#include "mg_sincos.h"
#include "mg_reset.cuh"

#define SAMPLE_WARPS 4
#define SAMPLE_MAX_TRIES 10000   /* geom.py:200 */
#define SAMPLE_MAX_RETRIES 10    /* geom.py:290 */

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], const uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
/* two uniform doubles in [0, 1) (53 random bits each) from one Philox4x32-10 block */
__device__ static void philox_uniform2(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                       double* u, double* v) {
  uint32_t c[4] = {c0, c1, c2, c3}, k[2] = {k0, k1};
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
  *u = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) * (1.0 / 9007199254740992.0);
  *v = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) * (1.0 / 9007199254740992.0);
}

struct WarpPlace {
  double x[MG_MAX_BODIES], y[MG_MAX_BODIES], a[MG_MAX_BODIES], c[MG_MAX_BODIES], s[MG_MAX_BODIES];
  double goal[MG_MAX_GOALS][4]; /* cx, cy, w, h */
};

__device__ static ShapeView placed_view(const DeviceScene* ds, int si, double px, double py, double rc, double rs) {
  const mg_shape_t& sh = ds->s.shapes[si];
  ShapeView v;
  v.kind = sh.kind;
  v.nvert = sh.nvert;
  v.lv = &ds->s.cverts[sh.vert0][0];
  v.ln = &ds->aux.cnorm[sh.vert0][0];
  v.radius = sh.radius;
  v.index = si;
  v.px = px; v.py = py; v.rc = rc; v.rs = rs;
  return v;
}

__device__ static bool views_touch(const ShapeView& a, const double* bba, const ShapeView& b, const double* bbb) {
  if (!bb_intersects(bba, bbb)) return false;
  Manifold m;
  /* cpCollide order: lower shape type first */
  if (a.kind <= b.kind) mg_collide(a, b, bba, bbb, m);
  else mg_collide(b, a, bbb, bba, m);
  return m.count > 0;
}

/* does shape view `v` (bounding box bb) touch any obstacle: walls, placed bodies' shapes, placed goal sensors */
__device__ static bool touches_obstacles(const DeviceScene* ds, const WarpPlace& wp, const ShapeView& v, const double* bb,
                                         uint32_t placed_bodies, uint32_t placed_goals) {
  const mg_scene_t& sc = ds->s;
  for (int g = 0; g < sc.n_cgroups; g++) {
    const int body = sc.cgroups[g].body;
    if (body >= 0 && ((placed_bodies >> body) & 1u) == 0u) continue;
    const int s0 = sc.cgroups[g].shape0, n = sc.cgroups[g].nshape;
    for (int si = s0; si < s0 + n; si++) {
      ShapeView o = body >= 0 ? placed_view(ds, si, wp.x[body], wp.y[body], wp.c[body], wp.s[body])
                              : placed_view(ds, si, 0.0, 0.0, 1.0, 0.0);
      double obb[4];
      sv_bb(o, obb);
      if (views_touch(v, bb, o, obb)) return true;
    }
  }
  for (int g = 0; g < sc.n_goals; g++) {
    if (((placed_goals >> g) & 1u) == 0u) continue;
    const double hw = wp.goal[g][2] / 2, hh = wp.goal[g][3] / 2;
    double lv[8] = {hw, -hh, hw, hh, -hw, hh, -hw, -hh}; /* Poly.create_box vertex order */
    double ln[8] = {0, -1, 1, 0, 0, 1, -1, 0};
    ShapeView o;
    o.kind = MG_SHAPE_POLY; o.nvert = 4; o.lv = lv; o.ln = ln; o.radius = 0.0; o.index = MG_MAX_SHAPES;
    o.px = wp.goal[g][0]; o.py = wp.goal[g][1]; o.rc = 1.0; o.rs = 0.0;
    double obb[4];
    sv_bb(o, obb);
    if (views_touch(v, bb, o, obb)) return true;
  }
  return false;
}

__global__ void __launch_bounds__(SAMPLE_WARPS * 32)
k_sample_layouts(EnvState* __restrict__ states, DeviceScene* __restrict__ scenes, const mg_placement_t* __restrict__ programs,
                 int n_templates, int batch, uint32_t seed, unsigned long long* __restrict__ failures) {
  __shared__ WarpPlace s_wp[SAMPLE_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int env = blockIdx.x * SAMPLE_WARPS + wid;
  if (env >= batch) return;
  EnvState& st = states[env];
  if (st.fresh != MG_FRESH_SAMPLE) return; /* warp-uniform */
  const int tmpl = st.scene;
  const int slot = n_templates + env;
  WarpPlace& wp = s_wp[wid];
  /* 1. the template becomes the environment's own scene */
  {
    const uint4* src = reinterpret_cast<const uint4*>(scenes + tmpl);
    uint4* dst = reinterpret_cast<uint4*>(scenes + slot);
    const int n16 = (int)(sizeof(DeviceScene) / sizeof(uint4));
    for (int i = lane; i < n16; i += 32) dst[i] = src[i];
  }
  __syncwarp();
  DeviceScene* ds = scenes + slot;
  mg_scene_t& sc = ds->s;
  const mg_placement_t& P = programs[tmpl];
  const uint32_t k0 = seed, k1 = (uint32_t)env * 0x9E3779B1u + 0x7F4A7C15u;
  const uint32_t resets = (uint32_t)st.resets;

  /* 2. pose table from the template (fixed entities keep these poses); which bodies / goals are randomised */
  uint32_t listed_bodies = 0u, listed_goals = 0u;
  for (int k = 0; k < P.n_ents; k++) {
    const mg_place_ent_t& E = P.ents[k];
    if (E.kind == 1) listed_goals |= 1u << E.goal;
    else for (int j = 0; j < E.n_bodies; j++) listed_bodies |= 1u << E.bodies[j];
  }
  if (lane < sc.n_bodies) {
    wp.x[lane] = sc.bodies[lane].p0[0]; wp.y[lane] = sc.bodies[lane].p0[1]; wp.a[lane] = sc.bodies[lane].a0;
    double sn, cs;
    mg_det_sincos(wp.a[lane], &sn, &cs);
    wp.c[lane] = cs; wp.s[lane] = sn;
  }
  if (lane < sc.n_goals) {
    wp.goal[lane][0] = sc.goals[lane].cx; wp.goal[lane][1] = sc.goals[lane].cy;
    wp.goal[lane][2] = sc.goals[lane].w; wp.goal[lane][3] = sc.goals[lane].h;
  }
  __syncwarp();

  /* 3. goal sizes (randomise_hw: one uniform (h, w) draw; the region keeps its TOP-LEFT corner) */
  if (lane == 0) {
    for (int i = 0; i < P.n_hw; i++) {
      const mg_place_hw_t& H = P.hw[i];
      double lo_h = H.min_side, hi_h = H.max_side, lo_w = H.min_side, hi_w = H.max_side;
      if (H.linf >= 0.0) {
        lo_h = fmax(lo_h, H.cur_h - H.linf); hi_h = fmin(hi_h, H.cur_h + H.linf);
        lo_w = fmax(lo_w, H.cur_w - H.linf); hi_w = fmin(hi_w, H.cur_w + H.linf);
      }
      double u, v;
      philox_uniform2(k0, k1, resets, 0xFFFF0000u + (uint32_t)i, 0u, 0u, &u, &v);
      const double h = lo_h + (hi_h - lo_h) * u, w = lo_w + (hi_w - lo_w) * v;
      wp.goal[H.goal][2] = w; wp.goal[H.goal][3] = h;
    }
    /* centres of the listed goals before their pose is randomised: top-left corner + half size */
    for (int k = 0; k < P.n_ents; k++) {
      const mg_place_ent_t& E = P.ents[k];
      if (E.kind != 1) continue;
      wp.goal[E.goal][0] = E.orig[0] + wp.goal[E.goal][2] / 2;
      wp.goal[E.goal][1] = E.orig[1] - wp.goal[E.goal][3] / 2;
    }
  }
  __syncwarp();

  /* 4. poses: entities in list order, 32 tries at a time, first success wins */
  bool placed_all = false;
  for (int retry = 0; retry < SAMPLE_MAX_RETRIES && !placed_all; retry++) {
    uint32_t placed_bodies = ~listed_bodies, placed_goals = ~listed_goals;
    bool failed = false;
    for (int k = 0; k < P.n_ents && !failed; k++) {
      const mg_place_ent_t& E = P.ents[k];
      /* the centre the relative limits refer to */
      double ox, oy, oa;
      if (E.kind == 1) { ox = wp.goal[E.goal][0]; oy = wp.goal[E.goal][1]; oa = 0.0; }
      else { ox = E.orig[0]; oy = E.orig[1]; oa = E.orig[2]; }
      double x_lo = P.arena[0], x_hi = P.arena[1], y_lo = P.arena[2], y_hi = P.arena[3];
      if (E.pos_limit >= 0.0) {
        x_lo = fmax(x_lo, ox - E.pos_limit); x_hi = fmin(x_hi, ox + E.pos_limit);
        y_lo = fmax(y_lo, oy - E.pos_limit); y_hi = fmin(y_hi, oy + E.pos_limit);
      }
      double a_lo = -3.14159265358979323846, a_hi = 3.14159265358979323846;
      if (E.rot_limit >= 0.0) { a_lo = oa - E.rot_limit; a_hi = oa + E.rot_limit; }
      /* rigid offsets of the entity's bodies in the main body's frame (from the template: a rigid configuration) */
      const int root = E.kind == 0 ? E.bodies[0] : 0;
      const double r_x = E.kind == 0 ? sc.bodies[root].p0[0] : 0.0, r_y = E.kind == 0 ? sc.bodies[root].p0[1] : 0.0;
      const double r_a = E.kind == 0 ? sc.bodies[root].a0 : 0.0;
      double rsn, rcs;
      mg_det_sincos(r_a, &rsn, &rcs);
      int won = -1;
      double wx = 0.0, wy = 0.0, wa = 0.0;
      for (int base = 0; base < SAMPLE_MAX_TRIES && won < 0; base += 32) {
        const int t = base + lane;
        double u0, u1, u2, u3;
        philox_uniform2(k0, k1, resets | ((uint32_t)retry << 24), (uint32_t)k, (uint32_t)t, 0u, &u0, &u1);
        philox_uniform2(k0, k1, resets | ((uint32_t)retry << 24), (uint32_t)k, (uint32_t)t, 1u, &u2, &u3);
        const double nx = E.rand_pos ? x_lo + (x_hi - x_lo) * u0 : ox;
        const double ny = E.rand_pos ? y_lo + (y_hi - y_lo) * u1 : oy;
        const double na = E.rand_rot ? a_lo + (a_hi - a_lo) * u2 : oa;
        bool ok = t < SAMPLE_MAX_TRIES;
        if (ok) {
          if (E.kind == 1) {
            const double hw = wp.goal[E.goal][2] / 2, hh = wp.goal[E.goal][3] / 2;
            double lv[8] = {hw, -hh, hw, hh, -hw, hh, -hw, -hh};
            double ln[8] = {0, -1, 1, 0, 0, 1, -1, 0};
            ShapeView v;
            v.kind = MG_SHAPE_POLY; v.nvert = 4; v.lv = lv; v.ln = ln; v.radius = 0.0; v.index = MG_MAX_SHAPES;
            v.px = nx; v.py = ny; v.rc = 1.0; v.rs = 0.0;
            double bb[4];
            sv_bb(v, bb);
            ok = !touches_obstacles(ds, wp, v, bb, placed_bodies, placed_goals);
          } else {
            double nsn, ncs;
            mg_det_sincos(na, &nsn, &ncs);
            for (int g = 0; g < E.n_groups && ok; g++) {
              const mg_cgroup_t& grp = sc.cgroups[E.groups[g]];
              const int b = grp.body;
              /* body b in the main body's frame, then at the candidate pose */
              const double dx = sc.bodies[b].p0[0] - r_x, dy = sc.bodies[b].p0[1] - r_y;
              const double lx = rcs * dx + rsn * dy, ly = -rsn * dx + rcs * dy;
              const double bx = nx + (ncs * lx - nsn * ly), by = ny + (nsn * lx + ncs * ly);
              const double ba = na + (sc.bodies[b].a0 - r_a);
              double bsn, bcs;
              mg_det_sincos(ba, &bsn, &bcs);
              for (int si = grp.shape0; si < grp.shape0 + grp.nshape && ok; si++) {
                ShapeView v = placed_view(ds, si, bx, by, bcs, bsn);
                double bb[4];
                sv_bb(v, bb);
                ok = !touches_obstacles(ds, wp, v, bb, placed_bodies, placed_goals);
              }
            }
          }
        }
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (okm) {
          const int src = __ffs(okm) - 1;
          won = base + src;
          wx = __shfl_sync(0xffffffffu, nx, src);
          wy = __shfl_sync(0xffffffffu, ny, src);
          wa = __shfl_sync(0xffffffffu, na, src);
        }
      }
      if (won < 0) { failed = true; break; }
      /* commit the entity's pose */
      if (E.kind == 1) {
        if (lane == 0) { wp.goal[E.goal][0] = wx; wp.goal[E.goal][1] = wy; }
        placed_goals |= 1u << E.goal;
      } else {
        double nsn, ncs;
        mg_det_sincos(wa, &nsn, &ncs);
        if (lane < E.n_bodies) {
          const int b = E.bodies[lane];
          const double dx = sc.bodies[b].p0[0] - r_x, dy = sc.bodies[b].p0[1] - r_y;
          const double lx = rcs * dx + rsn * dy, ly = -rsn * dx + rcs * dy;
          wp.x[b] = wx + (ncs * lx - nsn * ly);
          wp.y[b] = wy + (nsn * lx + ncs * ly);
          wp.a[b] = wa + (sc.bodies[b].a0 - r_a);
          double bsn, bcs;
          mg_det_sincos(wp.a[b], &bsn, &bcs);
          wp.c[b] = bcs; wp.s[b] = bsn;
        }
        for (int j = 0; j < E.n_bodies; j++) placed_bodies |= 1u << E.bodies[j];
      }
      __syncwarp();
    }
    placed_all = !failed;
    if (failed) {
      /* start over from the template's configuration (pm_randomise_all_poses restarts the whole list) */
      __syncwarp();
      if (lane < sc.n_bodies) {
        wp.x[lane] = sc.bodies[lane].p0[0]; wp.y[lane] = sc.bodies[lane].p0[1]; wp.a[lane] = sc.bodies[lane].a0;
        double sn, cs;
        mg_det_sincos(wp.a[lane], &sn, &cs);
        wp.c[lane] = cs; wp.s[lane] = sn;
      }
      __syncwarp();
    }
  }

  /* 5. patch the slot: body reset poses, goal sensors and their draw rectangles; no placement found within the
   * reference's limits: the template's own (host-sampled, valid) layout is played and the failure is counted */
  if (placed_all) {
    if (lane < sc.n_bodies) {
      sc.bodies[lane].p0[0] = wp.x[lane]; sc.bodies[lane].p0[1] = wp.y[lane]; sc.bodies[lane].a0 = wp.a[lane];
    }
    if (lane < sc.n_goals) {
      const double cx = wp.goal[lane][0], cy = wp.goal[lane][1], w = wp.goal[lane][2], h = wp.goal[lane][3];
      sc.goals[lane].cx = cx; sc.goals[lane].cy = cy; sc.goals[lane].w = w; sc.goals[lane].h = h;
      /* gym_render.make_rect order: (-w/2, h/2) (w/2, h/2) (w/2, -h/2) (-w/2, -h/2), plus the centre */
      const double rw = w / 2, rh = h / 2;
      const double px[4] = {-rw, rw, rw, -rw}, py[4] = {rh, rh, -rh, -rh};
      for (int q = 0; q < 2; q++) {
        const int pi = P.goal_prims[lane][q];
        if (pi < 0) continue;
        const int v0 = sc.prims[pi].vert0, rv0 = ds->ra.voff[pi];
        for (int c = 0; c < 4; c++) {
          const float fx = (float)(px[c] + cx), fy = (float)(py[c] + cy);
          sc.dverts[v0 + c][0] = fx; sc.dverts[v0 + c][1] = fy;
          ds->ra.lv[rv0 + c][0] = fx; ds->ra.lv[rv0 + c][1] = fy;
        }
      }
    }
  } else if (lane == 0 && failures) {
    atomicAdd(failures, 1ull);
  }
  __syncwarp();
  /* 6. the environment restarts from its slot's poses.  Its scene index stays the TEMPLATE: the physics kernel
   * reads nothing that the sampler changes, and 8192 environments walking 64 shared templates stay in L2
   * where 8192 private 76 KB copies would not (measured: k_physics_tpe 1.76x slower on private copies);
   * the kernels that need the sampled parts (render: goal rectangles, finish: goal sensors) address the slot. */
  if (lane == 0) mg_reset_state(st, ds, tmpl);
}

cudaError_t mg_launch_sample_layouts(EnvState* states, DeviceScene* scenes, const mg_placement_t* programs,
                                     int n_templates, int batch, uint32_t seed, unsigned long long* failures,
                                     cudaStream_t stream) {
  k_sample_layouts<<<(batch + SAMPLE_WARPS - 1) / SAMPLE_WARPS, SAMPLE_WARPS * 32, 0, stream>>>(
      states, scenes, programs, n_templates, batch, seed, failures);
  return cudaGetLastError();
}
