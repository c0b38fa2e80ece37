/*
 * mg_raster.cu — K3: 2D rasteriser + 4x4 area downsample + frame stack,
 * writing straight into the caller's device observation tensor.
 *
 * Replaces, for a whole batch in one launch:
 *   BaseEnv.render -> Viewer.render x2 (pyglet/OpenGL immediate mode + FBO
 *     readback)                          base_env.py:309-338, gym_render.py:208-249
 *   camera transforms                    gym_render.py:176-200, 362-380
 *   FlattenFrameStack / EagerDictFrameStack   benchmarks/__init__.py:46-136
 *   ResizeObservation / ResizeDictObservation (cv2 INTER_AREA 384->96)
 *                                        benchmarks/__init__.py:139-169, 234
 *   ChannelsFirst                        benchmarks/__init__.py:172-192
 *
 * One CTA per environment, all in shared memory:
 *   1. every draw primitive is transformed to window space for each needed
 *      camera (fp32, the CPU oracle's operation order) -> oriented edge
 *      equations;
 *   2. SPAN TABLE: for every (primitive, sample row) the exact interval
 *      [lo, hi] of covered sample columns.  A sample is covered iff every edge
 *      function fmaf(A, x, fmaf(B, y, C)) is >= 0; each edge function is
 *      monotone in x along a row (fp32 rounding is monotone), so the covered
 *      set is an interval whose ends are found by one division estimate and a
 *      +-1 correction using the exact expression.  This gives bit-identical
 *      coverage to brute-force per-sample testing with ~rows x edges work
 *      instead of pixels x samples x edges;
 *   3. primitives are binned into a 12x12 tile grid from the spans, together
 *      with the top-most primitive that covers each tile completely;
 *   4. half-warps own tiles (uniform control flow); per output pixel the
 *      primitives are walked FRONT TO BACK, the 4x4 sub-sample mask of a
 *      primitive is assembled from four span rows with integer ops only, and
 *      the colour sum of the 16 samples is rounded half-to-even -- identical
 *      to rendering 384x384 and box filtering (cv2 INTER_AREA);
 *   5. four pixels = 48 bytes of the frame stack are shifted and rewritten
 *      with three 128-bit load/store pairs, so HBM sees the algorithmic
 *      minimum: read the 3 frames that survive, write 4.
 */
#include <cstdlib>
#include "mg_device.cuh"

#define RGRID 12           /* tiles per side */
#define RMAXP 192          /* window-space primitives (draw prims + expanded line segments) */
#define RWORDS (RMAXP / 32)
#define RLONG MG_RLONG      /* capacity of the compact long-edge list */
#define RSHORT MG_RSHORT   /* polygon edges bounding fewer rows than this are processed one edge per lane */
#define RMAXLINES 16       /* thick line segments whose rows are spread over all warps (more: per-warp fallback) */
#define BG_R 231
#define BG_G 231
#define BG_B 234
#define ZOOM 1.02
#ifndef RASTER_THREADS
#define RASTER_THREADS 256
#endif
#ifndef RASTER_MIN_BLOCKS
#define RASTER_MIN_BLOCKS 4
#endif

#ifdef RASTER_PROF
__device__ unsigned long long g_rprof[16];
#define RPROF_DECL long long rp_t = clock64();
#define RPROF(i) do { if (threadIdx.x == 0) { long long n_ = clock64(); atomicAdd(&g_rprof[i], (unsigned long long)(n_ - rp_t)); rp_t = n_; } } while (0)
#else
#define RPROF_DECL
#define RPROF(i) ((void)0)
#endif
#if defined(RASTER_PROF) && defined(RASTER_COUNT)
#define RCOUNT(i, n) atomicAdd(&g_rprof[i], (unsigned long long)(n))
#else
#define RCOUNT(i, n) ((void)0)
#endif

struct RPrim {
  uint32_t rgb;         /* bits 0..23 colour; bit 24: stippled line */
  uint16_t e0;          /* first edge (polygons) / first of the 2 float4 of a line segment */
  uint16_t ne;          /* edge count; 0 = line segment */
  int16_t row0, nrows;  /* sample rows covered by the bounding box */
  int16_t col0, col1;   /* sample columns of the bounding box (inclusive) */
  int32_t span0;        /* offset of this primitive's rows in the span table */
  float sgn;            /* winding sign of the window-space polygon */
  int32_t tile0;        /* offset of this primitive's (primitive, tile row) work items in the binning pass */
  float ymin, ymax;     /* vertical extent of the window-space vertices */
};
struct Camera {
  float cc, cs, cx, cy, nx, ny, S;
};

__constant__ double c_unit[130][2]; /* unit circles with 10, 20, 100 vertices (gym_render.make_circle) */

__device__ __forceinline__ int unit_offset(int n) { return n == 10 ? 0 : (n == 20 ? 10 : 30); }

__device__ __forceinline__ float2 world_to_px(const Camera& cam, float wx, float wy) {
  float dx = wx - cam.cx, dy = wy - cam.cy;
  float rx = fmaf(cam.cc, dx, cam.cs * dy);
  float ry = fmaf(cam.cc, dy, -(cam.cs * dx));
  return make_float2((rx + cam.nx) * cam.S, (ry + cam.ny) * cam.S);
}
__device__ __forceinline__ float2 body_to_world(const EnvState& st, int b, float vx, float vy) {
  float bc = (float)st.R[b].x, bs = (float)st.R[b].y, bx = (float)st.P[b].x, by = (float)st.P[b].y;
  return make_float2(fmaf(bc, vx, fmaf(-bs, vy, bx)), fmaf(bs, vx, fmaf(bc, vy, by)));
}

__device__ __forceinline__ Camera make_camera(const EnvState& st, const mg_scene_t& sc, int view, int res_full) {
  Camera cam;
  cam.S = (float)((double)res_full / (2.0 * ZOOM));
  if (view == 0) {
    cam.cc = 1.0f; cam.cs = 0.0f; cam.cx = 0.0f; cam.cy = 0.0f;
    cam.nx = (float)ZOOM; cam.ny = (float)ZOOM;
  } else {
    int r = sc.robot_body;
    cam.cc = (float)st.R[r].x; cam.cs = (float)st.R[r].y;
    cam.cx = (float)st.P[r].x; cam.cy = (float)st.P[r].y;
    cam.nx = (float)(2.0 * ZOOM * 0.5); cam.ny = (float)(2.0 * ZOOM * 0.15);
  }
  return cam;
}

/* window-space vertex of draw primitive pr whose local-space coordinates are (vx, vy) (static table ra.lv) */
__device__ __forceinline__ float2 prim_vertex(const EnvState& st, const mg_prim_t& pr, float vx, float vy,
                                              const Camera& cam) {
  if (pr.kind == MG_PRIM_NGON) {
    if (pr.xform == MG_XFORM_PUPIL) {
      int b = pr.body, e = pr.body2;
      float pc = (float)__dadd_rn(__dmul_rn(st.R[e].x, st.R[b].x), __dmul_rn(st.R[e].y, st.R[b].y));
      float ps = (float)__dadd_rn(__dmul_rn(st.R[e].y, st.R[b].x), -__dmul_rn(st.R[e].x, st.R[b].y));
      float ux = vx + pr.ex, uy = vy + pr.ey;
      vx = fmaf(pc, ux, -(ps * uy));
      vy = fmaf(ps, ux, pc * uy);
    }
    vx += pr.cx; vy += pr.cy;
    float2 w = body_to_world(st, pr.body, vx, vy);
    return world_to_px(cam, w.x, w.y);
  }
  float2 w = make_float2(vx, vy);
  if (pr.xform == MG_XFORM_BODY) w = body_to_world(st, pr.body, vx, vy);
  return world_to_px(cam, w.x, w.y);
}

struct ViewSmem {
  int rcap, rwords; /* capacity in window-space primitives (multiple of 32, <= RMAXP) and mask words per tile */
  RPrim* prims;     /* [rcap] */
  float4* edges;    /* [ecap + 2*rcap]  (A, B, C, prim | rows << 8) ; line segments: 2 float4 behind the pool */
  float4* eaux;     /* [ecap] per edge: (m, q, ymin, ymax) with boundary x*(y) = m*y + q */
  float2* verts;    /* [ecap]; dead once the edge equations exist, so it lives inside the span table */
  short2* spans;    /* [scap] (lo, hi) per (primitive, row) */
  uint32_t* tiles;  /* [RGRID*RGRID*rwords] */
  int32_t* cover;   /* [RGRID*RGRID] top-most primitive covering the whole tile, -1 = none */
};

/* first index in [cmin, cmax+1] from which the monotone predicate holds (false...false true...true) */
template <class F>
__device__ __forceinline__ int first_true(float est, int cmin, int cmax, F ok) {
  float e = fminf(fmaxf(est, (float)cmin - 1.0f), (float)cmax + 1.0f);
  int i = (int)ceilf(e);
  i = i < cmin ? cmin : (i > cmax + 1 ? cmax + 1 : i);
  /* the estimate is almost always right: the predicate fails just below i and holds at i.  Both are evaluated
   * unconditionally so that the lanes of a warp stay converged; the search loops only run for the rare miss */
  const bool below = ok(i - 1) && i > cmin;
  const bool at = ok(i) || i > cmax;
  if (below || !at) {
    while (i > cmin && ok(i - 1)) i--;
    while (i <= cmax && !ok(i)) i++;
  }
  return i;
}
/* exact covered interval of sample row j for a thick line segment R (empty => lo > hi); polygons get their spans
 * from their edges in build_view.  A sample is covered unless along < 0 || along > L || |perp| > hw: four half-lines
 * in x, each bounded by a monotone fp32 expression, intersected in turn by the same search (columns mirrored where
 * the predicate falls instead of rising).  One loop instead of four unrolled searches: the kernel's size matters
 * more (instruction cache) than these few rows. */
__device__ __forceinline__ short2 row_span(const RPrim& R, const float4* __restrict__ edges, int j) {
  int lo = R.col0, hi = R.col1;
  const float y = (float)j + 0.5f;
  const float4 p = edges[R.e0], q = edges[R.e0 + 1];
  const float ry = y - p.y;
  const float ca = ry * p.w;      /* along = fmaf(rx, ux, ry*uy) */
  const float cp = -(ry * p.z);   /* perp  = fmaf(rx, uy, -(ry*ux)) */
  const float L = q.x, hw = q.z;
  if (L < 0.0f) hi = lo - 1;
#pragma unroll 1
  for (int c = 0; c < 4 && lo <= hi; c++) {
    /* c = 0, 1: perp >= -hw, perp <= hw (for the thin borders they leave a sample or two); 2, 3: along >= 0, <= L */
    const bool upper = (c & 1) != 0;
    const float coef = c < 2 ? p.w : p.z;
    const float off = c < 2 ? cp : ca;
    const float thr = c < 2 ? (upper ? hw : -hw) : (upper ? L : 0.0f);
    auto holds = [&](int i) {
      const float v = fmaf(((float)i + 0.5f) - p.x, coef, off);
      return upper ? !(v > thr) : !(v < thr);
    };
    if (coef == 0.0f) {
      if (!holds(lo)) hi = lo - 1; /* the same for every column */
      continue;
    }
    const int sg = ((coef > 0.0f) != upper) ? 1 : -1; /* +1: false ... true along x; -1: true ... false */
    const float est = p.x + (thr - off) / coef - 0.5f;
    const int u = first_true(sg > 0 ? est : -est, sg > 0 ? lo : -hi, sg > 0 ? hi : -lo, [&](int m) { return holds(sg * m); });
    if (sg > 0) lo = u; else hi = -u;
  }
  return make_short2((short)lo, (short)hi);
}

/* Build the window-space primitive set, span table and tile bins of one view. */
template <int SS>
__device__ void build_view(ViewSmem& vs, const EnvState& st, const mg_scene_t& sc, const mg_raster_aux_t& ra,
                           const mg_raster_aux_t* lv_own /* device-sampled layouts: the environment's own copy */,
                           int view, int res_out, int ecap, int scap, int* s_off /* [RLONG] */, int* s_misc) {
  RPROF_DECL
  const int tid = threadIdx.x;
  constexpr int nt = RASTER_THREADS; /* compile-time block size: no runtime divisions by the warp count */
  const int np = sc.n_prims;
  const int res_full = res_out * SS;
  const Camera cam = make_camera(st, sc, view, res_full);
  const float px_scale = (float)res_full / 384.0f;
  /* vertex offsets, window-primitive indices and the vertex -> primitive map are static per scene (ra) */
  const int nv = min((int)ra.nv, ecap);
  const int nrp = min((int)ra.nrp, vs.rcap);
#pragma unroll 1
  for (int i = tid; i < RGRID * RGRID * vs.rwords; i += nt) vs.tiles[i] = 0u;
  for (int i = tid; i < RGRID * RGRID; i += nt) vs.cover[i] = -1;
  if (tid == 0) {
    s_misc[7] = 0; /* number of thick line segments collected by phase C */
    s_misc[8] = 0; s_misc[9] = 0; s_misc[10] = 0; s_misc[11] = 0; /* flat / heavy / light tile lists, busy-tile hand-out */
    s_misc[3] = 0; /* primitive hand-out counter of the span phase */
  }
  /* B: window-space vertices */
  for (int v = tid; v < nv; v += nt) {
    const mg_prim_t& pr = sc.prims[ra.vprim[v]];
    /* world-fixed polygons and borders (the goal rectangles among them) come from the environment's own table
     * when layouts are sampled on the device; everything else from the shared, cache-resident template */
    const bool own = lv_own != nullptr && pr.kind != MG_PRIM_NGON && pr.xform != MG_XFORM_BODY;
    const float2 l = *reinterpret_cast<const float2*>(own ? lv_own->lv[v] : ra.lv[v]);
    vs.verts[v] = prim_vertex(st, pr, l.x, l.y, cam);
  }
  __syncthreads();
  RPROF(1);
  /* C: per-primitive records: bounding box in samples (the oracle's loop bounds), winding */
  for (int p = tid; p < np; p += nt) {
    const mg_prim_t& pr = sc.prims[p];
    int v0 = ra.voff[p], n = pr.nvert;
    int rp0 = ra.rp0[p];
    if (v0 + n > nv) continue;
    uint32_t rgb = pr.rgb[0] | (pr.rgb[1] << 8) | (pr.rgb[2] << 16);
    if (pr.kind == MG_PRIM_LINELOOP) {
      float hw = 0.5f * pr.radius * px_scale;
      float s0 = 0.0f;
      for (int k = 0; k < n; k++) {
        float2 a = vs.verts[v0 + k], b = vs.verts[v0 + (k + 1 == n ? 0 : k + 1)];
        float dx = b.x - a.x, dy = b.y - a.y;
        float L = sqrtf(fmaf(dx, dx, dy * dy));
        if (rp0 + k < vs.rcap) {
          /* a segment's two float4 (ax, ay, ux, uy), (L, s0, hw, stipple) live behind the edge pool */
          RPrim& R = vs.prims[rp0 + k];
          const int slot = ecap + 2 * (rp0 + k);
          R.rgb = rgb | ((uint32_t)(pr.stipple != 0xFFFF) << 24);
          R.ne = 0; R.e0 = (uint16_t)slot; R.sgn = 1.0f; R.span0 = 0; R.ymin = 0.0f; R.ymax = 0.0f;
          if (L > 0.0f) {
            float ux = dx / L, uy = dy / L;
            float minx = fminf(a.x, b.x) - hw - 1.0f, maxx = fmaxf(a.x, b.x) + hw + 1.0f;
            float miny = fminf(a.y, b.y) - hw - 1.0f, maxy = fmaxf(a.y, b.y) + hw + 1.0f;
            int i0 = (int)fmaxf(0.0f, floorf(minx)), i1 = (int)fminf((float)(res_full - 1), ceilf(maxx));
            int j0 = (int)fmaxf(0.0f, floorf(miny)), j1 = (int)fminf((float)(res_full - 1), ceilf(maxy));
            R.col0 = (int16_t)i0; R.col1 = (int16_t)i1; R.row0 = (int16_t)j0;
            R.nrows = (int16_t)((j1 >= j0 && i1 >= i0) ? (j1 - j0 + 1) : 0);
            vs.edges[slot] = make_float4(a.x, a.y, ux, uy);
            vs.edges[slot + 1] = make_float4(L, s0, hw, __uint_as_float((uint32_t)pr.stipple));
          } else {
            R.col0 = 0; R.col1 = -1; R.row0 = 0; R.nrows = 0;
            vs.edges[slot] = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
            vs.edges[slot + 1] = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
          }
        }
        s0 += L;
      }
    } else if (rp0 < vs.rcap) {
      RPrim& R = vs.prims[rp0];
      float area2 = 0.0f;
      float l = INFINITY, r = -INFINITY, b = INFINITY, t = -INFINITY;
      if (pr.kind == MG_PRIM_NGON && n >= 20) {
        /* regular many-gon: rigid maps keep it counter-clockwise, and a slightly generous box from two
         * opposite vertices + the radius is enough (the box only limits where spans are searched) */
        float2 a = vs.verts[v0], c = vs.verts[v0 + n / 2];
        float cx = 0.5f * (a.x + c.x), cy = 0.5f * (a.y + c.y);
        float rad = pr.radius * cam.S * 1.001f + 0.05f;
        l = cx - rad; r = cx + rad; b = cy - rad; t = cy + rad;
        area2 = 1.0f;
      } else {
#pragma unroll 2
        for (int k = 0; k < n; k++) {
          float2 a = vs.verts[v0 + k], c = vs.verts[v0 + (k + 1 == n ? 0 : k + 1)];
          area2 += a.x * c.y - a.y * c.x;
          l = fminf(l, a.x); r = fmaxf(r, a.x); b = fminf(b, a.y); t = fmaxf(t, a.y);
        }
      }
      int i0 = (int)fmaxf(0.0f, floorf(l - 1.0f)), i1 = (int)fminf((float)(res_full - 1), ceilf(r + 1.0f));
      int j0 = (int)fmaxf(0.0f, floorf(b - 1.0f)), j1 = (int)fminf((float)(res_full - 1), ceilf(t + 1.0f));
      R.rgb = rgb;
      R.sgn = (area2 >= 0.0f) ? 1.0f : -1.0f;
      R.ymin = b; R.ymax = t;
      R.e0 = (uint16_t)v0; R.ne = (uint16_t)n; R.span0 = 0;
      R.col0 = (int16_t)i0; R.col1 = (int16_t)i1; R.row0 = (int16_t)j0;
      R.nrows = (int16_t)((j1 >= j0 && i1 >= i0) ? (j1 - j0 + 1) : 0);
    }
  }
  __syncthreads();
  RPROF(2);
  /* D: span-table offsets (warp 0) + oriented edge equations and, per polygon edge, the sample rows it
   * can bound: row j is bounded by an edge only if j + 0.5 lies within one sample of the edge's own
   * y-extent (the polygon is convex, so edges further away hold with a margin far above fp32 rounding) */
  if (tid < 32) {
    const int TSd = (res_out / RGRID) * SS; /* samples per tile side */
    int off = 0, toff = 0, nlines = 0;
    for (int base = 0; base < nrp; base += 32) {
      int p = base + tid;
      int nr = (p < nrp) ? vs.prims[p].nrows : 0;
      int r0 = (p < nrp) ? vs.prims[p].row0 : 0;
      const bool is_line = (p < nrp) && vs.prims[p].ne == 0 && nr > 0;
      /* tile rows the primitive's sample rows touch = its work items in the binning pass (F) */
      int nt_rows = (nr > 0) ? (min((r0 + nr - 1) / TSd, RGRID - 1) - r0 / TSd + 1) : 0;
      /* rows are allocated in aligned groups of 4 sample rows (one output row): the allocation starts at the
       * primitive's row0 rounded down to a multiple of 4, so that the shading pass fetches the four spans of an
       * output row with one 16-byte load; the up to 6 padding rows hold empty spans */
      const int alloc = nr > 0 ? (((r0 & 3) + nr + 3) & ~3) : 0;
      int incl = alloc, tincl = nt_rows;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d), tt = __shfl_up_sync(0xffffffffu, tincl, d);
        if (tid >= d) { incl += t; tincl += tt; }
      }
      if (p < nrp) {
        int start = off + incl - alloc;
        if (start + alloc > scap) { vs.prims[p].nrows = 0; nr = 0; } /* cannot happen with the host's bound */
        vs.prims[p].span0 = start + (nr > 0 ? (r0 & 3) : 0);
        vs.prims[p].tile0 = toff + tincl - nt_rows;
        /* the span phase works in units of 32 consecutive table rows: primitive of each unit's first row */
        for (int u = (start + 31) >> 5; u <= (start + alloc - 1) >> 5 && alloc > 0; u++)
          if (u < RLONG - RMAXLINES) s_off[u] = p;
      }
      /* thick line segments with rows: their (segment, row) items are spread over all warps in E0 */
      const unsigned lm = __ballot_sync(0xffffffffu, is_line && nr > 0);
      if (is_line && nr > 0) {
        const int k = nlines + __popc(lm & ((1u << tid) - 1u));
        if (k < RMAXLINES) s_off[RLONG - RMAXLINES + k] = p;
      }
      nlines += __popc(lm);
      off += __shfl_sync(0xffffffffu, incl, 31);
      toff += __shfl_sync(0xffffffffu, tincl, 31);
    }
    if (tid == 0) { s_misc[2] = off > scap ? scap : off; s_misc[6] = toff; s_misc[7] = nlines; }
  }
  for (int v = tid; v < nv; v += nt) {
    const int dp = ra.vprim[v];
    const mg_prim_t& pr = sc.prims[dp];
    int rp0 = ra.rp0[dp];
    if (pr.kind == MG_PRIM_LINELOOP || rp0 >= vs.rcap) {
      vs.eaux[v] = make_float4(0.0f, 0.0f, __int_as_float(0), __int_as_float(0));
      vs.edges[v] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
      continue;
    }
    int v0 = ra.voff[dp], n = pr.nvert, k = v - v0;
    float2 a = vs.verts[v0 + k], b = vs.verts[v0 + (k + 1 == n ? 0 : k + 1)];
    const RPrim& R = vs.prims[rp0];
    float sgn = R.sgn;
    float A = a.y - b.y, B = b.x - a.x;
    float C = -fmaf(A, a.x, B * a.y);
    A *= sgn; B *= sgn; C *= sgn;

    /* boundary column of this edge on row y: x*(y) = -(B y + C)/A = m y + q (an ESTIMATE only) */
    float inv = (A != 0.0f) ? __fdividef(1.0f, A) : 0.0f;
    /* rows with  ymin_e - 1 <= j + 0.5 <= ymax_e + 1  (decided with the exact fp32 comparisons), clipped
     * to the rows of the primitive that are not empty anyway */
    const float elo = fminf(a.y, b.y) - 1.0f, ehi = fmaxf(a.y, b.y) + 1.0f;
    int j0 = (int)ceilf(elo - 0.5f), j1 = (int)floorf(ehi - 0.5f);
    while ((float)(j0 - 1) + 0.5f >= elo) j0--;
    while ((float)j0 + 0.5f < elo) j0++;
    while ((float)(j1 + 1) + 0.5f <= ehi) j1++;
    while ((float)j1 + 0.5f > ehi) j1--;
    int r0 = R.row0, r1 = R.row0 + R.nrows - 1;
    while (r0 <= r1 && (float)r0 + 0.5f < R.ymin - 0.01f) r0++;
    while (r1 >= r0 && (float)r1 + 0.5f > R.ymax + 0.01f) r1--;
    j0 = max(j0, r0); j1 = min(j1, r1);
    const int cnt = j1 >= j0 ? j1 - j0 + 1 : 0;
    /* edges.w = primitive | rows << 8;  eaux.w = rows that enter the item queue (long edges only: the
     * short ones -- the sides of the many-gons -- are handled one edge per lane) */
    vs.edges[v] = make_float4(A, B, C, __int_as_float(rp0 | (cnt << 8)));
    vs.eaux[v] = make_float4(-B * inv, -C * inv, __int_as_float(j0), __int_as_float(cnt));
  }
  __syncthreads();
  RPROF(3);
  /* E: SPAN TABLE.  A thread owns one (primitive, sample row) and computes the row's exact covered interval
   * [lo, hi] on its own: it collects the edges of the primitive that can bound the row (row inside the edge's
   * [j0, j0 + cnt), one shared-memory word per edge, no divergence) into a bit mask and then evaluates its
   * candidates one per trip; every edge function fmaf(A, x, fmaf(B, y, C)) is monotone in x, so the covered set is
   * [max of the lower bounds, min of the upper bounds] -- the set a brute-force per-sample test of all edges
   * yields.  No atomics: nobody else writes the row.  Warps take primitives from a counter (painter's order:
   * the arena, the largest, first); thick line segments are solved directly per row (row_span), their rows
   * spread over all threads afterwards. */
  const int lane = tid & 31;
  {
    /* exact bound of edge (E, X) on sample row j within the primitive's columns [c0, c1]: first (A > 0) or last
     * (A < 0) column that holds; one search for both orientations: for A < 0 the column axis is mirrored */
    auto edge_bound = [&](const float4& E, const float4& X, int j, int c0, int c1) {
      const float A = E.x;
      const float y = (float)j + 0.5f;
      const float t = fmaf(E.y, y, E.z);
      const int sg = (A < 0.0f) ? -1 : 1;
      if (A != 0.0f) {
        auto ok = [&](int u) { return fmaf(A, (float)(sg * u) + 0.5f, t) >= 0.0f; };
        const float est = fmaf(X.x, y, X.y) - 0.5f;
        const int u = first_true(sg > 0 ? est : -est, sg > 0 ? c0 : -c1, sg > 0 ? c1 : -c0, ok);
        return sg * u;
      }
      return (t >= 0.0f) ? c1 : c0 - 1; /* horizontal edge: the whole row holds or none of it */
    };
    const int nlines = s_misc[7];
    const bool spread = nlines <= RMAXLINES;
    const int total_rows = s_misc[2];
    /* work unit = 32 consecutive rows of the span table (rows are laid out primitive after primitive), handed
     * out to warps from a counter: full lane use whatever the primitives' heights are; a unit may straddle two or
     * three primitives */
    for (int u = tid >> 5; u * 32 < total_rows; u += nt / 32) {
      const int w = u * 32 + lane;
      if (w >= total_rows) continue;
      /* primitive of the unit's first row (table of phase D); the lane then steps forward to the primitive
       * holding its own row (primitives without rows share offsets with their successor) */
      int p = (u < RLONG - RMAXLINES) ? s_off[u] : 0;
      while (p + 1 < nrp && vs.prims[p].span0 + vs.prims[p].nrows <= w) p++;
      const RPrim R = vs.prims[p];
      const int r = w - R.span0;
      if (r < 0 || r >= R.nrows) { /* padding row of the aligned allocation: nothing covered */
        vs.spans[w] = make_short2(1, 0);
        continue;
      }
      if (R.ne == 0) continue; /* thick line segments: below */
      const int c0 = R.col0, c1 = R.col1;
      {
        const int j = R.row0 + r;
        short2 sp;
        {
          const float y = (float)j + 0.5f;
          int lo = c0, hi = c1;
          if (y > R.ymax + 0.01f || y < R.ymin - 0.01f) {
            hi = c0 - 1; /* outside the vertices' y-extent: empty */
          } else {
            /* candidates in chunks of 32 edges (polygons have <= 8, the many-gons 10 / 20 / 100) */
            for (int kb = 0; kb < (int)R.ne; kb += 32) {
              const int kn = min(32, (int)R.ne - kb);
              uint32_t cand = 0u;
              for (int k = 0; k < kn; k++) {
                const float2 jc = *reinterpret_cast<const float2*>(&vs.eaux[R.e0 + kb + k].z); /* j0, cnt */
                cand |= ((unsigned)(j - __float_as_int(jc.x)) < (unsigned)__float_as_int(jc.y) ? 1u : 0u) << k;
              }
              while (cand) {
                const int k = __ffs(cand) - 1;
                cand &= cand - 1u;
                const float4 E = vs.edges[R.e0 + kb + k];
                const int v = edge_bound(E, vs.eaux[R.e0 + kb + k], j, c0, c1);
                if (E.x > 0.0f) lo = max(lo, v); else hi = min(hi, v);
              }
            }
          }
          sp = make_short2((short)lo, (short)hi);
        }
        vs.spans[R.span0 + r] = sp;
      }
    }
    if (nlines > 0) {
      /* (segment, row) items of the thick line segments, spread evenly over all threads.  The segments come from
       * the list phase D made; a scene with more than RMAXLINES of them walks all primitives instead */
      const int nl = spread ? nlines : nrp;
      auto seg = [&](int k) { return spread ? s_off[RLONG - RMAXLINES + k] : k; };
      auto seg_rows = [&](int k) { const RPrim& R = vs.prims[seg(k)]; return R.ne == 0 ? (int)R.nrows : 0; };
      int total = 0;
#pragma unroll 1
      for (int k = 0; k < nl; k++) total += seg_rows(k);
      for (int it = tid; it < total; it += nt) {
        int k = 0, r = it;
#pragma unroll 1
        for (;;) {
          const int nr = seg_rows(k);
          if (r < nr) break;
          r -= nr; k++;
        }
        const RPrim& R = vs.prims[seg(k)];
        vs.spans[R.span0 + r] = row_span(R, vs.edges, R.row0 + r);
      }
    }
  }
  __syncthreads();
  RPROF(5);
  /* F: tile bins from the spans: one work item per (primitive, tile row) */
  const int TS = (res_out / RGRID) * SS; /* samples per tile side */
  const int npairs = s_misc[6];
  for (int w = tid; w < npairs; w += nt) {
    /* the primitive owning item w: last one whose tile0 is <= w (primitives without rows share offsets) */
    int lo = 0, hi = nrp - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (vs.prims[mid].tile0 <= w) lo = mid; else hi = mid - 1; }
    const int p = lo;
    const RPrim& R = vs.prims[p];
    if (R.nrows == 0) continue;
    const int ty = (int)R.row0 / TS + (w - R.tile0);
    int ja = ty * TS, jb = ja + TS - 1;
    int a = max(ja, (int)R.row0), b = min(jb, R.row0 + R.nrows - 1);
    if (a > b) continue;
    int umin = 32767, umax = -1;   /* hull of the covered columns */
    int cmin = -1, cmax = 32767;   /* columns covered by EVERY row of the tile row */
    bool all_rows = (a == ja && b == jb);
#pragma unroll 2
    for (int j = a; j <= b; j++) {
      short2 s = vs.spans[R.span0 + (j - R.row0)];
      if (s.x <= s.y) {
        umin = min(umin, (int)s.x); umax = max(umax, (int)s.y);
        cmin = max(cmin, (int)s.x); cmax = min(cmax, (int)s.y);
      } else {
        all_rows = false;
      }
    }
    if (umax < umin) continue;
    const bool solid = (R.rgb >> 24) == 0; /* stippled segments never cover a tile completely */
    for (int tx = umin / TS; tx <= umax / TS && tx < RGRID; tx++) {
      atomicOr(&vs.tiles[(ty * RGRID + tx) * vs.rwords + (p >> 5)], 1u << (p & 31));
      if (all_rows && solid && cmin <= tx * TS && cmax >= tx * TS + TS - 1) atomicMax(&vs.cover[ty * RGRID + tx], p);
    }
  }
  __syncthreads();
  RPROF(6);
}

/* round-half-even mean of SSxSS samples (cv2 INTER_AREA: saturate_cast<uchar>(sum / 16.f)), red and blue
 * together: (s + 7 + (bit 4 of s)) >> 4 rounds s / 16 to nearest, ties to even, and never carries out of a
 * 12-bit field (s <= 4080) */
template <int SS>
__device__ __forceinline__ uint32_t finish_colour(uint32_t srb, uint32_t sg) {
  if (SS == 4) {
    srb = ((srb + 0x00070007u + ((srb >> 4) & 0x00010001u)) >> 4) & 0x00FF00FFu;
    sg = (sg + 7u + ((sg >> 4) & 1u)) >> 4;
  }
  return srb | (sg << 8);
}

/* Colours of the 4 output pixels (X0..X0+3, Yg): front-to-back walk over the tile's primitives.
 * `words` has a bit per non-empty word of the tile's primitive mask (above the covering primitive). */
template <int SS>
__device__ __forceinline__ void shade4(const ViewSmem& vs, int X0, int Yg, int tile, int cover, uint32_t words,
                                       float px_scale, uint32_t out[4]) {
  /* SS == 4: a pixel's 16 samples sit in a 32-bit mask as  row0 -> bits 0..3, row2 -> 4..7, row1 -> 16..19,
   * row3 -> 20..23  (two packed 16-sample strips per word, so a pixel's mask is two shift+and pairs) */
  constexpr uint32_t FULLM = (SS == 4) ? 0x00FF00FFu : 1u;
  uint32_t unres[4] = {FULLM, FULLM, FULLM, FULLM};
  /* colour sums of the 16 samples, two channels per word: red in bits 0..11, blue in bits 16..27 (a sum is at
   * most 16 * 255 < 4096), green on its own */
  uint32_t srb[4] = {0, 0, 0, 0}, sg[4] = {0, 0, 0, 0};
  const uint32_t* tm = vs.tiles + tile * vs.rwords;
  const int x0 = X0 * SS, y0 = Yg * SS;
  uint32_t any = FULLM;
  while (words && any) {
    const int w = 31 - __clz(words);
    words &= ~(1u << w);
    uint32_t bits = tm[w];
    if (cover >= 0 && (cover >> 5) == w) bits &= ~((1u << (cover & 31)) - 1u); /* nothing shows through it */
    while (bits && any) {
      int b = 31 - __clz(bits);
      bits &= ~(1u << b);
      const int p = w * 32 + b;
      const RPrim& R = vs.prims[p];
      RCOUNT(10, 1);
      uint32_t m[4];
      if (p == cover) {
#pragma unroll
        for (int i = 0; i < 4; i++) m[i] = unres[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) m[i] = 0u;
        uint32_t strips[SS]; /* per sample row: bit c = sample column x0 + c is covered */
#pragma unroll
        for (int r = 0; r < SS; r++) strips[r] = 0u;
        if (SS == 4) {
          /* the four sample rows of this output row are one aligned group of the primitive's allocation (phase D),
           * inside it or not at all; padding rows hold empty spans */
          if (y0 >= (R.row0 & ~3) && y0 < (((int)R.row0 + R.nrows + 3) & ~3)) {
            const uint4 q4 = *reinterpret_cast<const uint4*>(&vs.spans[R.span0 + (y0 - R.row0)]);
            const uint32_t qs[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int r = 0; r < SS; r++) {
              const int l = max((int)(short)(qs[r % 4] & 0xFFFFu) - x0, 0), h = min((int)(short)(qs[r % 4] >> 16) - x0, 4 * SS - 1);
              if (l <= h) strips[r] = (0xFFFFFFFFu >> (31 - (h - l))) << l;
            }
          }
        } else {
#pragma unroll
          for (int r = 0; r < SS; r++) {
            int row = y0 + r - R.row0;
            if (row >= 0 && row < R.nrows) {
              short2 sp = vs.spans[R.span0 + row];
              /* columns of the 4*SS-sample strip covered by this row */
              int l = max((int)sp.x - x0, 0), h = min((int)sp.y - x0, 4 * SS - 1);
              if (l <= h) strips[r] = (0xFFFFFFFFu >> (31 - (h - l))) << l;
            }
          }
        }
        if (SS == 4) {
          const uint32_t v01 = strips[0] | (strips[1 % SS] << 16), v23 = strips[2 % SS] | (strips[3 % SS] << 16);
#pragma unroll
          for (int i = 0; i < 4; i++)
            m[i] = ((v01 >> (4 * i)) & 0x000F000Fu) | (((v23 >> (4 * i)) & 0x000F000Fu) << 4);
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) m[i] = (strips[0] >> i) & 1u;
        }
        const bool stippled = (R.rgb >> 24) != 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          m[i] &= unres[i];
          if (stippled && m[i]) {
            /* stippled line: GL stipple bit from the distance along the loop, per covered sample */
            float4 pp = vs.edges[R.e0], q = vs.edges[R.e0 + 1];
            uint32_t stipple = __float_as_uint(q.w);
            uint32_t keep = 0u, mm = m[i];
            while (mm) {
              int s = __ffs(mm) - 1;
              mm &= mm - 1;
              /* bit -> (column, row) of the sample inside the pixel, see the mask layout above */
              const int sc_ = (SS == 4) ? (s & 3) : 0;
              const int sr_ = (SS == 4) ? (((s >> 4) & 1) + 2 * ((s >> 2) & 1)) : 0;
              float x = ((float)(x0 + SS * i + sc_)) + 0.5f, y = ((float)(y0 + sr_)) + 0.5f;
              float along = fmaf(x - pp.x, pp.z, (y - pp.y) * pp.w);
              int bit = ((int)floorf((q.y + along) / px_scale)) & 15;
              keep |= ((stipple >> bit) & 1u) << s;
            }
            m[i] = keep;
          }
        }
      }
      const uint32_t crb = R.rgb & 0x00FF00FFu, cg = (R.rgb >> 8) & 0xFF;
      any = 0u;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint32_t cnt = __popc(m[i]);
        srb[i] += cnt * crb; sg[i] += cnt * cg;
        unres[i] &= ~m[i];
        any |= unres[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t cnt = __popc(unres[i]);
    out[i] = finish_colour<SS>(srb[i] + cnt * (BG_R | (BG_B << 16)), sg[i] + cnt * BG_G);
  }
}

/* shift one pixel's 12 stack bytes left by one frame and append colour n (or replicate when fresh);
 * push == false (mg_render of an unchanged step): only the newest frame's 3 bytes are replaced */
__device__ __forceinline__ void stack_push(uint32_t* w, uint32_t n, bool fresh, bool push = true) {
  if (fresh) {
    w[0] = n | (n << 24);
    w[1] = (n >> 8) | (n << 16);
    w[2] = (n >> 16) | (n << 8);
  } else if (push) {
    uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    w[0] = (w0 >> 24) | (w1 << 8);
    w[1] = (w1 >> 24) | (w2 << 8);
    w[2] = (w2 >> 24) | (n << 8);
  } else {
    w[2] = (w[2] & 0xFFu) | (n << 8);
  }
}

/* ------------------------------------------------------------------ kernel
 * MODE: MG_OBS_*;  SS: samples per output pixel side (4 for the LoRes modes, 1 for RAW). */
template <int MODE>
__global__ void __launch_bounds__(RASTER_THREADS, RASTER_MIN_BLOCKS)
k_raster(EnvState* __restrict__ states, const DeviceScene* __restrict__ scenes, uint8_t* __restrict__ obs,
         uint8_t* __restrict__ newest, size_t plane_stride /* bytes between the two view planes of obs */, int batch,
         int res_out_arg, int ecap, int scap, int rcap, int only_fresh, int push, int env0, int slot_base) {
  /* the LoRes layouts are 96 x 96 by definition (benchmarks/__init__.py:242-274): a compile-time resolution
   * turns every tile / group / row division and the address arithmetic into constants */
  const int res_out = (MODE == MG_OBS_RAW) ? res_out_arg : 96;
  constexpr int SS = (MODE == MG_OBS_RAW) ? 1 : 4;
  /* The layouts with two views (LoResStack, RAW: one plane per view; LoRes3EA: both views inside one pixel) get
   * one CTA per (environment, view), so every view goes through the single-view code and shared-memory footprint
   * (blocks 2 e and 2 e + 1 serve environment e; `pass` 0 = allocentric, 1 = egocentric).  In LoRes3EA the two CTAs
   * of an environment write disjoint BYTES of the same pixels: the allo CTA bytes 0..2, the ego CTA bytes 3..11. */
  constexpr bool SEQ = (MODE == MG_OBS_LORESSTACK || MODE == MG_OBS_RAW || MODE == MG_OBS_LORES3EA);
  constexpr int NV = 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_off[RLONG]; /* long-edge list; last RMAXLINES entries: thick line segments; later the tile lists */
  __shared__ int s_misc[12];
  /* this launch covers environments [env0, env0 + gridDim.x) (SEQ: gridDim.x / 2) */
  const int env = env0 + (SEQ ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
  const int pass = SEQ ? (int)(blockIdx.x & 1) : 0; /* plane / view of this CTA in the two-plane layouts */
  if (env >= batch) return;
  EnvState& stg = states[env];
  const EnvState& st = stg;
  /* block-uniform.  only_fresh 1: after a reset only the reset envs are (re)drawn; 2: everything EXCEPT the envs
   * that just reset (their layout is still being sampled on another stream; they are drawn by a second launch) */
  if (only_fresh == 1 && st.fresh == 0) return;
  if (only_fresh == 2 && st.fresh != 0) return;
  /* device-side layout sampling (slot_base >= 0): the environment plays template st.scene; of everything the
   * render reads only the goal rectangles' vertices were re-drawn, and those live in the environment's slot */
  const int scene_index = st.scene;
  const mg_raster_aux_t* lv_own = slot_base >= 0 ? &scenes[slot_base + env].ra : nullptr;
  const mg_scene_t& sc = scenes[scene_index].s;
  const float px_scale = (float)(res_out * SS) / 384.0f;

  /* carve shared memory per view */
  ViewSmem vsm[NV];
  {
    unsigned char* p = smem_raw;
    for (int v = 0; v < NV; v++) {
      vsm[v].rcap = rcap; vsm[v].rwords = rcap >> 5;
      vsm[v].edges = reinterpret_cast<float4*>(p); p += sizeof(float4) * (size_t)(ecap + 2 * rcap);
      vsm[v].eaux = reinterpret_cast<float4*>(p); p += sizeof(float4) * (size_t)ecap;
      vsm[v].prims = reinterpret_cast<RPrim*>(p); p += sizeof(RPrim) * (size_t)rcap;
      vsm[v].spans = reinterpret_cast<short2*>(p);
      vsm[v].verts = reinterpret_cast<float2*>(p); /* scap * 4 >= ecap * 8 (host) */
      p += sizeof(short2) * (size_t)scap;
      vsm[v].tiles = reinterpret_cast<uint32_t*>(p); p += sizeof(uint32_t) * RGRID * RGRID * (size_t)(rcap >> 5);
      vsm[v].cover = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * RGRID * RGRID;
    }
  }
  const bool fresh = st.fresh != 0;
#ifndef RASTER_NO_PREFETCH
  /* The shading passes at the end shift this environment's frame stack (110 592 B) through registers; ask for it
   * now, so that it travels HBM -> L2 while the span tables are built (one bulk prefetch of 4 KB per thread of the
   * first warp; the 592 resident CTAs hold 65 MB of the 126 MB L2) */
  if (MODE != MG_OBS_RAW && !(MODE == MG_OBS_LORES3EA && pass == 0) && !fresh && push && threadIdx.x < 27) {
    const size_t plane = (MODE == MG_OBS_LORESSTACK) ? (size_t)pass * plane_stride : 0;
    const uint8_t* src = obs + plane + (size_t)env * (96 * 96 * 12) + (size_t)threadIdx.x * 4096;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(4096) : "memory");
  }
#endif
  {
  for (int v = 0; v < NV; v++) {
    int view = SEQ ? pass : ((MODE == MG_OBS_LORES4A) ? 0 : 1);
    build_view<SS>(vsm[v], st, sc, scenes[scene_index].ra, lv_own, view, res_out, ecap, scap, s_off, s_misc);
  }
  RPROF_DECL
  const int T = res_out / RGRID;        /* output pixels per tile side (multiple of 4) */
  const int gpr = T / 4;                /* 4-pixel groups per tile row */
  const int gpt = gpr * T;              /* groups per tile */
  const size_t frame_px = (size_t)res_out * res_out;
  /* G0: one thread per tile: the covering primitive, which mask words hold primitives above it, and whether
   * the whole tile is one flat colour (nothing above the cover / nothing at all) -> tile descriptor
   *   bits 0..8 cover + 1 | bits 9..16 non-empty mask words above the cover | bit 31 flat
   * and the tile goes on the flat list or the busy list.  The lists reuse s_off (dead after E1). */
  uint8_t* const flat_list = reinterpret_cast<uint8_t*>(s_off);
  uint8_t* const busy_list = flat_list + RGRID * RGRID;
  for (int tile = threadIdx.x; tile < RGRID * RGRID; tile += RASTER_THREADS) {
    bool all_flat = true;
    int weight = 0; /* primitives a pixel of this tile may have to walk */
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const int cover = vsm[v].cover[tile];
      const uint32_t* tm = vsm[v].tiles + tile * vsm[v].rwords;
      uint32_t wm = 0u, above = 0u;
      for (int w = 0; w < vsm[v].rwords; w++) {
        uint32_t bits = tm[w];
        if (cover >= 0) {
          if ((cover >> 5) > w) bits = 0u;
          else if ((cover >> 5) == w) bits &= ~((1u << (cover & 31)) - 1u);
        }
        if (bits) wm |= 1u << w;
        weight += __popc(bits);
        uint32_t others = bits;
        if (cover >= 0 && (cover >> 5) == w) others &= ~(1u << (cover & 31));
        above |= others;
      }
      vsm[v].cover[tile] = (int32_t)((uint32_t)(cover + 1) | (wm << 9) | (above == 0u ? 0x80000000u : 0u));
      all_flat = all_flat && (above == 0u);
    }
    /* busy tiles: the heavy ones (robot, piled-up blocks) fill the list from the front and are handed out
     * first, the light ones from the back: longest-processing-time-first keeps the tail of G2 short */
    if (all_flat) flat_list[atomicAdd(&s_misc[8], 1)] = (uint8_t)tile;
    else if (weight >= 5) busy_list[atomicAdd(&s_misc[9], 1)] = (uint8_t)tile;
    else busy_list[RGRID * RGRID - 1 - atomicAdd(&s_misc[10], 1)] = (uint8_t)tile;
  }
  __syncthreads();
  const int n_flat = s_misc[8], n_heavy = s_misc[9], n_busy = n_heavy + s_misc[10];
  RPROF(8);
  RCOUNT(8, threadIdx.x == 0 ? n_flat : 0); RCOUNT(9, threadIdx.x == 0 ? n_busy : 0);

  /* read the surviving frames of one 4-pixel group / shift, append and write it back */
  auto load_pre = [&](int X0, int Y, uint4 (&pre)[NV][3]) {
      /* issue the read of the surviving frames before shading so HBM latency overlaps the ALU work */
      if ((MODE == MG_OBS_LORES4E || MODE == MG_OBS_LORES4A || MODE == MG_OBS_LORESSTACK ||
           (MODE == MG_OBS_LORES3EA && pass == 1 && push)) &&
          !fresh) {
#pragma unroll
        for (int v = 0; v < 1; v++) {
          const size_t plane = (MODE == MG_OBS_LORESSTACK) ? (size_t)pass * plane_stride : 0;
          const uint4* ptr =
              reinterpret_cast<const uint4*>(obs + plane + ((size_t)env * frame_px + (size_t)Y * res_out + X0) * 12);
          /* streaming accesses: every byte of the stack is touched exactly once per step */
          pre[v][0] = __ldcs(ptr); pre[v][1] = __ldcs(ptr + 1); pre[v][2] = __ldcs(ptr + 2);
        }
      }
  };
  auto store_group = [&](int X0, int Y, const uint4 (&pre)[NV][3], const uint32_t (&col)[NV][4]) {
      if (MODE == MG_OBS_LORES4E || MODE == MG_OBS_LORES4A || MODE == MG_OBS_LORESSTACK) {
        /* [B, R, R, 12] (LoResStack: [2, B, R, R, 12]): 4 pixels = 48 bytes = 3 x uint4 */
#pragma unroll
        for (int v = 0; v < 1; v++) {
          const size_t plane = (MODE == MG_OBS_LORESSTACK) ? (size_t)pass * plane_stride : 0;
          uint4* ptr = reinterpret_cast<uint4*>(obs + plane + ((size_t)env * frame_px + (size_t)Y * res_out + X0) * 12);
          uint32_t w[12];
          if (!fresh) {
            uint4 a = pre[v][0], b = pre[v][1], c = pre[v][2];
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) stack_push(&w[3 * i], col[v][i], fresh, push != 0);
          __stcs(ptr, make_uint4(w[0], w[1], w[2], w[3]));
          __stcs(ptr + 1, make_uint4(w[4], w[5], w[6], w[7]));
          __stcs(ptr + 2, make_uint4(w[8], w[9], w[10], w[11]));
          if (newest) {
            /* newest frame alone, [views][batch][R][R][3]: the send buffer of the multi-GPU observation
             * gather (27 648 B per environment instead of the 110 592 B stack) */
            uint32_t* np_ = reinterpret_cast<uint32_t*>(
                newest + (((MODE == MG_OBS_LORESSTACK ? (size_t)pass * batch : 0) + env) * frame_px + (size_t)Y * res_out + X0) * 3);
            const uint32_t c0 = col[v][0], c1 = col[v][1], c2 = col[v][2], c3 = col[v][3];
            __stcs(np_, c0 | (c1 << 24));
            __stcs(np_ + 1, (c1 >> 8) | (c2 << 16));
            __stcs(np_ + 2, (c2 >> 16) | (c3 << 8));
          }
        }
      } else if (MODE == MG_OBS_LORES3EA) {
        /* bytes 0..2 = newest allo frame; bytes 3..11 = 3 ego frames, oldest first.  This CTA owns one of the two
         * byte ranges of every pixel (the sibling CTA writes the other, concurrently), so it stores exactly its
         * bytes: the allo CTA 2 + 1 bytes per pixel, the ego CTA 1 + 4 + 4 (or only the newest 1 + 2 on a redraw) */
        uint8_t* px = obs + ((size_t)env * frame_px + (size_t)Y * res_out + X0) * 12;
        if (pass == 0) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t al = col[0][i];
            *reinterpret_cast<uint16_t*>(px + 12 * i) = (uint16_t)(al & 0xFFFFu);
            px[12 * i + 2] = (uint8_t)(al >> 16);
          }
        } else {
          uint32_t w[12];
          if (!fresh && push) {
            uint4 a = pre[0][0], b = pre[0][1], c = pre[0][2];
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t eg = col[0][i];
            if (fresh) {
              px[12 * i + 3] = (uint8_t)(eg & 0xFFu);
              *reinterpret_cast<uint32_t*>(px + 12 * i + 4) = (eg >> 8) | (eg << 16);
              *reinterpret_cast<uint32_t*>(px + 12 * i + 8) = (eg >> 16) | (eg << 8);
            } else if (!push) {
              px[12 * i + 9] = (uint8_t)(eg & 0xFFu);
              *reinterpret_cast<uint16_t*>(px + 12 * i + 10) = (uint16_t)(eg >> 8);
            } else {
              const uint32_t w1 = w[3 * i + 1], w2 = w[3 * i + 2];
              px[12 * i + 3] = (uint8_t)((w1 >> 16) & 0xFFu);  /* old bytes 6..11 -> 3..8 ; new ego -> 9..11 */
              *reinterpret_cast<uint32_t*>(px + 12 * i + 4) = (w1 >> 24) | (w2 << 8);
              *reinterpret_cast<uint32_t*>(px + 12 * i + 8) = (w2 >> 24) | (eg << 8);
            }
          }
        }
      } else if (MODE == MG_OBS_LORESCHW4E) {
        /* [B, 12, R, R]: plane c of frame f is channel 3f + c; 4 pixels = one u32 per plane */
        uint8_t* base = obs + (size_t)env * 12 * frame_px + (size_t)Y * res_out + X0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          uint32_t nw = ((col[0][0] >> (8 * c)) & 0xFF) | (((col[0][1] >> (8 * c)) & 0xFF) << 8) |
                        (((col[0][2] >> (8 * c)) & 0xFF) << 16) | (((col[0][3] >> (8 * c)) & 0xFF) << 24);
          uint32_t* p0 = reinterpret_cast<uint32_t*>(base + (size_t)(0 + c) * frame_px);
          uint32_t* p1 = reinterpret_cast<uint32_t*>(base + (size_t)(3 + c) * frame_px);
          uint32_t* p2 = reinterpret_cast<uint32_t*>(base + (size_t)(6 + c) * frame_px);
          uint32_t* p3 = reinterpret_cast<uint32_t*>(base + (size_t)(9 + c) * frame_px);
          if (fresh) { *p0 = nw; *p1 = nw; *p2 = nw; *p3 = nw; }
          else if (!push) { *p3 = nw; }
          else { uint32_t a = *p1, b = *p2, d = *p3; *p0 = a; *p1 = b; *p2 = d; *p3 = nw; }
        }
      } else {
        /* RAW [2, B, R, R, 3]: 4 pixels = 12 bytes = 3 x u32 */
#pragma unroll
        for (int v = 0; v < 1; v++) {
          uint32_t* ptr = reinterpret_cast<uint32_t*>(
              obs + (size_t)pass * plane_stride + ((size_t)env * frame_px + (size_t)Y * res_out + X0) * 3);
          uint32_t c0 = col[v][0], c1 = col[v][1], c2 = col[v][2], c3 = col[v][3];
          ptr[0] = c0 | (c1 << 24);
          ptr[1] = (c1 >> 8) | (c2 << 16);
          ptr[2] = (c2 >> 16) | (c3 << 8);
        }
      }
  };

  /* G1: flat tiles, all threads converged: item = (flat tile, 4-pixel group).  The loads of the next item are
   * issued before the current one is written back (two groups in flight per thread). */
  {
    const int n_items = n_flat * gpt;
    int item = threadIdx.x;
    uint4 pre_next[NV][3];
    auto item_xy = [&](int it, int& X0, int& Y, int& tile) {
      tile = flat_list[it / gpt];
      const int g = it % gpt;
      const int tx = tile % RGRID, ty = tile / RGRID; /* ty counts GL rows (bottom-up) */
      X0 = tx * T + (g % gpr) * 4;
      Y = res_out - 1 - (ty * T + g / gpr);
    };
    int X0 = 0, Y = 0, tile = 0;
    if (item < n_items) { item_xy(item, X0, Y, tile); load_pre(X0, Y, pre_next); }
    while (item < n_items) {
      uint4 pre[NV][3];
#pragma unroll
      for (int v = 0; v < NV; v++)
#pragma unroll
        for (int k = 0; k < 3; k++) pre[v][k] = pre_next[v][k];
      const int cX0 = X0, cY = Y, ctile = tile;
      const int next = item + RASTER_THREADS;
      if (next < n_items) { item_xy(next, X0, Y, tile); load_pre(X0, Y, pre_next); }
      uint32_t col[NV][4];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const int cover = (int)((uint32_t)vsm[v].cover[ctile] & 0x1FFu) - 1;
        const uint32_t c = (cover >= 0) ? (vsm[v].prims[cover].rgb & 0xFFFFFFu) : (BG_R | (BG_G << 8) | (BG_B << 16));
#pragma unroll
        for (int i = 0; i < 4; i++) col[v][i] = c;
      }
      store_group(cX0, cY, pre, col);
      item = next;
    }
  }
  /* no barrier: a warp that runs out of flat items starts on the busy tiles (own hand-out counter) */
  RPROF(9);

  /* G2: busy tiles, handed out dynamically to half-warps: all 16 lanes walk the same primitive list */
  const int hl = threadIdx.x & 15;
  const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
  for (;;) {
    int bi = 0;
    if (hl == 0) bi = atomicAdd(&s_misc[11], 1);
    bi = __shfl_sync(hmask, bi, 0, 16);
    if (bi >= n_busy) break;
    const int tile = busy_list[bi < n_heavy ? bi : RGRID * RGRID - 1 - (bi - n_heavy)];
    const int tx = tile % RGRID, ty = tile / RGRID;
    int cover[NV];
    uint32_t words[NV];
    bool flat[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const uint32_t d = (uint32_t)vsm[v].cover[tile];
      cover[v] = (int)(d & 0x1FFu) - 1;
      words[v] = (d >> 9) & 0xFFu;
      flat[v] = (d >> 31) != 0u;
    }
    for (int g = hl; g < gpt; g += 16) {
      const int Yg = ty * T + g / gpr;
      const int X0 = tx * T + (g % gpr) * 4;
      const int Y = res_out - 1 - Yg;       /* output row, 0 = top */
      /* issue the read of the surviving frames before shading so HBM latency overlaps the ALU work */
      uint4 pre[NV][3];
      load_pre(X0, Y, pre);
      uint32_t col[NV][4];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        if (flat[v]) {
          const uint32_t c = (cover[v] >= 0) ? (vsm[v].prims[cover[v]].rgb & 0xFFFFFFu) : (BG_R | (BG_G << 8) | (BG_B << 16));
#pragma unroll
          for (int i = 0; i < 4; i++) col[v][i] = c;
        } else {
          shade4<SS>(vsm[v], X0, Yg, tile, cover[v], words[v], px_scale, col[v]);
        }
      }
      store_group(X0, Y, pre, col);
    }
  }
  RPROF(7);
  }
  if (threadIdx.x == 0 && fresh) {
    if (!SEQ) {
      stg.fresh = 0;
    } else {
      /* two CTAs per environment: each leaves its mark in the flag, the second one to finish clears it (the flag
       * stays non-zero for a sibling that starts late) */
      const int old = atomicOr(&stg.fresh, 4 << pass);
      if (old & (4 << (1 - pass))) stg.fresh = 0;
    }
  }
}

static int n_views(int mode) { /* views resident in one CTA's shared memory */
  (void)mode;
  return 1;
}

size_t mg_raster_smem_bytes(int mode, int ecap, int scap, int rcap) {
  size_t per_view = sizeof(float4) * (size_t)(ecap + 2 * rcap) + sizeof(float4) * (size_t)ecap +
                    sizeof(RPrim) * (size_t)rcap +
                    sizeof(short2) * (size_t)scap + sizeof(uint32_t) * RGRID * RGRID * (size_t)(rcap >> 5) +
                    sizeof(int32_t) * RGRID * RGRID;
  return per_view * n_views(mode);
}

#ifdef RASTER_PROF
extern "C" void mg_raster_prof_read(unsigned long long* out16, int reset) {
  cudaMemcpyFromSymbol(out16, g_rprof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_rprof, z, sizeof(z)); }
}
#endif

cudaError_t mg_raster_upload_units(const double* units /* [130][2] */) {
  return cudaMemcpyToSymbol(c_unit, units, sizeof(double) * 130 * 2);
}

template <int MODE>
static cudaError_t launch_mode(EnvState* states, const DeviceScene* scenes, uint8_t* obs, uint8_t* newest,
                               size_t plane_stride, int batch, int res_out, int ecap, int scap, int rcap,
                               int only_fresh, int push, int env0, int count, int slot_base, cudaStream_t stream) {
  size_t smem = mg_raster_smem_bytes(MODE, ecap, scap, rcap);
  /* function attributes are per device and per instantiation: set them when the footprint differs from what this
   * device's k_raster<MODE> was last configured for (handles with different scenes may alternate) */
  static size_t configured_on[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (configured_on[dev & 63] != smem) {
    e = cudaFuncSetAttribute(k_raster<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    /* shared-memory carve-out: exactly what the resident CTAs need (registers allow 4 per SM), the rest of the
     * 256 KB stays L1 for the scene tables every CTA of the SM reads */
    const size_t per_cta = smem + 3200 /* static */ + 1024 /* reserved per CTA */;
    size_t ctas = (size_t)(227 * 1024) / per_cta;
    if (ctas > 4) ctas = 4;
    if (ctas < 1) ctas = 1;
    int pct = (int)((ctas * per_cta * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    if (const char* ev = getenv("MG_RASTER_CARVEOUT")) pct = atoi(ev);
    e = cudaFuncSetAttribute(k_raster<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
    configured_on[dev & 63] = smem;
  }
  constexpr int ctas_per_env = (MODE == MG_OBS_LORESSTACK || MODE == MG_OBS_RAW || MODE == MG_OBS_LORES3EA) ? 2 : 1;
  k_raster<MODE><<<count * ctas_per_env, RASTER_THREADS, smem, stream>>>(states, scenes, obs, newest, plane_stride, batch, res_out, ecap,
                                                          scap, rcap, only_fresh, push, env0, slot_base);
  return cudaGetLastError();
}

cudaError_t mg_launch_raster(int mode, EnvState* states, const DeviceScene* scenes, uint8_t* obs, uint8_t* newest,
                             size_t plane_stride, int batch, int res_out, int ecap, int scap, int rcap, int only_fresh,
                             int push, int env0, int count, int slot_base, cudaStream_t stream) {
#define MG_RASTER_CASE(M)                                                                                            \
  case M:                                                                                                            \
    return launch_mode<M>(states, scenes, obs, newest, plane_stride, batch, res_out, ecap, scap, rcap, only_fresh, push, \
                          env0, count, slot_base, stream);
  switch (mode) {
    MG_RASTER_CASE(MG_OBS_LORES4E)
    MG_RASTER_CASE(MG_OBS_LORES4A)
    MG_RASTER_CASE(MG_OBS_LORES3EA)
    MG_RASTER_CASE(MG_OBS_LORESSTACK)
    MG_RASTER_CASE(MG_OBS_LORESCHW4E)
    MG_RASTER_CASE(MG_OBS_RAW)
  }
#undef MG_RASTER_CASE
  return cudaErrorInvalidValue;
}
