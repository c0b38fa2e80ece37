/*
 * mgo_render.c — CPU ORACLE rasteriser (test infrastructure).
 *
 * Brute-force restatement of what the reference obtains from pyglet/OpenGL in
 * `BaseEnv.render` (magical/base_env.py:309-338) -> `Viewer.render`
 * (magical/gym_render.py:208-249): clear to the background colour, draw every
 * geom in insertion order (painter's algorithm) under the allocentric
 * (`set_bounds`, gym_render.py:176-182) or egocentric (`set_cam_follow` /
 * `TransformEgocentric`, gym_render.py:184-200, 362-380) camera, read back RGB
 * with row 0 = top.  Then `cv2.resize(..., INTER_AREA)` 384->96 as used by the
 * LoRes* preprocessors (benchmarks/__init__.py:159-169, 234) = exact 4x4 box
 * mean with round-half-to-even (checked against real cv2 in tests).
 *
 * PARITY UNPINNED w.r.t. a real GL driver (none available offline).  The
 * rasterisation rules restated here are the OpenGL ones for the state the
 * reference sets: no MSAA (msaa_samples=1, gym_render.py:150-151), convex
 * polygon fill by pixel-centre sampling, fp32 vertex transforms.  Documented
 * simplifications: line primitives (arena border, dashed goal
 * borders) are rendered as opaque width-w rectangles around each segment
 * instead of GL_LINE_SMOOTH alpha coverage, with the GL stipple counter
 * advancing one bit per pixel of length along the loop.
 *
 * This file evaluates EVERY sample of the full-resolution frame against every
 * primitive; the CUDA rasteriser gets the same bits through a hierarchical
 * tile/pixel/sample traversal.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mgo.h"

#define ZOOM 1.02 /* style.ARENA_ZOOM_OUT */

typedef struct { float x, y; } f2;

typedef struct {
  /* world -> GL window coordinates: q = S * (Rc * (p - c) + newpos) */
  float cc, cs; /* camera rotation (cos, sin of robot angle; 1, 0 for allo) */
  float cx, cy; /* camera centre */
  float nx, ny; /* new position of the centre, world units */
  float S;      /* pixels per world unit */
} camera;

static camera make_camera(const mgo_env* e, int view, int res) {
  camera cam;
  cam.S = (float)((double)res / (2.0 * ZOOM));
  if (view == 0) {
    cam.cc = 1.0f; cam.cs = 0.0f; cam.cx = 0.0f; cam.cy = 0.0f;
    cam.nx = (float)ZOOM; cam.ny = (float)ZOOM;
  } else {
    const mgo_body* r = &e->bodies[e->scene.robot_body];
    cam.cc = (float)r->rot.x; cam.cs = (float)r->rot.y;
    cam.cx = (float)r->p.x; cam.cy = (float)r->p.y;
    /* target (0.5, 0.15) of a 2.04 x 2.04 viewport (base_env.py:294-301) */
    cam.nx = (float)(2.0 * ZOOM * 0.5); cam.ny = (float)(2.0 * ZOOM * 0.15);
  }
  return cam;
}
static f2 world_to_px(const camera* cam, float wx, float wy) {
  float dx = wx - cam->cx, dy = wy - cam->cy;
  /* rotate by -theta */
  float rx = fmaf(cam->cc, dx, cam->cs * dy);
  float ry = fmaf(cam->cc, dy, -(cam->cs * dx));
  f2 q = {(rx + cam->nx) * cam->S, (ry + cam->ny) * cam->S};
  return q;
}
static f2 body_to_world(const mgo_body* b, float vx, float vy) {
  float bc = (float)b->rot.x, bs = (float)b->rot.y, bx = (float)b->p.x, by = (float)b->p.y;
  f2 w = {fmaf(bc, vx, fmaf(-bs, vy, bx)), fmaf(bs, vx, fmaf(bc, vy, by))};
  return w;
}

static double unit_table[101][2];
static int unit_n[3] = {10, 20, 100};
static int unit_off[3] = {0, 10, 30};
static double unit_all[130][2];
static int unit_ready = 0;
static void init_units(void) {
  if (unit_ready) return;
  for (int t = 0; t < 3; t++)
    for (int k = 0; k < unit_n[t]; k++) {
      double ang = 2 * M_PI * k / unit_n[t]; /* gym_render.make_circle, gym_render.py:438-446 */
      unit_all[unit_off[t] + k][0] = cos(ang);
      unit_all[unit_off[t] + k][1] = sin(ang);
    }
  (void)unit_table;
  unit_ready = 1;
}
static const double (*units_for(int n))[2] {
  for (int t = 0; t < 3; t++) if (unit_n[t] == n) return &unit_all[unit_off[t]];
  return NULL;
}

/* pixel-space vertices of primitive `pr` under `cam`; returns the count */
static int prim_vertices(const mgo_env* e, const mg_prim_t* pr, const camera* cam, f2* out) {
  const mg_scene_t* s = &e->scene;
  int n = pr->nvert;
  if (pr->kind == MG_PRIM_NGON) {
    const double(*U)[2] = units_for(n);
    const mgo_body* b = &e->bodies[pr->body];
    double r = (double)pr->radius;
    float pc = 1.0f, ps = 0.0f;
    if (pr->xform == MG_XFORM_PUPIL) {
      /* pupil spins by (eye angle - robot angle) about the eye centre (entities.py:488-490) */
      const mgo_body* eye = &e->bodies[pr->body2];
      pc = (float)(eye->rot.x * b->rot.x + eye->rot.y * b->rot.y);
      ps = (float)(eye->rot.y * b->rot.x - eye->rot.x * b->rot.y);
    }
    for (int k = 0; k < n; k++) {
      float vx = (float)(U[k][0] * r), vy = (float)(U[k][1] * r);
      if (pr->xform == MG_XFORM_PUPIL) {
        float ux = vx + pr->ex, uy = vy + pr->ey;
        vx = fmaf(pc, ux, -(ps * uy));
        vy = fmaf(ps, ux, pc * uy);
      }
      vx += pr->cx; vy += pr->cy;
      f2 w = body_to_world(b, vx, vy);
      out[k] = world_to_px(cam, w.x, w.y);
    }
    return n;
  }
  for (int k = 0; k < n; k++) {
    float vx = s->dverts[pr->vert0 + k][0], vy = s->dverts[pr->vert0 + k][1];
    f2 w = {vx, vy};
    if (pr->xform == MG_XFORM_BODY) w = body_to_world(&e->bodies[pr->body], vx, vy);
    out[k] = world_to_px(cam, w.x, w.y);
  }
  return n;
}

typedef struct { float A, B, C; } edge;
static edge make_edge(f2 a, f2 b) {
  edge ed;
  ed.A = a.y - b.y;
  ed.B = b.x - a.x;
  ed.C = -fmaf(ed.A, a.x, ed.B * a.y);
  return ed;
}
static inline float edge_eval(edge ed, float x, float y) { return fmaf(ed.A, x, fmaf(ed.B, y, ed.C)); }

static void fill_poly(uint8_t* img, int res, const f2* v, int n, const uint8_t rgb[3]) {
  /* orientation from the signed area (vertex lists come in both windings) */
  float area2 = 0.0f;
  for (int k = 0; k < n; k++) {
    f2 a = v[k], b = v[(k + 1) % n];
    area2 += a.x * b.y - a.y * b.x;
  }
  float sgn = area2 >= 0.0f ? 1.0f : -1.0f;
  edge ed[128];
  float minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
  for (int k = 0; k < n; k++) {
    ed[k] = make_edge(v[k], v[(k + 1) % n]);
    ed[k].A *= sgn; ed[k].B *= sgn; ed[k].C *= sgn;
    minx = fminf(minx, v[k].x); maxx = fmaxf(maxx, v[k].x);
    miny = fminf(miny, v[k].y); maxy = fmaxf(maxy, v[k].y);
  }
  int i0 = (int)fmaxf(0.0f, floorf(minx - 1.0f)), i1 = (int)fminf((float)(res - 1), ceilf(maxx + 1.0f));
  int j0 = (int)fmaxf(0.0f, floorf(miny - 1.0f)), j1 = (int)fminf((float)(res - 1), ceilf(maxy + 1.0f));
  for (int j = j0; j <= j1; j++) {
    float y = (float)j + 0.5f;
    for (int i = i0; i <= i1; i++) {
      float x = (float)i + 0.5f;
      int inside = 1;
      for (int k = 0; k < n && inside; k++) inside = edge_eval(ed[k], x, y) >= 0.0f;
      if (inside) {
        uint8_t* px = img + ((size_t)(res - 1 - j) * res + i) * 3;
        px[0] = rgb[0]; px[1] = rgb[1]; px[2] = rgb[2];
      }
    }
  }
}

static void draw_lineloop(uint8_t* img, int res, const f2* v, int n, float width, unsigned stipple, const uint8_t rgb[3],
                          float px_scale) {
  /* width is given in pixels of the 384-res frame; scale for other resolutions */
  float hw = 0.5f * width * px_scale;
  float s0 = 0.0f;
  for (int k = 0; k < n; k++) {
    f2 a = v[k], b = v[(k + 1) % n];
    float dx = b.x - a.x, dy = b.y - a.y;
    float L = sqrtf(fmaf(dx, dx, dy * dy));
    if (L > 0.0f) {
      float ux = dx / L, uy = dy / L;
      float minx = fminf(a.x, b.x) - hw - 1.0f, maxx = fmaxf(a.x, b.x) + hw + 1.0f;
      float miny = fminf(a.y, b.y) - hw - 1.0f, maxy = fmaxf(a.y, b.y) + hw + 1.0f;
      int i0 = (int)fmaxf(0.0f, floorf(minx)), i1 = (int)fminf((float)(res - 1), ceilf(maxx));
      int j0 = (int)fmaxf(0.0f, floorf(miny)), j1 = (int)fminf((float)(res - 1), ceilf(maxy));
      for (int j = j0; j <= j1; j++) {
        float y = (float)j + 0.5f;
        for (int i = i0; i <= i1; i++) {
          float x = (float)i + 0.5f;
          float rx = x - a.x, ry = y - a.y;
          float along = fmaf(rx, ux, ry * uy);
          float perp = fmaf(rx, uy, -(ry * ux));
          if (along < 0.0f || along > L || fabsf(perp) > hw) continue;
          int bit = ((int)floorf((s0 + along) / px_scale)) & 15;
          if (!((stipple >> bit) & 1u)) continue;
          uint8_t* px = img + ((size_t)(res - 1 - j) * res + i) * 3;
          px[0] = rgb[0]; px[1] = rgb[1]; px[2] = rgb[2];
        }
      }
    }
    s0 += L;
  }
}

void mgo_render_view(const mgo_env* e, int view, int res, uint8_t* out) {
  init_units();
  const mg_scene_t* s = &e->scene;
  camera cam = make_camera(e, view, res);
  /* background = lighten(grey, 4) (base_env.py:186) */
  for (size_t i = 0; i < (size_t)res * res; i++) { out[3 * i] = 231; out[3 * i + 1] = 231; out[3 * i + 2] = 234; }
  f2 verts[128];
  for (int p = 0; p < s->n_prims; p++) {
    const mg_prim_t* pr = &s->prims[p];
    int n = prim_vertices(e, pr, &cam, verts);
    if (pr->kind == MG_PRIM_LINELOOP) draw_lineloop(out, res, verts, n, pr->radius, pr->stipple, pr->rgb, (float)res / 384.0f);
    else fill_poly(out, res, verts, n, pr->rgb);
  }
}

void mgo_downsample4(const uint8_t* src, int n_out, uint8_t* dst) {
  int res = n_out * 4;
  for (int y = 0; y < n_out; y++)
    for (int x = 0; x < n_out; x++)
      for (int c = 0; c < 3; c++) {
        int sum = 0;
        for (int dy = 0; dy < 4; dy++)
          for (int dx = 0; dx < 4; dx++) sum += src[((size_t)(4 * y + dy) * res + (4 * x + dx)) * 3 + c];
        /* saturate_cast<uchar>(sum * (1/16.f)) = round half to even */
        int q = sum >> 4, rem = sum & 15;
        if (rem > 8 || (rem == 8 && (q & 1))) q++;
        dst[((size_t)y * n_out + x) * 3 + c] = (uint8_t)q;
      }
}
