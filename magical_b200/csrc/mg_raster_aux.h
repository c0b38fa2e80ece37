/*
 * mg_raster_aux.h — per-scene STATIC tables of the rasteriser, derived on the host once per compiled scene:
 * which draw primitive every window-space vertex belongs to, the vertex / window-primitive offsets of the
 * draw list, the local-space coordinates of every vertex (for the 10/20/100-gons of gym_render.make_circle,
 * gym_render.py:438-446, the unit circle times the radius) and the list of polygon edges that can be long
 * enough to enter the (edge, row) item queue.  None of this depends on the simulator state, so computing it
 * per environment and step (prefix sums, binary searches, fp64 multiplies) was wasted work.
 */
#ifndef MG_RASTER_AUX_H
#define MG_RASTER_AUX_H

#include <math.h>
#include <stdint.h>

#include "../../include/magical_b200.h"

#define MG_RV_MAX 3072   /* window-space vertices of one scene (= polygon edges + line-loop vertices) */
#define MG_RLONG 768     /* edges that may own (edge, row) items */
#define MG_RSHORT 12     /* polygon edges bounding fewer sample rows than this are processed one edge per lane */

typedef struct {
  int32_t nv, nrp, n_cand, n_lines;
  float lv[MG_RV_MAX][2];           /* local-space vertex (NGON: (float)(unit * radius), before centre / pupil) */
  uint8_t vprim[MG_RV_MAX];         /* draw primitive of vertex v */
  uint8_t vcand[MG_RV_MAX];         /* 1: v is on the candidate list (its edge may be long) */
  uint16_t voff[MG_MAX_PRIMS + 2];  /* first vertex of primitive p; voff[n_prims] = nv */
  uint16_t rp0[MG_MAX_PRIMS + 2];   /* first window-space primitive of p (a line loop expands to one per segment) */
  uint16_t cand[MG_RLONG];          /* vertices whose edge may bound >= MG_RSHORT rows, ascending */
} mg_raster_aux_t;

/* returns NULL or why the scene cannot be rasterised */
static inline const char* mg_build_raster_aux(const mg_scene_t* s, mg_raster_aux_t* ra) {
  memset(ra, 0, sizeof(*ra));
  int nv = 0, nrp = 0, nc = 0, nl = 0;
  const double S384 = 384.0 / 2.04;
  for (int p = 0; p < s->n_prims; p++) {
    const mg_prim_t* pr = &s->prims[p];
    const int n = pr->nvert;
    if (nv + n > MG_RV_MAX) return "too many draw vertices in one scene";
    ra->voff[p] = (uint16_t)nv;
    ra->rp0[p] = (uint16_t)nrp;
    int may_be_long = 1;
    if (pr->kind == MG_PRIM_NGON) {
      if (n != 10 && n != 20 && n != 100) return "NGON primitives must have 10, 20 or 100 sides";
      /* rows an edge can bound <= its length in samples + 3 (margins); a rigid map keeps the length */
      const double edge_len = 2.0 * (double)pr->radius * sin(M_PI / n) * S384;
      may_be_long = (edge_len + 4.0 >= MG_RSHORT);
    } else if ((int)pr->vert0 + n > MG_MAX_DVERTS) {
      return "draw vertex range";
    }
    for (int k = 0; k < n; k++) {
      if (pr->kind == MG_PRIM_NGON) {
        const double ang = 2 * M_PI * k / n; /* gym_render.make_circle */
        const double ux = cos(ang), uy = sin(ang), r = (double)pr->radius;
        ra->lv[nv][0] = (float)(ux * r);
        ra->lv[nv][1] = (float)(uy * r);
      } else {
        ra->lv[nv][0] = s->dverts[pr->vert0 + k][0];
        ra->lv[nv][1] = s->dverts[pr->vert0 + k][1];
      }
      ra->vprim[nv] = (uint8_t)p;
      if (pr->kind != MG_PRIM_LINELOOP && may_be_long) {
        if (nc >= MG_RLONG) return "too many polygon edges in one scene";
        ra->cand[nc++] = (uint16_t)nv;
        ra->vcand[nv] = 1;
      }
      nv++;
    }
    if (pr->kind == MG_PRIM_LINELOOP) { nrp += n; nl += n; } else nrp += 1;
  }
  ra->voff[s->n_prims] = (uint16_t)nv;
  ra->rp0[s->n_prims] = (uint16_t)nrp;
  ra->nv = nv; ra->nrp = nrp; ra->n_cand = nc; ra->n_lines = nl;
  return NULL;
}

#endif
