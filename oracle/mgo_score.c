/*
 * mgo_score.c — CPU ORACLE end-of-trajectory scores (test infrastructure).
 *
 * Restates `score_on_end_of_traj` of the seven task files under
 * /root/reference/magical/benchmarks/ and `GoalRegion.get_overlapping_ents`
 * (magical/entities.py:821-881) in fp64.  Each function cites the lines it
 * follows.  Geometric queries go through the oracle's own narrowphase
 * (mgo_collide), as pymunk's `space.shape_query` does through cpShapesCollide.
 */
#include <math.h>
#include <string.h>

#include "mgo.h"

/* Sensor box of goal `g` as a temporary poly shape in the scratch slot
 * (Poly.create_box vertex order; entities.py:794-797). */
static void load_goal_shape(mgo_env* e, int g, int slot) {
  const mg_goal_t* goal = &e->scene.goals[g];
  mgo_shape* s = &e->shapes[slot];
  memset(s, 0, sizeof(*s));
  s->kind = MG_SHAPE_POLY;
  s->body = -2; /* its own static body */
  s->nvert = 4;
  double hw = goal->w / 2, hh = goal->h / 2;
  double lx[4] = {hw, hw, -hw, -hw}, ly[4] = {-hh, hh, hh, -hh};
  static const double nx[4] = {0, 1, 0, -1}, ny[4] = {-1, 0, 1, 0};
  double l = INFINITY, r = -INFINITY, b = INFINITY, t = -INFINITY;
  for (int k = 0; k < 4; k++) {
    s->tv[k].x = lx[k] + goal->cx;
    s->tv[k].y = ly[k] + goal->cy;
    s->tn[k].x = nx[k];
    s->tn[k].y = ny[k];
    if (s->tv[k].x < l) l = s->tv[k].x;
    if (s->tv[k].x > r) r = s->tv[k].x;
    if (s->tv[k].y < b) b = s->tv[k].y;
    if (s->tv[k].y > t) t = s->tv[k].y;
  }
  s->bb[0] = l; s->bb[1] = b; s->bb[2] = r; s->bb[3] = t;
}

/* entities.py:821-881 with com_overlap=True: every shape of the block collides with the sensor AND
 * the body's position lies inside the sensor's bounding box. */
int mgo_block_in_goal(mgo_env* e, int block, int goal) {
  const mg_block_t* blk = &e->scene.blocks[block];
  int slot = MG_MAX_SHAPES;
  load_goal_shape(e, goal, slot);
  const mgo_shape* gs = &e->shapes[slot];
  const mgo_body* body = &e->bodies[blk->body];
  if (!(gs->bb[0] <= body->p.x && gs->bb[2] >= body->p.x && gs->bb[1] <= body->p.y && gs->bb[3] >= body->p.y))
    return 0;
  const mg_cgroup_t* cg = &e->scene.cgroups[blk->cgroup];
  for (int s = cg->shape0; s < cg->shape0 + cg->nshape; s++) {
    int ia, ib, count;
    v2 n, p1[2], p2[2];
    unsigned hash[2];
    /* cpSpaceShapeQuery only visits shapes whose bounding boxes touch the query box */
    const double* bb = e->shapes[s].bb;
    if (!(bb[0] <= gs->bb[2] && gs->bb[0] <= bb[2] && bb[1] <= gs->bb[3] && gs->bb[1] <= bb[3])) return 0;
    mgo_collide(e, slot, s, &ia, &ib, &n, &count, p1, p2, hash);
    if (count == 0) return 0;
  }
  return 1;
}

static double score_move_to_corner(mgo_env* e) {
  /* move_to_corner.py:66-75 */
  const mgo_body* b = &e->bodies[e->scene.blocks[0].body];
  double dx = -1.0 - b->p.x, dy = 1.0 - b->p.y;
  double dist = sqrt(dx * dx + dy * dy);
  double succeed_dist = sqrt(2.0) / 2;
  double furthest_dist = sqrt(2.0);
  double drange = furthest_dist - succeed_dist;
  double sc = fmax(0.0, furthest_dist - dist) / drange;
  return sc < 1.0 ? sc : 1.0;
}

static double score_move_to_region(mgo_env* e) {
  /* move_to_region.py:85-94: point_query distance <= 0 <=> not strictly outside any plane */
  const mg_goal_t* g = &e->scene.goals[0];
  const mgo_body* r = &e->bodies[e->scene.robot_body];
  double hw = g->w / 2, hh = g->h / 2;
  int outside = (r->p.x - (hw + g->cx) > 0.0) || (r->p.y - (hh + g->cy) > 0.0) || (-(r->p.x - (-hw + g->cx)) > 0.0) ||
                (-(r->p.y - (-hh + g->cy)) > 0.0);
  return outside ? 0.0 : 1.0;
}

static double score_match_regions(mgo_env* e) {
  /* match_regions.py:193-213 */
  int n_targets = 0, n_t_in = 0, n_d_in = 0, n_in = 0;
  for (int i = 0; i < e->scene.n_blocks; i++) {
    int role = e->scene.blocks[i].role;
    if (role == 1) n_targets++;
    if (mgo_block_in_goal(e, i, 0)) {
      n_in++;
      if (role == 1) n_t_in++;
      if (role == 2) n_d_in++;
    }
  }
  double target_frac_done = (double)n_t_in / (double)n_targets;
  double contamination = n_in == 0 ? 0.0 : (double)n_d_in / (double)n_in;
  return target_frac_done * (1 - contamination);
}

static double score_find_dupe(mgo_env* e) {
  /* find_dupe.py:203-216 */
  int n_t_in = 0, n_d_in = 0, n_in = 0;
  for (int i = 0; i < e->scene.n_blocks; i++) {
    int role = e->scene.blocks[i].role;
    if (mgo_block_in_goal(e, i, 0)) {
      n_in++;
      if (role == 1) n_t_in++;
      if (role == 2) n_d_in++;
    }
  }
  double have_two = n_t_in >= 2 ? 1.0 : 0.0;
  double contamination = n_in == 0 ? 0.0 : (double)n_d_in / (double)n_in;
  return have_two * (1 - contamination);
}

static double score_fix_colour(mgo_env* e) {
  /* fix_colour.py:193-202: each sensor must hold exactly its expected block list */
  for (int g = 0; g < e->scene.n_goals; g++) {
    int expect = e->scene.goals[g].expect_block;
    for (int i = 0; i < e->scene.n_blocks; i++) {
      int in = mgo_block_in_goal(e, i, g);
      if (in != (i == expect)) return 0.0;
    }
  }
  return 1.0;
}

static int longest_line(const double (*pts)[2], int npts, double inlier_dist, double max_sep) {
  /* make_line.py:31-71 */
  int best = npts < 1 ? npts : 1;
  for (int i = 0; i < npts - 1; i++)
    for (int j = i + 1; j < npts; j++) {
      double ox = pts[j][0] - pts[i][0], oy = pts[j][1] - pts[i][1];
      double nrm = sqrt(ox * ox + oy * oy);
      double ux = ox / nrm, uy = oy / nrm;
      double proj[MG_MAX_BLOCKS];
      int n_in = 0;
      for (int k = 0; k < npts; k++) {
        double dx = pts[k][0] - pts[i][0], dy = pts[k][1] - pts[i][1];
        double pl = dx * ux + dy * uy;
        double rx = dx - pl * ux, ry = dy - pl * uy;
        double dist = sqrt(rx * rx + ry * ry);
        if (dist <= inlier_dist) proj[n_in++] = pl;
      }
      if (n_in <= best) continue;
      for (int a = 1; a < n_in; a++) { /* insertion sort */
        double v = proj[a];
        int b = a - 1;
        while (b >= 0 && proj[b] > v) { proj[b + 1] = proj[b]; b--; }
        proj[b + 1] = v;
      }
      int run = 0, longest = 0;
      for (int a = 0; a + 1 < n_in; a++) {
        if (fabs(proj[a + 1] - proj[a]) <= max_sep) { run++; if (run > longest) longest = run; }
        else run = 0;
      }
      if (longest + 1 > best) best = longest + 1;
    }
  return best;
}

static double score_make_line(mgo_env* e) {
  /* make_line.py:142-152, thresholds :10-11, :90-91 */
  double pts[MG_MAX_BLOCKS][2];
  int n = e->scene.n_blocks;
  for (int i = 0; i < n; i++) {
    pts[i][0] = e->bodies[e->scene.blocks[i].body].p.x;
    pts[i][1] = e->bodies[e->scene.blocks[i].body].p.y;
  }
  double shape_rad = 0.2 * 0.6;
  int line_len = longest_line(pts, n, shape_rad * 1.5, shape_rad * 3.5);
  int min_line_len = n - 2 > 2 ? n - 2 : 2;
  int num = line_len - min_line_len;
  if (num < 0) num = 0;
  return (double)num / (double)(n - min_line_len);
}

static double score_cluster(mgo_env* e) {
  /* cluster.py:166-216 */
  int nvals = e->scene.n_labels, n = e->scene.n_blocks;
  double cent[MG_MAX_BLOCKS][2];
  for (int c = 0; c < nvals; c++) {
    double sx = 0.0, sy = 0.0;
    int cnt = 0;
    for (int i = 0; i < n; i++)
      if (e->scene.blocks[i].label == c) {
        sx += e->bodies[e->scene.blocks[i].body].p.x;
        sy += e->bodies[e->scene.blocks[i].body].p.y;
        cnt++;
      }
    cent[c][0] = cnt ? sx / cnt : 0.0;
    cent[c][1] = cnt ? sy / cnt : 0.0;
  }
  int n_correct = 0;
  for (int i = 0; i < n; i++) {
    int lab = e->scene.blocks[i].label;
    double px = e->bodies[e->scene.blocks[i].body].p.x, py = e->bodies[e->scene.blocks[i].body].p.y;
    double true_sse = 0.0, nearest_bad = INFINITY;
    for (int c = 0; c < nvals; c++) {
      double dx = px - cent[c][0], dy = py - cent[c][1];
      double sse = dx * dx + dy * dy;
      if (c == lab) true_sse = sse;
      else if (sse < nearest_bad) nearest_bad = sse;
    }
    double margin = 2.0 * true_sse; /* sic: margin uses the squared distance (cluster.py:203-206) */
    n_correct += sqrt(true_sse) < sqrt(nearest_bad) - margin;
  }
  double frac = (double)n_correct / (double)(n > 1 ? n : 1);
  double thresh = 0.75;
  return fmax(frac - thresh, 0.0) / (1 - thresh);
}

double mgo_score(mgo_env* e) {
  switch (e->scene.task) {
    case MG_TASK_MOVE_TO_CORNER: return score_move_to_corner(e);
    case MG_TASK_MOVE_TO_REGION: return score_move_to_region(e);
    case MG_TASK_MATCH_REGIONS: return score_match_regions(e);
    case MG_TASK_MAKE_LINE: return score_make_line(e);
    case MG_TASK_FIND_DUPE: return score_find_dupe(e);
    case MG_TASK_FIX_COLOUR: return score_fix_colour(e);
    case MG_TASK_CLUSTER_COLOUR:
    case MG_TASK_CLUSTER_SHAPE: return score_cluster(e);
  }
  return 0.0;
}

double mgo_debug_reward(mgo_env* e) {
  /* move_to_corner.py:84-98 (target is (0, 1) there, sic) */
  const mgo_body* sb = &e->bodies[e->scene.blocks[0].body];
  const mgo_body* rb = &e->bodies[e->scene.robot_body];
  double dx = sb->p.x - 0.0, dy = sb->p.y - 1.0;
  double shape_to_corner = sqrt(dx * dx + dy * dy);
  double ex = rb->p.x - sb->p.x, ey = rb->p.y - sb->p.y;
  double robot_to_shape = sqrt(ex * ex + ey * ey);
  double shaping = -shape_to_corner / 5 - fmax(robot_to_shape, 0.2) / 20;
  return shaping + score_move_to_corner(e);
}

int64_t mgo_sizeof_scene(void) { return (int64_t)sizeof(mg_scene_t); }
int64_t mgo_sizeof_state(void) { return (int64_t)sizeof(mg_state_t); }
