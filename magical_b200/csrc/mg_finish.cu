/*
 * mg_finish.cu — K2/K4: episode bookkeeping, end-of-trajectory score, reward
 * and (auto-)reset, one thread per environment.
 *
 * Replaces the tail of `BaseEnv.step` (magical/base_env.py:265-288: step
 * counter, `done`, `score_on_end_of_traj()` only on the final step, reward 0
 * except the DebugReward MoveToCorner variant, move_to_corner.py:77-98) and
 * the state half of `BaseEnv.reset` (base_env.py:177-223).  The seven score
 * functions restate benchmarks/*.py `score_on_end_of_traj` and
 * `GoalRegion.get_overlapping_ents(com_overlap=True)` (entities.py:821-881).
 *
 * This work is tiny next to K1 (it touches one env per thread, and only envs
 * whose episode just ended do more than increment a counter), so it favours
 * simplicity over lane cooperation.
 */
#include "mg_device.cuh"
#include "mg_narrowphase.h"
#include "mg_sincos.h"
#include "mg_reset.cuh"

__device__ static ShapeView state_view(const EnvState& st, const DeviceScene* ds, int si) {
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
    v.rc = st.R[b].x; v.rs = st.R[b].y; v.px = st.P[b].x; v.py = st.P[b].y;
  } else {
    v.rc = 1.0; v.rs = 0.0; v.px = 0.0; v.py = 0.0;
  }
  return v;
}

/* entities.py:821-881 with com_overlap=True */
__device__ static bool block_in_goal(const EnvState& st, const DeviceScene* ds, int block, int goal) {
  const mg_scene_t& sc = ds->s;
  const mg_goal_t& g = sc.goals[goal];
  double hw = g.w / 2, hh = g.h / 2;
  /* Poly.create_box vertex order (r,b) (r,t) (l,t) (l,b) */
  double lv[8] = {hw, -hh, hw, hh, -hw, hh, -hw, -hh};
  double ln[8] = {0, -1, 1, 0, 0, 1, -1, 0};
  ShapeView gv;
  gv.kind = MG_SHAPE_POLY; gv.nvert = 4; gv.lv = lv; gv.ln = ln;
  gv.rc = 1.0; gv.rs = 0.0; gv.px = g.cx; gv.py = g.cy; gv.radius = 0.0; gv.index = MG_MAX_SHAPES;
  double gbb[4];
  sv_bb(gv, gbb);
  const mg_block_t& blk = sc.blocks[block];
  double px = st.P[blk.body].x, py = st.P[blk.body].y;
  if (!(gbb[0] <= px && gbb[2] >= px && gbb[1] <= py && gbb[3] >= py)) return false;
  int s0 = sc.cgroups[blk.cgroup].shape0, n = sc.cgroups[blk.cgroup].nshape;
  for (int s = s0; s < s0 + n; s++) {
    ShapeView bv = state_view(st, ds, s);
    double bbb[4];
    sv_bb(bv, bbb);
    if (!bb_intersects(bbb, gbb)) return false;
    Manifold m;
    /* cpCollide order: lower shape type first; the query (goal) shape first among equals */
    if (bv.kind < gv.kind) mg_collide(bv, gv, bbb, gbb, m);
    else mg_collide(gv, bv, gbb, bbb, m);
    if (m.count == 0) return false;
  }
  return true;
}

__device__ static double score_move_to_corner(const EnvState& st, const mg_scene_t& sc) {
  int b = sc.blocks[0].body;
  double dx = -1.0 - st.P[b].x, dy = 1.0 - st.P[b].y;
  double dist = sqrt(dx * dx + dy * dy);
  double succeed_dist = sqrt(2.0) / 2;
  double furthest_dist = sqrt(2.0);
  double drange = furthest_dist - succeed_dist;
  double s = fmax(0.0, furthest_dist - dist) / drange;
  return s < 1.0 ? s : 1.0;
}

__device__ static int longest_line(const double (*pts)[2], int npts, double inlier_dist, double max_sep) {
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
      for (int a = 1; a < n_in; a++) {
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

__device__ static double compute_score(const EnvState& st, const DeviceScene* ds) {
  const mg_scene_t& sc = ds->s;
  switch (sc.task) {
    case MG_TASK_MOVE_TO_CORNER:
      return score_move_to_corner(st, sc); /* move_to_corner.py:66-75 */
    case MG_TASK_MOVE_TO_REGION: {         /* move_to_region.py:85-94 */
      const mg_goal_t& g = sc.goals[0];
      double x = st.P[sc.robot_body].x, y = st.P[sc.robot_body].y;
      double hw = g.w / 2, hh = g.h / 2;
      bool outside = (x - (hw + g.cx) > 0.0) || (y - (hh + g.cy) > 0.0) || (-(x - (-hw + g.cx)) > 0.0) ||
                     (-(y - (-hh + g.cy)) > 0.0);
      return outside ? 0.0 : 1.0;
    }
    case MG_TASK_MATCH_REGIONS: { /* match_regions.py:193-213 */
      int n_targets = 0, n_t_in = 0, n_d_in = 0, n_in = 0;
      for (int i = 0; i < sc.n_blocks; i++) {
        int role = sc.blocks[i].role;
        if (role == 1) n_targets++;
        if (block_in_goal(st, ds, i, 0)) {
          n_in++;
          if (role == 1) n_t_in++;
          if (role == 2) n_d_in++;
        }
      }
      double frac = (double)n_t_in / (double)n_targets;
      double contamination = n_in == 0 ? 0.0 : (double)n_d_in / (double)n_in;
      return frac * (1 - contamination);
    }
    case MG_TASK_MAKE_LINE: { /* make_line.py:142-152 */
      double pts[MG_MAX_BLOCKS][2];
      int n = sc.n_blocks;
      for (int i = 0; i < n; i++) { pts[i][0] = st.P[sc.blocks[i].body].x; pts[i][1] = st.P[sc.blocks[i].body].y; }
      double shape_rad = 0.2 * 0.6;
      int line_len = longest_line(pts, n, shape_rad * 1.5, shape_rad * 3.5);
      int min_line_len = n - 2 > 2 ? n - 2 : 2;
      int num = line_len - min_line_len;
      if (num < 0) num = 0;
      return (double)num / (double)(n - min_line_len);
    }
    case MG_TASK_FIND_DUPE: { /* find_dupe.py:203-216 */
      int n_t_in = 0, n_d_in = 0, n_in = 0;
      for (int i = 0; i < sc.n_blocks; i++) {
        int role = sc.blocks[i].role;
        if (block_in_goal(st, ds, i, 0)) {
          n_in++;
          if (role == 1) n_t_in++;
          if (role == 2) n_d_in++;
        }
      }
      double have_two = n_t_in >= 2 ? 1.0 : 0.0;
      double contamination = n_in == 0 ? 0.0 : (double)n_d_in / (double)n_in;
      return have_two * (1 - contamination);
    }
    case MG_TASK_FIX_COLOUR: { /* fix_colour.py:193-202 */
      for (int g = 0; g < sc.n_goals; g++) {
        int expect = sc.goals[g].expect_block;
        for (int i = 0; i < sc.n_blocks; i++) {
          bool in = block_in_goal(st, ds, i, g);
          if (in != (i == expect)) return 0.0;
        }
      }
      return 1.0;
    }
    case MG_TASK_CLUSTER_COLOUR:
    case MG_TASK_CLUSTER_SHAPE: { /* cluster.py:166-216 */
      int nvals = sc.n_labels, n = sc.n_blocks;
      double cent[MG_MAX_BLOCKS][2];
      for (int c = 0; c < nvals; c++) {
        double sx = 0.0, sy = 0.0;
        int cnt = 0;
        for (int i = 0; i < n; i++)
          if (sc.blocks[i].label == c) { sx += st.P[sc.blocks[i].body].x; sy += st.P[sc.blocks[i].body].y; cnt++; }
        cent[c][0] = cnt ? sx / cnt : 0.0;
        cent[c][1] = cnt ? sy / cnt : 0.0;
      }
      int n_correct = 0;
      for (int i = 0; i < n; i++) {
        int lab = sc.blocks[i].label;
        double px = st.P[sc.blocks[i].body].x, py = st.P[sc.blocks[i].body].y;
        double true_sse = 0.0, nearest_bad = MG_INF;
        for (int c = 0; c < nvals; c++) {
          double dx = px - cent[c][0], dy = py - cent[c][1];
          double sse = dx * dx + dy * dy;
          if (c == lab) true_sse = sse;
          else if (sse < nearest_bad) nearest_bad = sse;
        }
        double margin = 2.0 * true_sse;
        n_correct += (sqrt(true_sse) < sqrt(nearest_bad) - margin) ? 1 : 0;
      }
      double frac = (double)n_correct / (double)(n > 1 ? n : 1);
      double thresh = 0.75;
      return fmax(frac - thresh, 0.0) / (1 - thresh);
    }
  }
  return 0.0;
}

/* mode 0: full step tail; mode 1: score of the current state only (mg_score) */
__device__ __forceinline__ uint32_t mix32(uint32_t x) { /* lowbias32 integer hash */
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void k_finish(EnvState* __restrict__ states, const DeviceScene* __restrict__ scenes, int env0, int count,
                         int auto_reset, int mode, int draw_first, int draw_count, uint32_t reset_seed, int slot_base,
                         float* __restrict__ reward,
                         uint8_t* __restrict__ done, float* __restrict__ score,
                         unsigned long long* __restrict__ overflow_count) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= count) return;
  env += env0; /* this launch covers environments [env0, env0 + count) */
  EnvState& st = states[env];
  /* device-side layout sampling (slot_base >= 0): the environment's sampled goal sensors live in its slot */
  const bool sample_mode = slot_base >= 0;
  const DeviceScene* ds = scenes + (sample_mode ? slot_base + env : st.scene);
  const mg_scene_t& sc = ds->s;
  if (mode == 1) {
    if (score) score[env] = (float)compute_score(st, ds);
    return;
  }
  /* capacity overflows of the physics (flags 2 / 4: that sub-step was solved without contacts / the contact
   * cache was truncated) are counted once per environment and episode, so a caller can check a whole run */
  if (st.overflow != 0 && (st.overflow & 0x100) == 0) {
    st.overflow |= 0x100;
    if (overflow_count) atomicAdd(overflow_count, 1ull);
  }
  int steps = st.episode_steps + 1;
  st.episode_steps = steps;
  bool d = sc.max_steps > 0 && steps >= sc.max_steps;
  double s = 0.0;
  if (d) s = compute_score(st, ds);
  double rew = 0.0;
  if (sc.debug_reward) {
    /* move_to_corner.py:84-98 (its shaping target is (0, 1), sic) */
    int sb = sc.blocks[0].body, rb = sc.robot_body;
    double dx = st.P[sb].x - 0.0, dy = st.P[sb].y - 1.0;
    double shape_to_corner = sqrt(dx * dx + dy * dy);
    double ex = st.P[rb].x - st.P[sb].x, ey = st.P[rb].y - st.P[sb].y;
    double robot_to_shape = sqrt(ex * ex + ey * ey);
    double shaping = -shape_to_corner / 5 - fmax(robot_to_shape, 0.2) / 20;
    rew = shaping + score_move_to_corner(st, sc);
  }
  if (reward) reward[env] = (float)rew;
  if (done) done[env] = d ? 1 : 0;
  if (score) score[env] = (float)s;
  if (d && auto_reset) {
    int scene = st.scene;
    if (sample_mode && !(draw_count > 0 && (draw_count > 1 || draw_first != scene))) ++st.resets;
    if (draw_count > 0 && (draw_count > 1 || draw_first != scene)) {
      /* randomised variants: the next episode plays a freshly drawn scene of the pool's current draw range */
      const int resets = ++st.resets;
      scene = draw_first +
              (int)(mix32(mix32(reset_seed ^ mix32((uint32_t)env)) + (uint32_t)resets) % (uint32_t)draw_count);
    }
    mg_reset_state(st, scenes + scene, scene);
    /* device-side layout sampling: `scene` is a template; k_sample_layouts (next on the stream) draws sizes
     * and poses into the environment's own slot and resets it again from there */
    if (sample_mode) st.fresh = MG_FRESH_SAMPLE;
  }
}

/* explicit reset of selected envs (env_ids == nullptr: all), optional new scene index per env */
__global__ void k_reset(EnvState* __restrict__ states, const DeviceScene* __restrict__ scenes, int n,
                        const int32_t* __restrict__ env_ids, const int32_t* __restrict__ scene_ids, int first_time,
                        int sample_mode) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int env = env_ids ? env_ids[i] : i;
  EnvState& st = states[env];
  int scene = scene_ids ? scene_ids[i] : (first_time ? 0 : st.scene);
  mg_reset_state(st, scenes + scene, scene);
  for (int k = 0; k < MG_NCACHE; k++) {
    CEntry e;
    e.a = e.b = e.used = e.pad_ = 0; e.hash = 0; e.stamp = -100; e.pad2_ = 0; e.jn = e.jt = 0.0;
    st.cache[k] = e;
  }
  if (sample_mode && !first_time) { st.resets++; st.fresh = MG_FRESH_SAMPLE; }
}

cudaError_t mg_launch_finish(EnvState* states, const DeviceScene* scenes, int env0, int count, int auto_reset, int mode,
                             int draw_first, int draw_count, uint32_t reset_seed, int slot_base, float* reward,
                             uint8_t* done, float* score, unsigned long long* overflow_count, cudaStream_t stream) {
  int threads = 128;
  k_finish<<<(count + threads - 1) / threads, threads, 0, stream>>>(states, scenes, env0, count, auto_reset, mode,
                                                                    draw_first, draw_count, reset_seed, slot_base,
                                                                    reward, done, score, overflow_count);
  return cudaGetLastError();
}

cudaError_t mg_launch_reset(EnvState* states, const DeviceScene* scenes, int n, const int32_t* env_ids,
                            const int32_t* scene_ids, int first_time, int sample_mode, cudaStream_t stream) {
  int threads = 128;
  k_reset<<<(n + threads - 1) / threads, threads, 0, stream>>>(states, scenes, n, env_ids, scene_ids, first_time,
                                                               sample_mode);
  return cudaGetLastError();
}
