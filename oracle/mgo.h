/*
 * mgo.h — CPU ORACLE internal types (test infrastructure; see mgo_physics.c).
 * Consumes the same compiled-scene tables as the CUDA library
 * (include/magical_b200.h) so tests drive both through identical inputs.
 */
#ifndef MGO_H
#define MGO_H

#include <stdint.h>

#include "../include/magical_b200.h"

#define MGO_DBL_MIN 2.2250738585072014e-308
#define MGO_MAX_ARBITERS 256
#define MGO_MAX_POLY 16

typedef struct { double x, y; } v2;

typedef struct {
  double m_inv, i_inv;
  v2 p, v;
  double a, w;
  v2 rot;
  v2 v_bias;
  double w_bias;
  int kind;
} mgo_body;

typedef struct {
  int kind, body, nvert, group;
  double radius, friction;
  v2 lv[MGO_MAX_POLY], ln[MGO_MAX_POLY]; /* local verts / plane normals */
  v2 tv[MGO_MAX_POLY], tn[MGO_MAX_POLY]; /* world-space cache */
  double bb[4];                          /* l, b, r, t */
} mgo_shape;

typedef struct {
  v2 r1, r2;
  double nMass, tMass, bounce, jnAcc, jtAcc, jBias, bias;
  unsigned hash;
} mgo_contact;

enum { MGO_ARB_FIRST = 0, MGO_ARB_NORMAL = 1, MGO_ARB_CACHED = 2 };

typedef struct {
  int a, b; /* shape indices, type-ordered */
  int body_a, body_b;
  v2 n;
  double u;
  int count;
  mgo_contact contacts[2];
  int stamp, state;
} mgo_arbiter;

typedef struct {
  int kind, a, b;
  v2 anchor_a, anchor_b;
  double p0, p1, p2, max_force, max_bias, error_bias;
  /* solver state */
  v2 r1, r2, n, bias_v, jAcc;
  double k[4];
  double bias, iSum, nMass, w_coef, target_wrn, rate;
} mgo_joint;

typedef struct mgo_env {
  mg_scene_t scene;
  int n_bodies, n_shapes, n_joints;
  mgo_body static_body;
  mgo_body bodies[MG_MAX_BODIES];
  mgo_shape shapes[MG_MAX_SHAPES + 1]; /* +1: scratch slot for goal sensor boxes at score time */
  mgo_joint joints[MG_MAX_JOINTS];
  mgo_arbiter cached[MGO_MAX_ARBITERS];
  int n_cached;
  int active[MGO_MAX_ARBITERS];
  int n_active;
  int stamp;
  double curr_dt;
  double collision_bias;
  int episode_steps;
  int overflow;
  int det_sincos;
  long stat_substeps, stat_bb_pass, stat_circle_pass, stat_hits, stat_contacts; /* instrumentation */
  int32_t* pair_perm; /* optional permutation of the canonical pair order (sensitivity study) */
  /* Robot.set_action state */
  double rel_turn_angle, target_speed, target_finger_angle;
} mgo_env;

#ifdef __cplusplus
extern "C" {
#endif

mgo_env* mgo_create(const mg_scene_t* scene);
void mgo_destroy(mgo_env* e);
void mgo_reset(mgo_env* e);
void mgo_set_det_sincos(mgo_env* e, int on);
void mgo_set_pair_permutation(mgo_env* e, const int32_t* perm);
void mgo_set_action(mgo_env* e, int action);
void mgo_robot_update(mgo_env* e);
void mgo_space_step(mgo_env* e, double dt);
void mgo_phys_steps_on_frame(mgo_env* e);
void mgo_step(mgo_env* e, int action, float* reward, uint8_t* done, float* score);
void mgo_get_state(const mgo_env* e, mg_state_t* out);
void mgo_set_state(mgo_env* e, const mg_state_t* in);
void mgo_set_pose(mgo_env* e, int body, double x, double y, double angle);
void mgo_collide(const mgo_env* e, int ia, int ib, int* out_a, int* out_b, v2* n, int* count, v2 p1[2], v2 p2[2],
                 unsigned hash[2]);
void mgo_get_stats(const mgo_env* e, long out[5]);
double mgo_score(mgo_env* e);
double mgo_debug_reward(mgo_env* e);
int mgo_block_in_goal(mgo_env* e, int block, int goal);
/* render one view at `res` x `res` (res = 384 in the reference): view 0 = allo, 1 = ego. out: u8[res*res*3] */
void mgo_render_view(const mgo_env* e, int view, int res, uint8_t* out);
/* 4x4 INTER_AREA box mean of a (4n x 4n x 3) image -> (n x n x 3) */
void mgo_downsample4(const uint8_t* src, int n_out, uint8_t* dst);
int64_t mgo_sizeof_scene(void);
int64_t mgo_sizeof_state(void);

#ifdef __cplusplus
}
#endif
#endif
