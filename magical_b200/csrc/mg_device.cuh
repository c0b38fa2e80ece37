/*
 * mg_device.cuh — device-side data layout shared by the physics, finish
 * (score/reset) and raster kernels.
 */
#ifndef MG_DEVICE_CUH
#define MG_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/magical_b200.h"
#include "mg_scene_aux.h"
#include "mg_raster_aux.h"

#define MG_NCACHE 48 /* cached contacts: this sub-step's (<= 32) + those of pairs that collided in the last 3 sub-steps;
                        at most 64 (the thread-per-environment kernel tracks matched entries in a 64-bit mask) */
#define MG_NCON 24   /* solver contacts per sub-step (16 when two environments share a warp) */
#define MG_NCAND 64  /* narrowphase candidates per sub-step */
#define MG_PERSISTENCE 3
#define MG_MAX_PINS 4

/* One cached contact: Chipmunk keeps jnAcc/jtAcc per contact hash inside the shape pair's arbiter for
 * `collisionPersistence` steps; a pair's entries all carry the stamp of its last collision. */
struct __align__(16) CEntry {
  uint8_t a, b;      /* shape indices, type-ordered as Chipmunk's cpCollide orders them */
  uint8_t used;      /* scratch: matched by a collision of the current sub-step */
  uint8_t pad_;
  uint32_t hash;     /* contact feature id */
  int32_t stamp;     /* sub-step stamp of the pair's last collision */
  int32_t pad2_;
  double jn, jt;
};

/* Per-environment simulator state in HBM: one contiguous, 16-byte-aligned record that a warp (or
 * half-warp) streams in and out with 128-bit loads. */
struct __align__(16) EnvState {
  int32_t scene, episode_steps, stamp, n_cache, overflow, fresh, last_contacts, resets;
  double4 V[MG_MAX_BODIES];  /* vx, vy, w, - */
  double4 Bv[MG_MAX_BODIES]; /* bias velocities vbx, vby, wb, - (consumed by the next position update) */
  double4 P[MG_MAX_BODIES];  /* px, py, angle, - */
  double2 R[MG_MAX_BODIES];  /* cos, sin of angle */
  double2 jacc[MG_MAX_JOINTS];
  CEntry cache[MG_NCACHE];
};

struct DeviceScene {
  mg_scene_t s;
  mg_scene_aux_t aux;
  mg_raster_aux_t ra;
};

#endif
