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

#define MG_NARB 24   /* cached arbiters (colliding shape pairs incl. those seen in the last 3 sub-steps) */
#define MG_NCON 24   /* solver contacts per sub-step */
#define MG_NCAND 64  /* narrowphase candidates per sub-step */
#define MG_PERSISTENCE 3

/* One cached arbiter = one colliding shape pair with its (<=2) contact accumulators. */
struct __align__(16) ArbEntry {
  uint8_t a, b;      /* shape indices, type-ordered as Chipmunk's cpCollide orders them */
  uint8_t count;     /* contacts */
  uint8_t pad_;
  int32_t stamp;     /* sub-step stamp of the last collision */
  uint32_t hash[2];  /* contact feature ids */
  double jn[2], jt[2];
};

/* Per-environment simulator state in HBM: one contiguous, 16-byte-aligned record that a
 * warp streams in and out with 128-bit loads (7 per lane). */
struct __align__(16) EnvState {
  int32_t scene, episode_steps, stamp, n_arb, overflow, fresh, last_contacts, pad_;
  double4 V[MG_MAX_BODIES];  /* vx, vy, w, - */
  double4 Bv[MG_MAX_BODIES]; /* bias velocities vbx, vby, wb, - (consumed by the next position update) */
  double4 P[MG_MAX_BODIES];  /* px, py, angle, - */
  double2 R[MG_MAX_BODIES];  /* cos, sin of angle */
  double2 jacc[MG_MAX_JOINTS];
  ArbEntry arb[MG_NARB];
};

struct DeviceScene {
  mg_scene_t s;
  mg_scene_aux_t aux;
};

#endif
