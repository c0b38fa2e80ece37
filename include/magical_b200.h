/*
 * magical_b200.h — C ABI of the B200-native MAGICAL hot path.
 *
 * The reference (qxcv/magical) has no FFI of its own: its hot path is
 * `BaseEnv.step()` (magical/base_env.py:255-292), which hands the physics to
 * pymunk/Chipmunk2D (`pm.Space.step`, base_env.py:236-243), the two-camera
 * render to pyglet/OpenGL (base_env.py:309-338, gym_render.py:208-249) and the
 * 96x96 downsample + frame stack to OpenCV/gym wrappers
 * (benchmarks/__init__.py:46-274).  This header is the boundary that replaces
 * those three native call sites with ONE batched library: plain pointers and
 * sizes, no torch types, loadable with ctypes (see INTEGRATION.md for the
 * reference-side binding).
 *
 * Data model
 *   - A *compiled scene* (`mg_scene_t`) is the flat-table form of what
 *     `Entity.setup()` builds imperatively in the reference
 *     (entities.py:238-437, 502-537, 614-757, 790-819): bodies, collision
 *     shapes, joints, draw primitives, goal sensors and score metadata.  The
 *     Python host code (magical_b200/entities.py) produces it.
 *   - A handle owns `n_scenes` compiled scenes and `batch` environments, each
 *     bound to one scene index; all simulator state lives in device memory.
 *   - The observation buffer is caller-owned device memory (e.g. a torch uint8
 *     tensor's data_ptr) that stays bound to the handle: the frame stack is
 *     shifted in place inside it every step.
 *
 * Error model: every entry point returns 0 on success or a negative MG_E_*
 * code; `mg_last_error()` returns a thread-local message.  Nothing throws
 * across the ABI.  A handle must be used from one host thread at a time.
 * Kernels are enqueued on the stream given at `mg_create`; `mg_step` does not
 * synchronise.
 */
#ifndef MAGICAL_B200_H
#define MAGICAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 2

/* capacity limits of one compiled scene */
#define MG_MAX_BODIES 16   /* non-static bodies (<=15 dynamic + 1 kinematic) */
#define MG_MAX_SHAPES 72   /* solid collision shapes incl. the 4 arena walls */
#define MG_MAX_CVERTS 384  /* collision vertex pool (double2) */
#define MG_MAX_JOINTS 32
#define MG_MAX_CGROUPS 20  /* collision groups: one per shaped body + 4 walls */
#define MG_MAX_BPAIRS 160  /* candidate collision-group pairs */
#define MG_MAX_GOALS 3
#define MG_MAX_BLOCKS 10
#define MG_MAX_PRIMS 160   /* draw primitives */
#define MG_MAX_DVERTS 704  /* draw vertex pool (float2) */

/* error codes */
#define MG_OK 0
#define MG_E_INVALID -1   /* bad argument */
#define MG_E_CUDA -2      /* CUDA runtime failure (message has the detail) */
#define MG_E_NOMEM -3
#define MG_E_STATE -4     /* call not valid in the handle's current state */

/* shape kinds (Chipmunk's type order: circle < segment < poly) */
enum { MG_SHAPE_CIRCLE = 0, MG_SHAPE_SEGMENT = 1, MG_SHAPE_POLY = 2 };
/* body kinds */
enum { MG_BODY_DYNAMIC = 0, MG_BODY_KINEMATIC = 1 };
/* joint kinds, the six pymunk constraints the reference uses
 * (entities.py:255-277, 334-354, 703-711) */
enum {
  MG_JOINT_PIVOT = 0,
  MG_JOINT_GEAR = 1,
  MG_JOINT_ROTARY_SPRING = 2,
  MG_JOINT_PIN = 3,
  MG_JOINT_ROTARY_LIMIT = 4,
  MG_JOINT_MOTOR = 5
};
/* draw primitive kinds / transforms */
enum { MG_PRIM_POLY = 0, MG_PRIM_NGON = 1, MG_PRIM_LINELOOP = 2 };
enum { MG_XFORM_WORLD = 0, MG_XFORM_BODY = 1, MG_XFORM_PUPIL = 2 };
/* tasks (score function selector), benchmarks/<task>.py */
enum {
  MG_TASK_MOVE_TO_CORNER = 0,
  MG_TASK_MOVE_TO_REGION = 1,
  MG_TASK_MATCH_REGIONS = 2,
  MG_TASK_MAKE_LINE = 3,
  MG_TASK_FIND_DUPE = 4,
  MG_TASK_FIX_COLOUR = 5,
  MG_TASK_CLUSTER_COLOUR = 6,
  MG_TASK_CLUSTER_SHAPE = 7
};
/* observation layouts = the reference's preprocessors
 * (benchmarks/__init__.py:242-274) plus the raw two-view dict */
enum {
  MG_OBS_LORES4E = 0,    /* u8 [B,96,96,12]  4 ego frames, oldest first     */
  MG_OBS_LORES4A = 1,    /* u8 [B,96,96,12]  4 allo frames                  */
  MG_OBS_LORES3EA = 2,   /* u8 [B,96,96,12]  1 allo + 3 ego                 */
  MG_OBS_LORESSTACK = 3, /* u8 [2,B,96,96,12] plane 0 = allo, plane 1 = ego */
  MG_OBS_LORESCHW4E = 4, /* u8 [B,12,96,96]                                 */
  MG_OBS_RAW = 5         /* u8 [2,B,res,res,3] plane 0 = allo, 1 = ego      */
};

typedef struct {
  double m_inv, i_inv; /* 0 for kinematic bodies */
  double p0[2];        /* pose at reset */
  double a0;
  int32_t kind;
  int32_t pad_;
} mg_body_t;

typedef struct {
  int32_t kind;   /* MG_SHAPE_* */
  int32_t body;   /* -1 = the static body */
  int32_t vert0;  /* first vertex in cverts: circle 1 (centre), segment 2, poly n */
  int32_t nvert;
  double radius;  /* circle radius / segment radius / poly bevel radius */
  double friction;
  int32_t group;  /* Chipmunk ShapeFilter group, 0 = none */
  int32_t pad_;
} mg_shape_t;

typedef struct {
  int32_t kind; /* MG_JOINT_* */
  int32_t a, b; /* body indices, -1 = static body */
  int32_t pad_;
  double anchor_a[2], anchor_b[2];
  /* gear: p0=phase p1=ratio | spring: p0=rest p1=stiffness p2=damping |
   * limit: p0=min p1=max | pin: p0=rest distance */
  double p0, p1, p2;
  double max_force, max_bias, error_bias;
} mg_joint_t;

typedef struct {
  uint8_t shape0, nshape; /* contiguous shape range */
  int8_t body;            /* -1 static */
  uint8_t robot_group;    /* 1 if the shapes carry the robot's filter group */
} mg_cgroup_t;

typedef struct {
  uint8_t kind;   /* MG_PRIM_* */
  uint8_t xform;  /* MG_XFORM_* */
  uint8_t body;   /* body whose pose places the primitive (XFORM_BODY/PUPIL) */
  uint8_t body2;  /* XFORM_PUPIL: the eye body whose angle spins the pupil */
  uint8_t rgb[3];
  uint8_t nvert;  /* POLY/LINELOOP: vertex count; NGON: number of sides */
  uint16_t vert0; /* POLY/LINELOOP: first vertex in dverts */
  uint16_t stipple; /* LINELOOP: 16-bit GL stipple pattern, 0xFFFF = solid */
  float cx, cy;   /* NGON: centre in the body frame (eye offset for pupils) */
  float radius;   /* NGON: circumradius; LINELOOP: width in 384-res pixels */
  float ex, ey;   /* PUPIL: offset applied before the pupil's own rotation */
} mg_prim_t;

typedef struct {
  double cx, cy, w, h; /* sensor box centre and size (entities.py:794-797) */
  int32_t colour;
  int32_t expect_block; /* FixColour: block that must be the only one inside, -1 = must be empty */
} mg_goal_t;

typedef struct {
  int32_t body;       /* body index */
  int32_t cgroup;     /* collision group holding all its shapes */
  int32_t shape_type; /* 0 square 1 pentagon 2 star 3 circle 4 triangle 5 hexagon 6 octagon */
  int32_t colour;     /* 0 red 1 green 2 blue 3 yellow */
  int32_t role;       /* task-specific: 1 target, 2 distractor, 0 none */
  int32_t label;      /* Cluster*: index of the block's cluster value */
} mg_block_t;

typedef struct {
  /* ---- header ---- */
  int32_t task;           /* MG_TASK_* */
  int32_t max_steps;      /* episode length (benchmarks/__init__.py:407-813) */
  int32_t debug_reward;   /* MoveToCorner DebugReward variant */
  int32_t n_bodies, n_shapes, n_cverts, n_joints, n_cgroups, n_bpairs;
  int32_t n_goals, n_blocks, n_prims, n_dverts;
  int32_t n_labels;       /* Cluster*: number of distinct cluster values */
  /* robot wiring for Robot.update (entities.py:459-479) */
  int32_t robot_body, control_body;
  int32_t finger_body[2], motor_joint[2], eye_body[2];
  int32_t pad_[2];
  double robot_radius;
  /* ---- tables ---- */
  mg_body_t bodies[MG_MAX_BODIES];
  mg_shape_t shapes[MG_MAX_SHAPES];
  double cverts[MG_MAX_CVERTS][2];
  mg_joint_t joints[MG_MAX_JOINTS];
  mg_cgroup_t cgroups[MG_MAX_CGROUPS];
  uint8_t bpairs[MG_MAX_BPAIRS][2]; /* canonical arbiter order */
  mg_goal_t goals[MG_MAX_GOALS];
  mg_block_t blocks[MG_MAX_BLOCKS];
  mg_prim_t prims[MG_MAX_PRIMS];   /* painter's order */
  float dverts[MG_MAX_DVERTS][2];
} mg_scene_t;

typedef struct {
  int32_t device;       /* CUDA device ordinal */
  int32_t batch;        /* environments owned by this handle */
  int32_t n_scenes;
  int32_t obs_mode;     /* MG_OBS_* */
  int32_t res;          /* MG_OBS_RAW: render resolution (384); ignored otherwise */
  int32_t auto_reset;   /* 1: envs that finish an episode are reset inside mg_step
                              and the returned observation is the new episode's first */
  int32_t reserved0_;   /* must be 0 (was `fast_math`: the library has ONE arithmetic, fp64 with contraction off,
                           the reference's -- a faster, narrower solver would not be the reference's path) */
  int32_t reset_seed;   /* n_scenes > 1 with auto_reset: an env that finishes an episode draws its next
                           scene on the device as hash(reset_seed, env, resets so far) % n_scenes (the
                           reference re-randomises the layout on every reset, base_env.py:177-234) */
  int32_t keep_scene;   /* 1: auto-reset restarts an env on the scene it is bound to (e.g. a mixed-task batch)
                           instead of drawing a new pool entry */
  int32_t device_sampling; /* 1: the n_scenes scenes are TEMPLATES (structure: shape types, colours, counts, dynamics);
                           every reset draws a template and rejection-samples fresh goal sizes and poses for it ON
                           THE DEVICE into a scene slot owned by the environment (mg_set_placement) */
  int32_t reserved_[6];
} mg_config_t;

/* ---- placement program of one template: what the task's on_reset() asks the reference's rejection sampler
 * to randomise (pm_randomise_all_poses / randomise_hw, magical/geom.py:116-359), recorded by the host while it
 * builds the template.  Entities are placed in list order, each avoiding the arena walls, every entity that is
 * not on the list, and the list entries placed before it (goal sensors count as obstacles). ---- */
#define MG_MAX_PLACE_ENTS 16
#define MG_MAX_PLACE_BODIES 6
typedef struct {
  int32_t kind;      /* 0: bodies that move rigidly together (robot: body, control, eyes, fingers); 1: goal region */
  int32_t goal;      /* kind 1: goal index */
  int32_t n_bodies;  /* kind 0 */
  int32_t n_groups;  /* kind 0: collision groups carrying the entity's shapes */
  int32_t bodies[MG_MAX_PLACE_BODIES]; /* main body first */
  int32_t groups[4];
  int32_t rand_pos, rand_rot;
  double pos_limit;  /* l-infinity bound around orig (< 0: anywhere in the arena) */
  double rot_limit;  /* bound around orig[2] (< 0: [-pi, pi]) */
  double orig[3];    /* main body pose before randomisation; goal: TOP-LEFT corner x, y (entities.py:790-798) */
} mg_place_ent_t;
typedef struct {
  int32_t goal, pad_;
  double min_side, max_side, cur_h, cur_w;
  double linf;       /* bound around (cur_h, cur_w); < 0: none */
} mg_place_hw_t;
typedef struct {
  int32_t n_ents, n_hw;
  int32_t goal_prims[MG_MAX_GOALS][2]; /* fill and border primitive of every goal (world-space rectangles) */
  double arena[4];   /* l, r, b, t */
  mg_place_ent_t ents[MG_MAX_PLACE_ENTS];
  mg_place_hw_t hw[MG_MAX_GOALS];
} mg_placement_t;

/* One environment's COMPLETE simulator state in host-readable form: what mg_get_state returns and
 * mg_set_state restores (checkpoint / resume, parity tests).  Everything Chipmunk carries from one
 * cpSpaceStep to the next is here: poses, velocities, the pending bias velocities, the joints'
 * accumulated impulses and the arbiter cache (one entry per cached contact: shape pair, contact
 * hash, age of the pair's last collision in sub-steps, accumulated impulses). */
#define MG_STATE_CACHE 48
typedef struct {
  int32_t n_bodies, n_joints, n_contacts, episode_steps;
  int32_t scene, overflow;
  int32_t n_cache;  /* valid entries of cache_* */
  int32_t stamp;    /* sub-steps simulated since the episode's reset */
  double pos[MG_MAX_BODIES][2];
  double angle[MG_MAX_BODIES];
  double vel[MG_MAX_BODIES][2];
  double angvel[MG_MAX_BODIES];
  double joint_acc[MG_MAX_JOINTS][2]; /* accumulated impulses (jAcc / jnAcc) */
  /* contacts of the last sub-step (a view of the cache entries with age 0; ignored by mg_set_state) */
  int32_t contact_shapes[32][2];
  double contact_jn[32], contact_jt[32];
  /* pending bias velocities (cpBody v_bias / w_bias: consumed by the next position update) */
  double bias_vel[MG_MAX_BODIES][2];
  double bias_angvel[MG_MAX_BODIES];
  /* arbiter cache (Chipmunk keeps a pair's contacts for collision_persistence = 3 steps) */
  int32_t cache_shapes[MG_STATE_CACHE][2]; /* type-ordered shape indices */
  uint32_t cache_hash[MG_STATE_CACHE];     /* contact feature id */
  int32_t cache_age[MG_STATE_CACHE];       /* stamp - stamp of the pair's last collision: 0, 1 or 2 */
  double cache_jn[MG_STATE_CACHE], cache_jt[MG_STATE_CACHE];
} mg_state_t;

typedef struct mg_handle mg_handle;

int mg_version(void);
const char* mg_last_error(void);
/* sizeof(mg_scene_t) / sizeof(mg_state_t) as compiled, so a binding can check its struct layout. */
int64_t mg_sizeof_scene(void);
int64_t mg_sizeof_state(void);

/* Create `cfg->batch` environments on `cfg->device`.  `scenes` is a host array of
 * cfg->n_scenes compiled scenes (copied).  `cuda_stream` is a cudaStream_t (NULL = legacy
 * default stream).  Replaces BaseEnv.__init__ + Viewer creation (base_env.py:78-122,184-190). */
int mg_create(const mg_config_t* cfg, const mg_scene_t* scenes, void* cuda_stream, mg_handle** out);
int mg_destroy(mg_handle* h);

/* Bind the caller-owned DEVICE observation buffer (layout per cfg->obs_mode). */
int mg_bind_obs(mg_handle* h, void* obs_dev, int64_t nbytes);
int64_t mg_obs_nbytes(const mg_handle* h);
/* Two-plane layouts (LoResStack, RAW) with the planes `plane_stride` bytes apart instead of back to back, so a
 * rank's shard can be rendered straight into its slice of a larger [2, B_global, ...] tensor (multi-GPU, 8(e)). */
int mg_bind_obs_planes(mg_handle* h, void* plane0_dev, int64_t plane_nbytes, int64_t plane_stride);

/* Optional second output of the render: every environment's NEWEST frame alone, u8 [views, B, 96, 96, 3]
 * (views = 2 for LoResStack: allo then ego; LoRes4E / LoRes4A: 1).  It is the send buffer of the multi-GPU
 * observation all-gather: 27 648 B per environment and view instead of the 110 592 B stack.
 * NULL unbinds.  mg_newest_nbytes is 0 for layouts without this output (3EA, CHW4E, RAW). */
int64_t mg_newest_nbytes(const mg_handle* h);
int mg_bind_newest(mg_handle* h, void* newest_dev, int64_t nbytes);

/* FlattenFrameStack for environments rendered elsewhere (benchmarks/__init__.py:118-136): for every env in
 * [env_first, env_first + env_count) of `stacks_dev` (u8 [n, res, res, 12]) drop the oldest frame and append
 * the env's frame from `newest_dev`; where fresh_dev[env] != 0 (the env auto-reset in this step) all four
 * slots are filled with it.  The frame of env e is read at
 *   newest_dev + (e / shard) * newest_rank_stride + (e % shard) * res * res * 3
 * (the layout an all-gather of per-rank [views, shard, res, res, 3] buffers produces).  Handle-free: runs on
 * the current device on `cuda_stream`. */
int mg_stack_push(void* stacks_dev, const void* newest_dev, const uint8_t* fresh_dev, int64_t env_first,
                  int64_t env_count, int32_t shard, int64_t newest_rank_stride, int32_t res, void* cuda_stream);

/* Reset environments: BaseEnv.reset (base_env.py:177-234) + FlattenFrameStack.reset
 * (benchmarks/__init__.py:130-136).  env_ids: HOST int32[n] (NULL => all); scene_ids: HOST
 * int32[n] scene index per reset env (NULL => keep each env's current scene, initially 0).
 * Renders the first observation into the bound buffer (frame replicated over the stack). */
int mg_reset(mg_handle* h, const int32_t* env_ids, int32_t n, const int32_t* scene_ids);

/* Scene pool maintenance for the randomised variants (the reference samples a fresh layout on every reset,
 * base_env.py:177-234; here the host streams freshly sampled scenes into the pool while the GPU steps):
 * mg_update_scenes overwrites pool entries [first, first + n) -- the caller must make sure that no environment is
 * still playing them, e.g. by keeping them outside the draw range for at least one episode length;
 * mg_set_draw_range restricts the entries an auto-reset draws from (n = 0: keep each env on its scene).
 * The rasteriser's shared-memory layout grows with the largest scene seen; the physics kernel's layout is fixed
 * at mg_create, so a new scene with more bodies / shape groups than any scene of the original pool is rejected
 * (MG_E_INVALID, "exceeds the physics layout", nothing copied). */
int mg_update_scenes(mg_handle* h, int32_t first, int32_t n, const mg_scene_t* scenes);
int mg_set_draw_range(mg_handle* h, int32_t first, int32_t n);

/* Device-side reset randomisation (cfg->device_sampling; replaces geom.py:116-359 at reset time).
 * mg_set_placement uploads the placement programs of templates [first, first + n) (required before the first
 * reset; again after mg_update_scenes streams new templates).  With it, every reset -- mg_reset and the
 * auto-reset inside mg_step -- draws a template, samples goal sizes and poses with the reference's procedure
 * (uniform draws, same try / retry limits, same obstacle rules; Philox streams keyed by reset_seed, env and the
 * env's reset count: same distribution as the reference, not the same numpy stream) and plays the result from a
 * scene slot of its own.  mg_get_env_scene returns the scene an environment is playing (what the oracle needs to
 * reproduce the episode); mg_sampler_failures counts resets that fell back to the template's own host-sampled
 * poses because no placement was found within the reference's limits. */
int64_t mg_sizeof_placement(void);
int mg_set_placement(mg_handle* h, int32_t first, int32_t n, const mg_placement_t* programs);
int mg_get_env_scene(mg_handle* h, int32_t env, mg_scene_t* out);
int mg_sampler_failures(mg_handle* h, int64_t* out);

/* One env-step for the whole batch: Robot.set_action + 10 x (Robot.update + Space.step)
 * + episode bookkeeping + score + render + stack (base_env.py:255-292).
 * actions: DEVICE int32[batch] in [0,18).  reward/done/score: DEVICE, caller-owned, [batch]
 * (any may be NULL).  The observation lands in the bound buffer. */
int mg_step(mg_handle* h, const int32_t* actions_dev, float* reward_dev, uint8_t* done_dev,
            float* score_dev);

/* Physics only (no render): used by parity tests and the physics-only bench leg. */
int mg_step_physics(mg_handle* h, const int32_t* actions_dev, float* reward_dev,
                    uint8_t* done_dev, float* score_dev);
/* The render half of mg_step (mg_step == mg_step_physics + mg_step_render): rasterise the current state and
 * PUSH the frame onto the stacks.  Call exactly once per mg_step_physics. */
int mg_step_render(mg_handle* h);
/* Re-render the current state WITHOUT advancing the stacks: the newest frame is replaced in place (raw layout:
 * the frame is redrawn), so calling it any number of times leaves the next observation unchanged, like the
 * reference's env.render() (base_env.py:309-338). */
int mg_render(mg_handle* h);

/* Evaluate each env's end-of-trajectory score for its CURRENT state
 * (score_on_end_of_traj of the env's task); score_dev: DEVICE float[batch]. */
int mg_score(mg_handle* h, float* score_dev);

/* Host-readable snapshot / overwrite of one environment (synchronises the stream). */
int mg_get_state(mg_handle* h, int32_t env, mg_state_t* out);
/* Restore one environment from a snapshot taken with mg_get_state (of this or another handle running the
 * same scene): poses, velocities, bias velocities, joint accumulators, arbiter cache, episode step counter.
 * The environment keeps its scene binding; `in->n_bodies` / `in->n_joints` must match the scene (MG_E_INVALID
 * otherwise).  Rotations are re-derived from the angles with the library's sincos, so get -> set -> step
 * continues bit for bit.  The observation stack is not touched (call mg_render for a frame of the new state).
 * SURVEY 8(b) `mg_set_state`; no reference counterpart (pymunk spaces are pickled whole). */
int mg_set_state(mg_handle* h, int32_t env, const mg_state_t* in);
/* Poses of every body of every environment in one copy: out_host is HOST double [batch][MG_MAX_BODIES][4]
 * (x, y, angle, 0).  Synchronises the stream. */
int mg_get_poses(mg_handle* h, double* out_host);
int mg_set_pose(mg_handle* h, int32_t env, int32_t body, double x, double y, double angle);

/* Number of kernels launched by this handle since creation (bench bookkeeping). */
int64_t mg_launch_count(const mg_handle* h);
int mg_synchronize(mg_handle* h);

/* Environments x episodes in which the physics hit a capacity limit since mg_create (more than 32 simultaneous
 * solver contacts or 48 cached contacts: that sub-step is then solved without contacts / with a truncated cache,
 * and mg_state_t.overflow of the environment is non-zero until its next reset).  0 on every workload measured;
 * a caller that needs the reference's behaviour bit for bit checks this after a run.  Synchronises the stream.
 * (No reference counterpart: Chipmunk's arrays grow.) */
int mg_overflow_count(mg_handle* h, int64_t* out);

/* ---------------------------------------------------------------------------------------------------------
 * Multi-GPU transport over NVLink peer memory (one process per GPU on one node).
 * Each rank owns a region its peers map through CUDA IPC: barrier flags + `n_buffers` send buffers, each
 * holding the rank's packed scalars and its newest frames ([views, shard, res, res, 3]).  Bind
 * mg_comm_frame_ptr(c, b) with mg_bind_newest and let mg_step write reward / score / done into
 * mg_comm_scalar_ptr(c, b) (layout is the caller's; the library only moves the bytes).  Per step:
 *     mg_comm_barrier           every rank's send buffer b is complete and visible
 *     mg_comm_gather_scalars    dst[world][scalar_bytes] <- every rank's scalars
 *     mg_comm_stack_push        FlattenFrameStack of remote environments, frames read from the owners'
 *                               buffers over NVLink inside the same kernel (all-gather fused with the rebuild)
 * Rendezvous: mg_comm_export gives MG_COMM_HANDLE_BYTES bytes to exchange by any means (the Python host uses
 * torch.distributed.all_gather); mg_comm_connect takes the world's handles in rank order.
 * The barrier kernel gives up after 20 s (a dead peer must not hang the GPU); mg_comm_error then reports 1. */
#define MG_COMM_HANDLE_BYTES 64
typedef struct mg_comm mg_comm;
int mg_comm_create(int32_t rank, int32_t world, int32_t n_buffers, int64_t scalar_bytes, int64_t frame_bytes,
                   mg_comm** out);
int mg_comm_export(mg_comm* c, void* handle_out);
int mg_comm_connect(mg_comm* c, const void* all_handles);
void* mg_comm_scalar_ptr(mg_comm* c, int32_t buffer);
void* mg_comm_frame_ptr(mg_comm* c, int32_t buffer);
int mg_comm_barrier(mg_comm* c, void* cuda_stream);
int mg_comm_gather_scalars(mg_comm* c, int32_t buffer, void* dst_dev, void* cuda_stream);
/* stacks_dev: u8 [n_global, res, res, 12] (one view plane); the frame of env e is read from rank e / shard at
 * byte view_offset + (e % shard) * res * res * 3 of its frame buffer `buffer`.  env_modulo = world * shard makes
 * the range [env_first, env_first + env_count) wrap around: rank r passes env_first = (r + 1) * shard and
 * env_count = (world - 1) * shard and so reads its peers in the order r+1, r+2, ... -- every rank starts at a
 * different owner, no GPU's NVLink egress serves all readers at once.  env_modulo = 0: no wrap. */
int mg_comm_stack_push(mg_comm* c, int32_t buffer, int64_t view_offset, void* stacks_dev, const uint8_t* fresh_dev,
                       int64_t env_first, int64_t env_count, int64_t env_modulo, int32_t shard, int32_t res,
                       void* cuda_stream);
int mg_comm_error(mg_comm* c, int32_t* out);
int mg_comm_destroy(mg_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* MAGICAL_B200_H */
