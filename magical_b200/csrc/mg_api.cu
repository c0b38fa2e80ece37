/*
 * mg_api.cu — the C ABI declared in include/magical_b200.h (host side).
 *
 * One handle = one batch of environments on one GPU.  It owns the compiled
 * scenes (+ derived constants) and the per-environment state records in HBM,
 * and enqueues three kernels per env-step on the caller's stream:
 *   K1 k_physics  (mg_physics.cu)  10 sub-steps of rigid-body physics
 *   K2 k_finish   (mg_finish.cu)   step counter, done, score, reward, auto-reset
 *   K3 k_raster   (mg_raster.cu)   render + downsample + frame stack into the bound obs buffer
 * Reference call sites replaced: base_env.py:177-338 (reset/step/render).
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "mg_device.cuh"
#include "mg_physics_tpe.h"
#include "mg_sincos.h"
#include "mg_state_io.h"

cudaError_t mg_launch_physics(EnvState* states, const DeviceScene* scenes, const int32_t* actions, int batch,
                              int lanes_per_env, int block_threads, cudaStream_t stream);
cudaError_t mg_launch_physics_tpe(EnvState* states, const DeviceScene* scenes, const int32_t* actions, int env0,
                                  int count, const TpeLayout* L, double* spill, uint32_t* scratch,
                                  cudaStream_t stream);
size_t mg_tpe_smem_bytes(const TpeLayout* L);
size_t mg_tpe_spill_doubles_per_env(const TpeLayout* L);
cudaError_t mg_launch_finish(EnvState* states, const DeviceScene* scenes, int env0, int count, int auto_reset, int mode,
                             int draw_first, int draw_count, uint32_t reset_seed, int slot_base, float* reward,
                             uint8_t* done, float* score, unsigned long long* overflow_count, cudaStream_t stream);
cudaError_t mg_launch_reset(EnvState* states, const DeviceScene* scenes, int n, const int32_t* env_ids,
                            const int32_t* scene_ids, int first_time, int sample_mode, cudaStream_t stream);
cudaError_t mg_launch_sample_layouts(EnvState* states, DeviceScene* scenes, const mg_placement_t* programs,
                                     int n_templates, int batch, uint32_t seed, unsigned long long* failures,
                                     cudaStream_t stream);
cudaError_t mg_launch_raster(int mode, EnvState* states, const DeviceScene* scenes, uint8_t* obs, uint8_t* newest,
                             size_t plane_stride, int batch, int res_out, int ecap, int scap, int rcap, int only_fresh,
                             int push, int env0, int count, int slot_base, cudaStream_t stream);
cudaError_t mg_launch_stack_push(uint8_t* stacks, const uint8_t* newest, const uint8_t* fresh, long long env_first,
                                 long long env_count, int shard, long long rank_stride, int res, cudaStream_t stream);
cudaError_t mg_raster_upload_units(const double* units);
size_t mg_raster_smem_bytes(int mode, int ecap, int scap, int rcap);

struct mg_handle {
  mg_config_t cfg;
  cudaStream_t stream;
  EnvState* d_states;
  DeviceScene* d_scenes;
  int32_t* d_ids;       /* scratch for mg_reset */
  int32_t* d_scene_ids;
  uint8_t* obs;
  int64_t obs_bytes;
  int64_t plane_stride; /* bytes between the two view planes of obs (LoResStack / RAW) */
  uint8_t* newest;      /* optional second output: every environment's newest frame (mg_bind_newest) */
  int res_out;
  int ecap;
  int scap;             /* span-table rows the rasteriser reserves per view */
  int rcap;             /* window-space primitives (draw prims + expanded line segments) it reserves */
  int lanes_per_env;    /* 16: two environments share a warp in K1; 32: one warp per environment */
  int block_threads;    /* K1 block size: 512 = one phase-aligned block per SM, 128 = small blocks */
  int use_tpe;          /* K1 variant: 1 = thread per environment (mg_physics_tpe.h), 0 = lanes per environment */
  TpeLayout tpe;        /* private-word layout of the thread-per-environment kernel */
  double* d_spill;      /* its contact spill area */
  uint32_t* d_scratch;  /* its per-environment work-item / separation-cache records (scratch_global layout) */
  unsigned long long* d_overflow; /* environments x episodes that hit a physics capacity limit (k_finish counts) */
  /* device-side layout sampling (cfg.device_sampling): d_scenes = [n_scenes templates | batch slots] */
  mg_placement_t* d_programs;
  unsigned long long* d_failures;
  int programs_set;
  /* the auto-reset sampler runs on its own stream, next to the render of the environments that did not reset
   * (one warp per resetting environment is latency-bound: ~1.5 ms for a handful of warps) */
  cudaStream_t sample_stream;
  cudaEvent_t ev_finish, ev_sampled;
  int sample_pending;
  /* mg_step software pipeline: the batch is cut into chunks whose physics and raster kernels run on two
   * internal streams, staggered so that the raster of chunk c overlaps the physics of chunk c + 1 (the
   * physics is latency-bound at ~13 % issue utilisation, the raster issue-bound: they share SMs well) */
  int draw_first, draw_count; /* pool entries an auto-reset draws from (mg_set_draw_range) */
  int n_chunks;
  cudaStream_t side[2];
  cudaEvent_t ev_start, ev_phys[2], ev_done[2];
  int64_t launches;
};

static thread_local char g_err[512] = "";

void mg_set_error_(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

static int fail(int code, const char* fmt, const char* detail) {
  snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
  return code;
}
#define CUDA_TRY(expr)                                                      \
  do {                                                                      \
    cudaError_t e_ = (expr);                                                \
    if (e_ != cudaSuccess) return fail(MG_E_CUDA, #expr ": %s", cudaGetErrorString(e_)); \
  } while (0)

/* what the rasteriser must reserve for one scene: polygon edges, window-space primitives, span-table rows */
static const char* scene_raster_needs(const mg_scene_t& sc, int res_full, int* edges_out, int* rprims_out,
                                      int* rows_out) {
  const double S_full = (double)res_full / 2.04;
  int edges = 0, rprims = 0, rows = 0;
  for (int p = 0; p < sc.n_prims; p++) {
    const mg_prim_t& pr = sc.prims[p];
    edges += pr.nvert;
    rprims += pr.kind == MG_PRIM_LINELOOP ? pr.nvert : 1;
    /* span-table rows: a primitive is rigid, so in any camera it spans at most its diameter */
    if (pr.kind == MG_PRIM_NGON) {
      int r = (int)ceil(2.0 * pr.radius * S_full) + 14; /* box margins + the aligned allocation's padding */
      rows += r < res_full ? r : res_full;
    } else if ((int)pr.vert0 + pr.nvert <= MG_MAX_DVERTS) {
      const float(*dv)[2] = &sc.dverts[pr.vert0];
      if (pr.kind == MG_PRIM_LINELOOP) {
        for (int k = 0; k < pr.nvert; k++) {
          int k2 = (k + 1) % pr.nvert;
          double len = hypot((double)dv[k][0] - dv[k2][0], (double)dv[k][1] - dv[k2][1]);
          int r = (int)ceil(len * S_full + pr.radius) + 14;
          rows += r < res_full ? r : res_full;
        }
      } else {
        double diam = 0.0;
        for (int a = 0; a < pr.nvert; a++)
          for (int b = a + 1; b < pr.nvert; b++) {
            double d = hypot((double)dv[a][0] - dv[b][0], (double)dv[a][1] - dv[b][1]);
            if (d > diam) diam = d;
          }
        int r = (int)ceil(diam * S_full) + 14;
        rows += r < res_full ? r : res_full;
      }
    }
    if (pr.kind == MG_PRIM_NGON && pr.nvert != 10 && pr.nvert != 20 && pr.nvert != 100)
      return "NGON primitives must have 10, 20 or 100 sides";
    if (pr.kind != MG_PRIM_NGON && (int)pr.vert0 + pr.nvert > MG_MAX_DVERTS) return "draw vertex range";
  }
  if (rprims > 192) return "too many draw primitives in one scene";
  *edges_out = edges; *rprims_out = rprims; *rows_out = rows;
  return nullptr;
}

extern "C" {

int mg_version(void) { return MG_ABI_VERSION; }
const char* mg_last_error(void) { return g_err; }
int64_t mg_sizeof_scene(void) { return (int64_t)sizeof(mg_scene_t); }
int64_t mg_sizeof_state(void) { return (int64_t)sizeof(mg_state_t); }

static int64_t obs_bytes_for(const mg_config_t* cfg, int res_out) {
  int64_t px = (int64_t)res_out * res_out;
  switch (cfg->obs_mode) {
    case MG_OBS_LORES4E:
    case MG_OBS_LORES4A:
    case MG_OBS_LORES3EA:
    case MG_OBS_LORESCHW4E:
      return (int64_t)cfg->batch * px * 12;
    case MG_OBS_LORESSTACK:
      return 2 * (int64_t)cfg->batch * px * 12;
    case MG_OBS_RAW:
      return 2 * (int64_t)cfg->batch * px * 3;
  }
  return -1;
}

int mg_create(const mg_config_t* cfg, const mg_scene_t* scenes, void* cuda_stream, mg_handle** out) {
  if (!cfg || !scenes || !out) return fail(MG_E_INVALID, "mg_create: null argument%s", "");
  if (cfg->batch <= 0 || cfg->n_scenes <= 0) return fail(MG_E_INVALID, "mg_create: batch and n_scenes must be > 0%s", "");
  if (cfg->obs_mode < MG_OBS_LORES4E || cfg->obs_mode > MG_OBS_RAW) return fail(MG_E_INVALID, "mg_create: bad obs_mode%s", "");
  int res_out = 96;
  if (cfg->obs_mode == MG_OBS_RAW) {
    res_out = cfg->res > 0 ? cfg->res : 384;
    if (res_out % 48 != 0) return fail(MG_E_INVALID, "mg_create: raw resolution must be a multiple of 48%s", "");
  }
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MG_E_INVALID, "mg_create: no such CUDA device%s", "");
  CUDA_TRY(cudaSetDevice(cfg->device));
  {
    /* the library holds sm_100a SASS only (no PTX): say so instead of failing at the first launch */
    int major = 0, minor = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
    CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, cfg->device));
    if (major != 10 || minor != 0)
      return fail(MG_E_INVALID, "mg_create: this build contains sm_100a (B200) code only; the selected device is not sm_100%s", "");
  }

  /* derive per-scene constants on the host */
  std::vector<DeviceScene> host(cfg->n_scenes);
  int ecap = 64, scap = 64, rcap = 32;
  const int ss = cfg->obs_mode == MG_OBS_RAW ? 1 : 4;
  const int res_full = res_out * ss;
  for (int i = 0; i < cfg->n_scenes; i++) {
    host[i].s = scenes[i];
    const char* why = mg_build_scene_aux(&scenes[i], &host[i].aux);
    if (why) return fail(MG_E_INVALID, "mg_create: scene rejected: %s", why);
    why = mg_build_raster_aux(&scenes[i], &host[i].ra);
    if (why) return fail(MG_E_INVALID, "mg_create: scene rejected: %s", why);
    int edges = 0, rprims = 0, rows = 0;
    const char* bad = scene_raster_needs(scenes[i], res_full, &edges, &rprims, &rows);
    if (bad) return fail(MG_E_INVALID, "mg_create: %s", bad);
    /* device-side sampling re-draws goal sizes (up to 0.8 a side, base_env.py:RAND_GOAL_MAX_SIZE): room for
     * the largest fill (its diagonal) and four border segments per goal, whatever the template's size is */
    if (cfg->device_sampling) rows += 900 * ss * scenes[i].n_goals / 4;
    if (edges > ecap) ecap = edges;
    if (rows > scap) scap = rows;
    if (rprims > rcap) rcap = rprims;
  }
  ecap = (ecap + 63) / 64 * 64;
  scap = (scap + 63) / 64 * 64;
  if (scap < 2 * ecap) scap = 2 * ecap; /* the window-space vertices are staged inside the span table */
  rcap = (rcap + 31) / 32 * 32;
  /* K1 serves one environment with 16 lanes when every schedule level fits (it does for all registered
   * tasks), which doubles the environments in flight per SM; MG_LANES_PER_ENV=32 forces a full warp */
  int lanes = 16;
  for (int i = 0; i < cfg->n_scenes; i++)
    if (host[i].aux.max_per_level > 16) lanes = 32;
  /* the cooperative fallback kernel pairs two environments per warp only when they run the same scene */
  if (cfg->n_scenes > 1) lanes = 32;
  if (const char* ev = getenv("MG_LANES_PER_ENV")) {
    int v = atoi(ev);
    if (v == 32 || (v == 16 && lanes == 16)) lanes = v;
  }
  if (mg_raster_smem_bytes(cfg->obs_mode, ecap, scap, rcap) > 200 * 1024)
    return fail(MG_E_INVALID, "mg_create: scene has too many draw edges for the rasteriser's shared memory%s", "");
  if (getenv("MG_VERBOSE"))
    fprintf(stderr, "mg_create: raster capacities: %d edges, %d span rows, %d primitives -> %zu B shared memory per CTA\n",
            ecap, scap, rcap, mg_raster_smem_bytes(cfg->obs_mode, ecap, scap, rcap));

  mg_handle* h = new (std::nothrow) mg_handle();
  if (!h) return fail(MG_E_NOMEM, "mg_create: out of host memory%s", "");
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->stream = (cudaStream_t)cuda_stream;
  h->res_out = res_out;
  h->ecap = ecap;
  h->scap = scap;
  h->rcap = rcap;
  h->lanes_per_env = lanes;
  /* phase-aligned 512-thread blocks need enough environments to fill the 148 SMs */
  h->block_threads = (cfg->n_scenes == 1 && cfg->batch / (512 / lanes) >= 2 * 148) ? 512 : 128;
  if (const char* ev = getenv("MG_BLOCK_THREADS")) {
    int v = atoi(ev);
    if (v == 128 || v == 512) h->block_threads = v;
  }
  h->obs_bytes = obs_bytes_for(cfg, res_out);
  h->draw_first = 0;
  h->draw_count = cfg->keep_scene ? 0 : cfg->n_scenes;
  /* K1 variant: every MAGICAL scene has the canonical robot + drag-jointed blocks structure and runs one
   * environment per thread; anything else (hand-built scenes) falls back to the cooperative kernel */
  h->use_tpe = 1;
  int max_slots = 0, max_blocks = 0, max_groups = 0, max_pairs = 0;
  for (int i = 0; i < cfg->n_scenes; i++) {
    if (!host[i].aux.tpe_ok) h->use_tpe = 0;
    if (host[i].aux.tpe_nslots > max_slots) max_slots = host[i].aux.tpe_nslots;
    if (host[i].aux.tpe_nblocks > max_blocks) max_blocks = host[i].aux.tpe_nblocks;
    if (scenes[i].n_cgroups > max_groups) max_groups = scenes[i].n_cgroups;
    if (scenes[i].n_bpairs > max_pairs) max_pairs = scenes[i].n_bpairs;
  }
  if (const char* ev = getenv("MG_PHYSICS")) {
    if (!strcmp(ev, "warp")) h->use_tpe = 0;
  }
  if (!h->use_tpe && cfg->n_scenes > 1) {
    /* the cooperative fallback kernel is validated for one shared scene only */
    delete h;
    return fail(MG_E_INVALID, "mg_create: scene pools need scenes with MAGICAL's canonical structure "
                              "(thread-per-environment kernel)%s", "");
  }
  if (h->use_tpe) {
    int kcon = 4;
    if (const char* ev = getenv("MG_TPE_KCON")) kcon = atoi(ev);
    if (kcon < 1) kcon = 1;
    if (kcon > TPE_MAX_CONTACTS) kcon = TPE_MAX_CONTACTS;
    int nitems = 48;
    if (const char* ev = getenv("MG_TPE_NITEMS")) nitems = atoi(ev);
    int scratch_global = 1;
    if (const char* ev = getenv("MG_TPE_SCRATCH")) scratch_global = strcmp(ev, "smem") != 0;
    h->tpe = tpe_make_layout(max_slots, max_blocks, max_groups, max_pairs, kcon, nitems, scratch_global);
    while (mg_tpe_smem_bytes(&h->tpe) > 200 * 1024 && kcon > 1)
      h->tpe = tpe_make_layout(max_slots, max_blocks, max_groups, max_pairs, --kcon, nitems, scratch_global);
    /* residency experiments: unused words at the end of every environment's private area */
    if (const char* ev = getenv("MG_TPE_PAD_WORDS")) h->tpe.words += atoi(ev);
  }
  cudaError_t e;
  if ((e = cudaMalloc(&h->d_states, sizeof(EnvState) * (size_t)cfg->batch)) != cudaSuccess ||
      (e = cudaMalloc(&h->d_scenes, sizeof(DeviceScene) * ((size_t)cfg->n_scenes +
                                                             (cfg->device_sampling ? (size_t)cfg->batch : 0)))) != cudaSuccess ||
      (cfg->device_sampling &&
       ((e = cudaMalloc(&h->d_programs, sizeof(mg_placement_t) * (size_t)cfg->n_scenes)) != cudaSuccess ||
        (e = cudaMemset(h->d_programs, 0, sizeof(mg_placement_t) * (size_t)cfg->n_scenes)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_failures, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemset(h->d_failures, 0, sizeof(unsigned long long))) != cudaSuccess)) ||
      (e = cudaMalloc(&h->d_ids, sizeof(int32_t) * (size_t)cfg->batch)) != cudaSuccess ||
      (e = cudaMalloc(&h->d_scene_ids, sizeof(int32_t) * (size_t)cfg->batch)) != cudaSuccess ||
      (e = cudaMalloc(&h->d_overflow, sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(h->d_overflow, 0, sizeof(unsigned long long))) != cudaSuccess ||
      (h->use_tpe && mg_tpe_spill_doubles_per_env(&h->tpe) > 0 &&
       (e = cudaMalloc(&h->d_spill, sizeof(double) * mg_tpe_spill_doubles_per_env(&h->tpe) * ((size_t)cfg->batch + 64))) !=
           cudaSuccess) ||
      (h->use_tpe && h->tpe.scratch_global &&
       (e = cudaMalloc(&h->d_scratch, sizeof(uint32_t) * (size_t)h->tpe.scratch_u32 * ((size_t)cfg->batch + 32))) !=
           cudaSuccess)) {
    mg_destroy(h);
    return fail(MG_E_NOMEM, "mg_create: cudaMalloc: %s", cudaGetErrorString(e));
  }
  /* measured on B200 (ClusterColour, 65536 envs): 1 chunk 23.8 ms, 2: 24.0, 4: 25.7, 8: 29.1 per env-step --
   * both kernels are shared-memory-capacity bound per SM, so co-residency costs more than the overlap
   * gains; the pipeline stays available through MG_CHUNKS but is off by default */
  h->n_chunks = 1;
  if (const char* ev = getenv("MG_CHUNKS")) {
    int v = atoi(ev);
    if (v >= 1 && v <= 16 && h->use_tpe) h->n_chunks = v;
  }
  if (h->n_chunks > 1) {
    bool ok = cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++)
      ok = cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_phys[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      mg_destroy(h);
      return fail(MG_E_CUDA, "mg_create: stream / event creation failed%s", "");
    }
  }
  if (cfg->device_sampling &&
      /* highest priority: its few blocks must get the first SM slot the render of the others frees, not the
       * last one (block scheduling is otherwise first launched, first placed) */
      (cudaStreamCreateWithPriority(&h->sample_stream, cudaStreamNonBlocking, -5) != cudaSuccess ||
       cudaEventCreateWithFlags(&h->ev_finish, cudaEventDisableTiming) != cudaSuccess ||
       cudaEventCreateWithFlags(&h->ev_sampled, cudaEventDisableTiming) != cudaSuccess)) {
    mg_destroy(h);
    return fail(MG_E_CUDA, "mg_create: stream / event creation failed%s", "");
  }
  if ((e = cudaMemcpyAsync(h->d_scenes, host.data(), sizeof(DeviceScene) * (size_t)cfg->n_scenes, cudaMemcpyHostToDevice,
                           h->stream)) != cudaSuccess ||
      (cfg->device_sampling && /* slots are empty scenes until the first reset samples them */
       (e = cudaMemsetAsync(h->d_scenes + cfg->n_scenes, 0, sizeof(DeviceScene) * (size_t)cfg->batch, h->stream)) !=
           cudaSuccess) ||
      (e = cudaMemsetAsync(h->d_states, 0, sizeof(EnvState) * (size_t)cfg->batch, h->stream)) != cudaSuccess) {
    mg_destroy(h);
    return fail(MG_E_CUDA, "mg_create: upload: %s", cudaGetErrorString(e));
  }
  /* unit circles for 10/20/100-gons: gym_render.make_circle (gym_render.py:438-446) */
  double units[130][2];
  const int ns[3] = {10, 20, 100}, offs[3] = {0, 10, 30};
  for (int t = 0; t < 3; t++)
    for (int k = 0; k < ns[t]; k++) {
      double ang = 2 * M_PI * k / ns[t];
      units[offs[t] + k][0] = cos(ang);
      units[offs[t] + k][1] = sin(ang);
    }
  if ((e = mg_raster_upload_units(&units[0][0])) != cudaSuccess ||
      (e = mg_launch_reset(h->d_states, h->d_scenes, cfg->batch, nullptr, nullptr, 1, 0, h->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(h->stream)) != cudaSuccess) {
    mg_destroy(h);
    return fail(MG_E_CUDA, "mg_create: init: %s", cudaGetErrorString(e));
  }
  h->launches = 1;
  *out = h;
  return MG_OK;
}

int mg_destroy(mg_handle* h) {
  if (!h) return MG_OK;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_states);
  cudaFree(h->d_scenes);
  cudaFree(h->d_ids);
  cudaFree(h->d_scene_ids);
  cudaFree(h->d_spill);
  cudaFree(h->d_scratch);
  cudaFree(h->d_overflow);
  cudaFree(h->d_programs);
  cudaFree(h->d_failures);
  if (h->sample_stream) { cudaStreamSynchronize(h->sample_stream); cudaStreamDestroy(h->sample_stream); }
  if (h->ev_finish) cudaEventDestroy(h->ev_finish);
  if (h->ev_sampled) cudaEventDestroy(h->ev_sampled);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  for (int i = 0; i < 2; i++) {
    if (h->side[i]) { cudaStreamSynchronize(h->side[i]); cudaStreamDestroy(h->side[i]); }
    if (h->ev_phys[i]) cudaEventDestroy(h->ev_phys[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
  }
  delete h;
  return MG_OK;
}

int mg_bind_obs(mg_handle* h, void* obs_dev, int64_t nbytes) {
  if (!h || !obs_dev) return fail(MG_E_INVALID, "mg_bind_obs: null argument%s", "");
  if (nbytes != h->obs_bytes) return fail(MG_E_INVALID, "mg_bind_obs: buffer size does not match the observation layout%s", "");
  if (((uintptr_t)obs_dev & 15) != 0) return fail(MG_E_INVALID, "mg_bind_obs: buffer must be 16-byte aligned%s", "");
  h->obs = (uint8_t*)obs_dev;
  h->plane_stride = nbytes / 2; /* only read by the two-plane layouts */
  return MG_OK;
}

int mg_bind_obs_planes(mg_handle* h, void* plane0_dev, int64_t plane_nbytes, int64_t plane_stride) {
  if (!h || !plane0_dev) return fail(MG_E_INVALID, "mg_bind_obs_planes: null argument%s", "");
  if (h->cfg.obs_mode != MG_OBS_LORESSTACK && h->cfg.obs_mode != MG_OBS_RAW)
    return fail(MG_E_INVALID, "mg_bind_obs_planes: the observation layout has a single plane (use mg_bind_obs)%s", "");
  if (plane_nbytes * 2 != h->obs_bytes) return fail(MG_E_INVALID, "mg_bind_obs_planes: plane size does not match the layout%s", "");
  if (plane_stride < plane_nbytes || (plane_stride & 15) != 0 || ((uintptr_t)plane0_dev & 15) != 0)
    return fail(MG_E_INVALID, "mg_bind_obs_planes: planes must not overlap and must be 16-byte aligned%s", "");
  h->obs = (uint8_t*)plane0_dev;
  h->plane_stride = plane_stride;
  return MG_OK;
}

int64_t mg_newest_nbytes(const mg_handle* h) {
  if (!h) return -1;
  switch (h->cfg.obs_mode) {
    case MG_OBS_LORES4E:
    case MG_OBS_LORES4A: return (int64_t)h->cfg.batch * h->res_out * h->res_out * 3;
    case MG_OBS_LORESSTACK: return 2 * (int64_t)h->cfg.batch * h->res_out * h->res_out * 3;
  }
  return 0; /* layout has no newest-frame output */
}

int mg_bind_newest(mg_handle* h, void* newest_dev, int64_t nbytes) {
  if (!h) return fail(MG_E_INVALID, "mg_bind_newest: null handle%s", "");
  if (!newest_dev) { h->newest = nullptr; return MG_OK; }
  if (mg_newest_nbytes(h) <= 0)
    return fail(MG_E_INVALID, "mg_bind_newest: only the LoRes4E / LoRes4A / LoResStack layouts have a newest-frame output%s", "");
  if (nbytes != mg_newest_nbytes(h)) return fail(MG_E_INVALID, "mg_bind_newest: buffer size does not match%s", "");
  if (((uintptr_t)newest_dev & 3) != 0) return fail(MG_E_INVALID, "mg_bind_newest: buffer must be 4-byte aligned%s", "");
  h->newest = (uint8_t*)newest_dev;
  return MG_OK;
}

int mg_stack_push(void* stacks_dev, const void* newest_dev, const uint8_t* fresh_dev, int64_t env_first,
                  int64_t env_count, int32_t shard, int64_t newest_rank_stride, int32_t res, void* cuda_stream) {
  if (!stacks_dev || !newest_dev) return fail(MG_E_INVALID, "mg_stack_push: null argument%s", "");
  if (env_first < 0 || env_count < 0 || shard <= 0 || res <= 0 || res % 4 != 0 || newest_rank_stride < 0)
    return fail(MG_E_INVALID, "mg_stack_push: bad range%s", "");
  if (((uintptr_t)stacks_dev & 15) != 0 || ((uintptr_t)newest_dev & 3) != 0 || (newest_rank_stride & 3) != 0)
    return fail(MG_E_INVALID, "mg_stack_push: stacks must be 16-byte and frames 4-byte aligned%s", "");
  CUDA_TRY(mg_launch_stack_push((uint8_t*)stacks_dev, (const uint8_t*)newest_dev, fresh_dev, env_first, env_count, shard,
                                newest_rank_stride, res, (cudaStream_t)cuda_stream));
  return MG_OK;
}

int64_t mg_obs_nbytes(const mg_handle* h) { return h ? h->obs_bytes : -1; }

/* the caller's stream waits for an auto-reset sampler that is still running on the side stream */
static int join_sampler(mg_handle* h) {
  if (h->sample_pending) {
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_sampled, 0));
    h->sample_pending = 0;
  }
  return MG_OK;
}

static int do_raster(mg_handle* h, int only_fresh, int push) {
  if (!h->obs) return fail(MG_E_STATE, "no observation buffer bound (call mg_bind_obs first)%s", "");
  const int slot_base = h->cfg.device_sampling ? h->cfg.n_scenes : -1;
  if (h->sample_pending && only_fresh == 0) {
    /* render everything that did NOT reset while the sampler is still placing the environments that did
     * (only_fresh = 2 skips them), then their first frames */
    CUDA_TRY(mg_launch_raster(h->cfg.obs_mode, h->d_states, h->d_scenes, h->obs, h->newest, (size_t)h->plane_stride,
                              h->cfg.batch, h->res_out, h->ecap, h->scap, h->rcap, 2, push, 0, h->cfg.batch, slot_base,
                              h->stream));
    h->launches++;
    only_fresh = 1;
  }
  int rc = join_sampler(h);
  if (rc != MG_OK) return rc;
  CUDA_TRY(mg_launch_raster(h->cfg.obs_mode, h->d_states, h->d_scenes, h->obs, h->newest, (size_t)h->plane_stride,
                            h->cfg.batch, h->res_out, h->ecap, h->scap, h->rcap, only_fresh, push, 0, h->cfg.batch,
                            slot_base, h->stream));
  h->launches++;
  return MG_OK;
}

int mg_reset(mg_handle* h, const int32_t* env_ids, int32_t n, const int32_t* scene_ids) {
  if (!h) return fail(MG_E_INVALID, "mg_reset: null handle%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  if (!env_ids) n = h->cfg.batch;
  if (n < 0 || n > h->cfg.batch) return fail(MG_E_INVALID, "mg_reset: bad env count%s", "");
  if (n == 0) return MG_OK;
  if (env_ids) {
    for (int i = 0; i < n; i++)
      if (env_ids[i] < 0 || env_ids[i] >= h->cfg.batch) return fail(MG_E_INVALID, "mg_reset: env id out of range%s", "");
    CUDA_TRY(cudaMemcpyAsync(h->d_ids, env_ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  }
  if (scene_ids) {
    for (int i = 0; i < n; i++)
      if (scene_ids[i] < 0 || scene_ids[i] >= h->cfg.n_scenes)
        return fail(MG_E_INVALID, "mg_reset: scene id out of range%s", "");
    CUDA_TRY(cudaMemcpyAsync(h->d_scene_ids, scene_ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  }
  if (h->cfg.device_sampling && (!scene_ids || !h->programs_set))
    return fail(MG_E_STATE, "mg_reset: device-side layout sampling needs template ids (scene_ids) and mg_set_placement first%s", "");
  CUDA_TRY(mg_launch_reset(h->d_states, h->d_scenes, n, env_ids ? h->d_ids : nullptr,
                           scene_ids ? h->d_scene_ids : nullptr, 0, h->cfg.device_sampling, h->stream));
  h->launches++;
  if (h->cfg.device_sampling) {
    CUDA_TRY(mg_launch_sample_layouts(h->d_states, h->d_scenes, h->d_programs, h->cfg.n_scenes, h->cfg.batch,
                                      (uint32_t)h->cfg.reset_seed, h->d_failures, h->stream));
    h->launches++;
  }
  /* the host arrays may be reused by the caller as soon as we return */
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->obs) return do_raster(h, 1, 1);
  return MG_OK;
}

static int do_physics(mg_handle* h, const int32_t* actions_dev, float* reward_dev, uint8_t* done_dev, float* score_dev) {
  if (!actions_dev) return fail(MG_E_INVALID, "mg_step: null actions%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc = join_sampler(h); if (rc != MG_OK) return rc; }
  if (h->use_tpe)
    CUDA_TRY(mg_launch_physics_tpe(h->d_states, h->d_scenes, actions_dev, 0, h->cfg.batch, &h->tpe, h->d_spill,
                                   h->d_scratch, h->stream));
  else
    CUDA_TRY(mg_launch_physics(h->d_states, h->d_scenes, actions_dev, h->cfg.batch, h->lanes_per_env, h->block_threads,
                               h->stream));
  CUDA_TRY(mg_launch_finish(h->d_states, h->d_scenes, 0, h->cfg.batch, h->cfg.auto_reset, 0, h->draw_first, h->draw_count,
                            (uint32_t)h->cfg.reset_seed, h->cfg.device_sampling ? h->cfg.n_scenes : -1, reward_dev, done_dev,
                            score_dev, h->d_overflow, h->stream));
  h->launches += 2;
  if (h->cfg.device_sampling && h->cfg.auto_reset) {
    /* environments that finished an episode get a freshly sampled layout, on the side stream: the render of
     * all the others (do_raster) runs next to it */
    CUDA_TRY(cudaEventRecord(h->ev_finish, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->sample_stream, h->ev_finish, 0));
    CUDA_TRY(mg_launch_sample_layouts(h->d_states, h->d_scenes, h->d_programs, h->cfg.n_scenes, h->cfg.batch,
                                      (uint32_t)h->cfg.reset_seed, h->d_failures, h->sample_stream));
    CUDA_TRY(cudaEventRecord(h->ev_sampled, h->sample_stream));
    h->sample_pending = 1;
    h->launches++;
  }
  return MG_OK;
}

/* mg_step as a two-stream software pipeline over chunks of the batch (see mg_handle::n_chunks).  Order on
 * the GPU: phys(0) | phys(1) + raster(0) | phys(2) + raster(1) | ... | raster(C-1).  Everything is fenced
 * against the caller's stream with events, so the call stays asynchronous and stream-ordered. */
static int step_pipelined(mg_handle* h, const int32_t* actions_dev, float* reward_dev, uint8_t* done_dev,
                          float* score_dev) {
  if (!actions_dev) return fail(MG_E_INVALID, "mg_step: null actions%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const int B = h->cfg.batch, C = h->n_chunks;
  /* chunk boundaries on multiples of 32 environments (one warp of the physics kernel) */
  const int per = ((B + C - 1) / C + 31) / 32 * 32;
  CUDA_TRY(cudaEventRecord(h->ev_start, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(h->side[0], h->ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->side[1], h->ev_start, 0));
  for (int c = 0; c * per < B; c++) {
    const int env0 = c * per, count = (env0 + per <= B) ? per : B - env0;
    cudaStream_t st = h->side[c & 1];
    /* stagger: the physics of chunk c starts when the physics of chunk c - 1 has finished */
    if (c > 0) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_phys[(c - 1) & 1], 0));
    CUDA_TRY(mg_launch_physics_tpe(h->d_states, h->d_scenes, actions_dev, env0, count, &h->tpe, h->d_spill, h->d_scratch,
                                   st));
    CUDA_TRY(cudaEventRecord(h->ev_phys[c & 1], st));
    CUDA_TRY(mg_launch_finish(h->d_states, h->d_scenes, env0, count, h->cfg.auto_reset, 0, h->draw_first, h->draw_count,
                              (uint32_t)h->cfg.reset_seed, -1, reward_dev, done_dev, score_dev, h->d_overflow, st));
    CUDA_TRY(mg_launch_raster(h->cfg.obs_mode, h->d_states, h->d_scenes, h->obs, h->newest, (size_t)h->plane_stride, B,
                              h->res_out, h->ecap, h->scap, h->rcap, 0, 1, env0, count, -1, st));
    h->launches += 3;
  }
  for (int i = 0; i < 2; i++) {
    CUDA_TRY(cudaEventRecord(h->ev_done[i], h->side[i]));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_done[i], 0));
  }
  return MG_OK;
}

int mg_step(mg_handle* h, const int32_t* actions_dev, float* reward_dev, uint8_t* done_dev, float* score_dev) {
  if (!h) return fail(MG_E_INVALID, "mg_step: null handle%s", "");
  if (!h->obs) return fail(MG_E_STATE, "mg_step: no observation buffer bound (call mg_bind_obs first)%s", "");
  if (h->n_chunks > 1 && !h->cfg.device_sampling) return step_pipelined(h, actions_dev, reward_dev, done_dev, score_dev);
  int rc = do_physics(h, actions_dev, reward_dev, done_dev, score_dev);
  if (rc != MG_OK) return rc;
  return do_raster(h, 0, 1);
}

int mg_step_physics(mg_handle* h, const int32_t* actions_dev, float* reward_dev, uint8_t* done_dev, float* score_dev) {
  if (!h) return fail(MG_E_INVALID, "mg_step_physics: null handle%s", "");
  return do_physics(h, actions_dev, reward_dev, done_dev, score_dev);
}

int mg_step_render(mg_handle* h) {
  if (!h) return fail(MG_E_INVALID, "mg_step_render: null handle%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  return do_raster(h, 0, 1);
}

int mg_render(mg_handle* h) {
  if (!h) return fail(MG_E_INVALID, "mg_render: null handle%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  return do_raster(h, 0, 0);
}

int mg_score(mg_handle* h, float* score_dev) {
  if (!h || !score_dev) return fail(MG_E_INVALID, "mg_score: null argument%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  CUDA_TRY(mg_launch_finish(h->d_states, h->d_scenes, 0, h->cfg.batch, 0, 1, 0, 1, 0u,
                            h->cfg.device_sampling ? h->cfg.n_scenes : -1, nullptr, nullptr, score_dev, nullptr, h->stream));
  h->launches++;
  return MG_OK;
}

int mg_get_state(mg_handle* h, int32_t env, mg_state_t* out) {
  if (!h || !out) return fail(MG_E_INVALID, "mg_get_state: null argument%s", "");
  if (env < 0 || env >= h->cfg.batch) return fail(MG_E_INVALID, "mg_get_state: env out of range%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  EnvState st;
  CUDA_TRY(cudaMemcpyAsync(&st, h->d_states + env, sizeof(EnvState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  DeviceScene* ds = (DeviceScene*)malloc(sizeof(DeviceScene));
  if (!ds) return fail(MG_E_NOMEM, "mg_get_state: out of host memory%s", "");
  cudaError_t e = cudaMemcpy(ds, h->d_scenes + st.scene, sizeof(DeviceScene), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(ds); return fail(MG_E_CUDA, "mg_get_state: %s", cudaGetErrorString(e)); }
  mg_state_export(st, ds->s, out);
  free(ds);
  return MG_OK;
}

int mg_set_state(mg_handle* h, int32_t env, const mg_state_t* in) {
  if (!h || !in) return fail(MG_E_INVALID, "mg_set_state: null argument%s", "");
  if (env < 0 || env >= h->cfg.batch) return fail(MG_E_INVALID, "mg_set_state: env out of range%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  EnvState st;
  CUDA_TRY(cudaMemcpyAsync(&st, h->d_states + env, sizeof(EnvState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  /* the snapshot must describe the scene this environment is bound to (header counts only are fetched) */
  int32_t hdr[16];
  CUDA_TRY(cudaMemcpy(hdr, &h->d_scenes[st.scene].s, sizeof(hdr), cudaMemcpyDeviceToHost));
  const mg_scene_t* sc = reinterpret_cast<const mg_scene_t*>(hdr);
  const char* why = mg_state_import(st, sc->n_bodies, sc->n_joints, sc->n_shapes, in);
  if (why) return fail(MG_E_INVALID, "mg_set_state: %s", why);
  CUDA_TRY(cudaMemcpy(h->d_states + env, &st, sizeof(EnvState), cudaMemcpyHostToDevice));
  return MG_OK;
}

int mg_set_pose(mg_handle* h, int32_t env, int32_t body, double x, double y, double angle) {
  if (!h) return fail(MG_E_INVALID, "mg_set_pose: null handle%s", "");
  if (env < 0 || env >= h->cfg.batch || body < 0 || body >= MG_MAX_BODIES)
    return fail(MG_E_INVALID, "mg_set_pose: index out of range%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  double4 P = make_double4(x, y, angle, 0.0);
  double sn, cs;
  mg_det_sincos(angle, &sn, &cs);
  double2 R = make_double2(cs, sn);
  EnvState* st = h->d_states + env;
  CUDA_TRY(cudaMemcpy(&st->P[body], &P, sizeof(P), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(&st->R[body], &R, sizeof(R), cudaMemcpyHostToDevice));
  return MG_OK;
}

int mg_update_scenes(mg_handle* h, int32_t first, int32_t n, const mg_scene_t* scenes) {
  if (!h || !scenes) return fail(MG_E_INVALID, "mg_update_scenes: null argument%s", "");
  if (first < 0 || n <= 0 || first + n > h->cfg.n_scenes) return fail(MG_E_INVALID, "mg_update_scenes: range%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  const int ss = h->cfg.obs_mode == MG_OBS_RAW ? 1 : 4;
  std::vector<DeviceScene> host(n);
  int ecap = h->ecap, scap = h->scap, rcap = h->rcap;
  for (int i = 0; i < n; i++) {
    host[i].s = scenes[i];
    const char* why = mg_build_scene_aux(&scenes[i], &host[i].aux);
    if (why) return fail(MG_E_INVALID, "mg_update_scenes: scene rejected: %s", why);
    why = mg_build_raster_aux(&scenes[i], &host[i].ra);
    if (why) return fail(MG_E_INVALID, "mg_update_scenes: scene rejected: %s", why);
    int edges = 0, rprims = 0, rows = 0;
    const char* bad = scene_raster_needs(scenes[i], h->res_out * ss, &edges, &rprims, &rows);
    if (bad) return fail(MG_E_INVALID, "mg_update_scenes: %s", bad);
    /* the rasteriser's shared-memory layout is a launch parameter: it grows with the largest scene seen */
    if (edges > ecap) ecap = edges;
    if (rows > scap) scap = rows;
    if (rprims > rcap) rcap = rprims;
    if (h->use_tpe) {
      const mg_scene_aux_t& ax = host[i].aux;
      const int con_words = h->tpe.scratch_global ? (h->tpe.words - h->tpe.off_con) : (h->tpe.off_it - h->tpe.off_con);
      if (!ax.tpe_ok || ax.tpe_nslots > h->tpe.nslots || ax.tpe_nblocks > h->tpe.nblocks ||
          scenes[i].n_cgroups * 3 > con_words || (scenes[i].n_bpairs + 1) / 2 + h->tpe.nitems > h->tpe.scratch_u32)
        return fail(MG_E_INVALID, "mg_update_scenes: scene exceeds the physics layout the handle was created with%s", "");
    }
  }
  ecap = (ecap + 63) / 64 * 64;
  scap = (scap + 63) / 64 * 64;
  if (scap < 2 * ecap) scap = 2 * ecap;
  rcap = (rcap + 31) / 32 * 32;
  if (mg_raster_smem_bytes(h->cfg.obs_mode, ecap, scap, rcap) > 200 * 1024)
    return fail(MG_E_INVALID, "mg_update_scenes: scene has too many draw edges for the rasteriser's shared memory%s", "");
  h->ecap = ecap;
  h->scap = scap;
  h->rcap = rcap;
  /* environments bound to these entries must have finished their episodes (caller's contract, see header) */
  CUDA_TRY(cudaMemcpyAsync(h->d_scenes + first, host.data(), sizeof(DeviceScene) * (size_t)n, cudaMemcpyHostToDevice,
                           h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return MG_OK;
}

int64_t mg_sizeof_placement(void) { return (int64_t)sizeof(mg_placement_t); }

int mg_set_placement(mg_handle* h, int32_t first, int32_t n, const mg_placement_t* programs) {
  if (!h || !programs) return fail(MG_E_INVALID, "mg_set_placement: null argument%s", "");
  if (!h->cfg.device_sampling) return fail(MG_E_STATE, "mg_set_placement: the handle was not created with device_sampling%s", "");
  if (first < 0 || n <= 0 || first + n > h->cfg.n_scenes) return fail(MG_E_INVALID, "mg_set_placement: range%s", "");
  for (int i = 0; i < n; i++) {
    const mg_placement_t& P = programs[i];
    if (P.n_ents < 0 || P.n_ents > MG_MAX_PLACE_ENTS || P.n_hw < 0 || P.n_hw > MG_MAX_GOALS)
      return fail(MG_E_INVALID, "mg_set_placement: bad entity / size-draw count%s", "");
    for (int k = 0; k < P.n_ents; k++) {
      const mg_place_ent_t& E = P.ents[k];
      bool ok = (E.kind == 0 || E.kind == 1);
      if (E.kind == 1) ok = ok && E.goal >= 0 && E.goal < MG_MAX_GOALS;
      if (E.kind == 0) {
        ok = ok && E.n_bodies >= 1 && E.n_bodies <= MG_MAX_PLACE_BODIES && E.n_groups >= 0 && E.n_groups <= 4;
        for (int j = 0; ok && j < E.n_bodies; j++) ok = E.bodies[j] >= 0 && E.bodies[j] < MG_MAX_BODIES;
        for (int j = 0; ok && j < E.n_groups; j++) ok = E.groups[j] >= 0 && E.groups[j] < MG_MAX_CGROUPS;
      }
      if (!ok || !(E.rand_pos || E.rand_rot)) return fail(MG_E_INVALID, "mg_set_placement: bad entity entry%s", "");
    }
    for (int k = 0; k < P.n_hw; k++)
      if (P.hw[k].goal < 0 || P.hw[k].goal >= MG_MAX_GOALS) return fail(MG_E_INVALID, "mg_set_placement: bad size-draw entry%s", "");
    for (int g = 0; g < MG_MAX_GOALS; g++)
      for (int q = 0; q < 2; q++)
        if (P.goal_prims[g][q] >= MG_MAX_PRIMS) return fail(MG_E_INVALID, "mg_set_placement: bad goal primitive%s", "");
  }
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaMemcpyAsync(h->d_programs + first, programs, sizeof(mg_placement_t) * (size_t)n, cudaMemcpyHostToDevice,
                           h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->programs_set = 1;
  return MG_OK;
}

int mg_get_env_scene(mg_handle* h, int32_t env, mg_scene_t* out) {
  if (!h || !out) return fail(MG_E_INVALID, "mg_get_env_scene: null argument%s", "");
  if (env < 0 || env >= h->cfg.batch) return fail(MG_E_INVALID, "mg_get_env_scene: env out of range%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  int32_t scene = 0;
  CUDA_TRY(cudaMemcpyAsync(&scene, &h->d_states[env].scene, sizeof(scene), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->cfg.device_sampling) scene = h->cfg.n_scenes + env; /* the environment's own sampled layout */
  CUDA_TRY(cudaMemcpy(out, &h->d_scenes[scene].s, sizeof(mg_scene_t), cudaMemcpyDeviceToHost));
  return MG_OK;
}

int mg_get_poses(mg_handle* h, double* out_host) {
  if (!h || !out_host) return fail(MG_E_INVALID, "mg_get_poses: null argument%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  CUDA_TRY(cudaMemcpy2DAsync(out_host, sizeof(double4) * MG_MAX_BODIES, &h->d_states[0].P[0], sizeof(EnvState),
                             sizeof(double4) * MG_MAX_BODIES, (size_t)h->cfg.batch, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return MG_OK;
}

int mg_sampler_failures(mg_handle* h, int64_t* out) {
  if (!h || !out) return fail(MG_E_INVALID, "mg_sampler_failures: null argument%s", "");
  *out = 0;
  if (!h->cfg.device_sampling) return MG_OK;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  unsigned long long v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, h->d_failures, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out = (int64_t)v;
  return MG_OK;
}

int mg_set_draw_range(mg_handle* h, int32_t first, int32_t n) {
  if (!h) return fail(MG_E_INVALID, "mg_set_draw_range: null handle%s", "");
  if (first < 0 || n < 0 || first + n > h->cfg.n_scenes) return fail(MG_E_INVALID, "mg_set_draw_range: range%s", "");
  h->draw_first = first;
  h->draw_count = n;
  return MG_OK;
}

int64_t mg_launch_count(const mg_handle* h) { return h ? h->launches : 0; }

int mg_overflow_count(mg_handle* h, int64_t* out) {
  if (!h || !out) return fail(MG_E_INVALID, "mg_overflow_count: null argument%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  unsigned long long v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, h->d_overflow, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  *out = (int64_t)v;
  return MG_OK;
}

int mg_synchronize(mg_handle* h) {
  if (!h) return fail(MG_E_INVALID, "mg_synchronize: null handle%s", "");
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  { int rc_ = join_sampler(h); if (rc_ != MG_OK) return rc_; }
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return MG_OK;
}

} /* extern "C" */
