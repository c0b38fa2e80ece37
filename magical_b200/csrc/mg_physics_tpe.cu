/*
 * mg_physics_tpe.cu — K1 launcher, thread-per-environment form (see mg_physics_tpe.h for the design and
 * the reference calls it replaces: entities.py:439-479, base_env.py:236-243, pymunk Space.step).
 *
 * One warp-sized block = 32 environments; each block owns `words * 32` 8-byte words of shared memory laid
 * out [word][lane].  Lanes never communicate, so there is no barrier in the kernel.  Blocks per SM are
 * bounded by shared memory (about 1.7-2.2 KB per environment), which is what sizes `kcon` (the number of
 * solver contacts held on chip; rarer, larger contact sets continue in a per-environment spill area in HBM).
 */
#define MG_NP_FORCE_INLINE
#include "mg_physics_tpe.h"

#define TPE_THREADS 32

__global__ void __launch_bounds__(TPE_THREADS)
k_physics_tpe(EnvState* __restrict__ states, const DeviceScene* __restrict__ scenes, const int32_t* __restrict__ actions,
              int env0, int count, TpeLayout L, double* __restrict__ spill, uint32_t* __restrict__ scratch) {
  extern __shared__ __align__(16) double tpe_words[];
  int env = env0 + blockIdx.x * TPE_THREADS + threadIdx.x;
  const int batch = env0 + count; /* this launch covers environments [env0, env0 + count) */
  /* lanes beyond the batch keep their warp complete for the cooperative narrowphase (they do no work of
   * their own and store nothing) */
  const bool live = env < batch;
  if (!live) env = batch - 1;
  EnvState* G = states + env;
  const DeviceScene* ds = scenes + G->scene;
  Tpe<TPE_THREADS> T;
  T.wd = tpe_words + threadIdx.x;
  T.wf = reinterpret_cast<float*>(tpe_words) + threadIdx.x;
  T.wh = reinterpret_cast<uint16_t*>(tpe_words) + threadIdx.x;
  T.L = L;
  /* spill record of this launch slot's block, interleaved like the private words */
  T.spill = spill ? spill + (size_t)(env0 / TPE_THREADS + blockIdx.x) * (size_t)((TPE_MAX_CONTACTS - L.kcon) * TPE_CON_WORDS) *
                                TPE_THREADS + threadIdx.x
                  : nullptr;
  T.slotmap = 0;
  T.static_slot = 0;
  /* scratch records are indexed by launch slot (not by the clamped env), so every lane has its own */
  T.bind_scratch(scratch ? scratch + (size_t)(env0 + blockIdx.x * TPE_THREADS + threadIdx.x) * (size_t)L.scratch_u32 : nullptr);
  tpe_env_step<TPE_THREADS>(T, G, ds, actions[env], live);
}

size_t mg_tpe_smem_bytes(const TpeLayout* L) { return (size_t)L->words * sizeof(double) * TPE_THREADS; }
size_t mg_tpe_spill_doubles_per_env(const TpeLayout* L) { return (size_t)(TPE_MAX_CONTACTS - L->kcon) * TPE_CON_WORDS; }

cudaError_t mg_launch_physics_tpe(EnvState* states, const DeviceScene* scenes, const int32_t* actions, int env0,
                                  int count, const TpeLayout* L, double* spill, uint32_t* scratch,
                                  cudaStream_t stream) {
  const size_t smem = mg_tpe_smem_bytes(L);
  /* the opt-in is a per-DEVICE function attribute: one process may hold handles on several GPUs */
  static size_t configured_on[64] = {0};
  int dev = 0;
  cudaError_t de = cudaGetDevice(&dev);
  if (de != cudaSuccess) return de;
  size_t& configured = configured_on[dev & 63];
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(k_physics_tpe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_physics_tpe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  k_physics_tpe<<<(count + TPE_THREADS - 1) / TPE_THREADS, TPE_THREADS, smem, stream>>>(states, scenes, actions, env0,
                                                                                      count, *L, spill, scratch);
  return cudaGetLastError();
}
