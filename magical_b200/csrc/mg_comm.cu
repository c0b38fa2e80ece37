/*
 * mg_comm.cu — multi-GPU transport of the observation gather over NVLink peer memory.
 *
 * One process per GPU.  Every rank owns a REGION of device memory that its peers map through CUDA IPC:
 *     [ barrier flags | per buffer b: packed scalars (reward, score, done) | newest frames [views][n][R][R][3] ]
 * k_raster writes the rank's newest frames and k_finish its scalars straight into the region (they are the
 * "send buffers", double-buffered).  After a cross-rank barrier (k_xbarrier: release-store of the epoch into
 * every peer's flag slot, acquire-spin on the own slots) each rank runs
 *     k_peer_gather       the peers' packed scalars  -> a local [world, bytes] buffer      (12 B per env)
 *     k_stack_push_p2p    FlattenFrameStack of every REMOTE environment, the frame being read directly from
 *                         the owner's region over NVLink (384 contiguous bytes per warp) while the 48-byte
 *                         stack groups stream through local HBM
 * i.e. the all-gather and the stack rebuild are ONE kernel: no receive buffer, no NCCL kernel competing with
 * k_physics_tpe for shared memory, the NVLink transfer overlapped with the HBM work tile by tile.
 * Reference semantics: FlattenFrameStack.observation / reset, benchmarks/__init__.py:118-136.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/magical_b200.h"

#define MG_COMM_MAX_RANKS 16
#define MG_COMM_FLAG_BYTES 256

struct PeerPtrs {
  uint8_t* p[MG_COMM_MAX_RANKS];
};

struct mg_comm {
  int rank, world, device, n_buf, connected;
  int64_t scalar_bytes, frame_bytes; /* per buffer */
  int64_t buf_stride, region_bytes;
  uint8_t* local;
  PeerPtrs peers;
  uint32_t epoch;
  uint32_t* d_err;
};

void mg_set_error_(const char* msg); /* mg_api.cu: the message mg_last_error() returns */
static int cfail(int code, const char* what, const char* detail) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
  mg_set_error_(buf);
  return code;
}
#define COMM_TRY(expr)                                                         \
  do {                                                                         \
    cudaError_t e_ = (expr);                                                   \
    if (e_ != cudaSuccess) return cfail(MG_E_CUDA, #expr, cudaGetErrorString(e_)); \
  } while (0)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

/* Cross-rank barrier: everything this rank's stream executed before is visible to the peers that pass it
 * (kernel boundaries write back to L2, the point of coherence peers read through), and everything the peers
 * executed before THEIR barrier is visible to kernels launched after this one (fresh L1). */
__global__ void k_xbarrier(const __grid_constant__ PeerPtrs peers, int rank, int world, uint32_t epoch, unsigned long long timeout_ns,
                           uint32_t* err) {
  const int q = threadIdx.x;
  if (q >= world || q == rank) return;
  __threadfence_system();
  st_release_sys(reinterpret_cast<uint32_t*>(peers.p[q]) + rank, epoch);
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(peers.p[rank]) + q;
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(err, 1u); break; } /* a peer died: do not hang the GPU */
    __nanosleep(256);
  }
}

/* dst[r][0..bytes) <- peer r's region at `offset` (16-byte units), all ranks incl. the own one */
__global__ void __launch_bounds__(256)
k_peer_gather(const __grid_constant__ PeerPtrs peers, int64_t offset, uint4* __restrict__ dst, int64_t n16) {
  const int r = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(peers.p[r] + offset);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
    dst[r * n16 + i] = __ldcs(src + i);
}

/* FlattenFrameStack of the remote environments [env_first, env_first + count), frames read from the owners'
 * regions.  One thread = 4 pixels = 48 B of stack + 12 B of frame; a warp reads 384 contiguous frame bytes
 * over NVLink (three non-coherent 32-bit loads per lane: the second and third hit the 128-byte lines the
 * first one brought into L1; the buffers are read-only while the kernel runs, L1 is fresh at every launch).
 * No shared memory and ~32 registers, so its blocks fit into the registers k_physics_tpe leaves free (that
 * kernel is bound by shared memory at 4 warps per SM) and the NVLink + HBM streaming hides under the next
 * step's physics.  The remote loads are issued first and consumed last: the NVLink round trip (~3 us) overlaps
 * the HBM read of the stack. */
__global__ void __launch_bounds__(256)
k_stack_push_p2p(uint8_t* __restrict__ stacks, const __grid_constant__ PeerPtrs peers, int64_t frames_offset /* of this buffer + view */,
                 const uint8_t* __restrict__ fresh, long long env_first, long long n_groups_total, int groups_per_env,
                 int shard, long long env_modulo) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups_total) return;
  /* env_modulo > 0: the range wraps around, so that rank r can walk its peers in the order r+1, r+2, ...:
   * with every rank starting at a different owner no GPU's NVLink egress serves all readers at once */
  long long env = env_first + g / groups_per_env;
  if (env_modulo > 0 && env >= env_modulo) env -= env_modulo;
  const int gi = (int)(g % groups_per_env);
  const uint32_t* np_ = reinterpret_cast<const uint32_t*>(peers.p[env / shard] + frames_offset +
                                                          ((env % shard) * (long long)groups_per_env + gi) * 12);
  const uint32_t n0 = __ldg(np_), n1 = __ldg(np_ + 1), n2 = __ldg(np_ + 2);
  uint4* sp = reinterpret_cast<uint4*>(stacks + (env * groups_per_env + gi) * 48);
  const uint4 a = __ldcs(sp), b = __ldcs(sp + 1), c = __ldcs(sp + 2); /* evict-first: stream past the L2-resident physics state */
  const bool f = fresh != nullptr && fresh[env] != 0;
  uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  const uint32_t col[4] = {n0 & 0xFFFFFFu, (n0 >> 24) | ((n1 & 0xFFFFu) << 8), (n1 >> 16) | ((n2 & 0xFFu) << 16), n2 >> 8};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t n = col[i];
    if (f) {
      w[3 * i] = n | (n << 24);
      w[3 * i + 1] = (n >> 8) | (n << 16);
      w[3 * i + 2] = (n >> 16) | (n << 8);
    } else {
      const uint32_t w0 = w[3 * i], w1 = w[3 * i + 1], w2 = w[3 * i + 2];
      w[3 * i] = (w0 >> 24) | (w1 << 8);
      w[3 * i + 1] = (w1 >> 24) | (w2 << 8);
      w[3 * i + 2] = (w2 >> 24) | (n << 8);
    }
  }
  __stcs(sp, make_uint4(w[0], w[1], w[2], w[3]));
  __stcs(sp + 1, make_uint4(w[4], w[5], w[6], w[7]));
  __stcs(sp + 2, make_uint4(w[8], w[9], w[10], w[11]));
}

static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

extern "C" {

int mg_comm_create(int32_t rank, int32_t world, int32_t n_buffers, int64_t scalar_bytes, int64_t frame_bytes,
                   mg_comm** out) {
  if (!out || rank < 0 || world < 1 || rank >= world || world > MG_COMM_MAX_RANKS || n_buffers < 1 || n_buffers > 4 ||
      scalar_bytes < 0 || frame_bytes < 0 || (scalar_bytes & 15) != 0 || (frame_bytes & 15) != 0)
    return cfail(MG_E_INVALID, "mg_comm_create: bad argument (sizes must be multiples of 16, world <= 16)", nullptr);
  mg_comm* c = new (std::nothrow) mg_comm();
  if (!c) return cfail(MG_E_NOMEM, "mg_comm_create: out of host memory", nullptr);
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->n_buf = n_buffers;
  c->scalar_bytes = scalar_bytes; c->frame_bytes = frame_bytes;
  c->buf_stride = align256(scalar_bytes) + align256(frame_bytes);
  c->region_bytes = MG_COMM_FLAG_BYTES + c->buf_stride * n_buffers;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(&c->local, (size_t)c->region_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, (size_t)c->region_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_err, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(c->d_err, 0, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(c->local); cudaFree(c->d_err);
    delete c;
    return cfail(MG_E_CUDA, "mg_comm_create", cudaGetErrorString(e));
  }
  c->peers.p[rank] = c->local;
  *out = c;
  return MG_OK;
}

int mg_comm_export(mg_comm* c, void* handle_out) {
  if (!c || !handle_out) return cfail(MG_E_INVALID, "mg_comm_export: null argument", nullptr);
  static_assert(sizeof(cudaIpcMemHandle_t) == MG_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  COMM_TRY(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle_out, &h, sizeof(h));
  return MG_OK;
}

int mg_comm_connect(mg_comm* c, const void* all_handles) {
  if (!c || !all_handles) return cfail(MG_E_INVALID, "mg_comm_connect: null argument", nullptr);
  if (c->connected) return cfail(MG_E_STATE, "mg_comm_connect: already connected", nullptr);
  COMM_TRY(cudaSetDevice(c->device));
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)all_handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    COMM_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peers.p[r] = (uint8_t*)p;
  }
  c->connected = 1;
  return MG_OK;
}

void* mg_comm_scalar_ptr(mg_comm* c, int32_t buffer) {
  if (!c || buffer < 0 || buffer >= c->n_buf) return nullptr;
  return c->local + MG_COMM_FLAG_BYTES + c->buf_stride * buffer;
}
void* mg_comm_frame_ptr(mg_comm* c, int32_t buffer) {
  if (!c || buffer < 0 || buffer >= c->n_buf) return nullptr;
  return c->local + MG_COMM_FLAG_BYTES + c->buf_stride * buffer + align256(c->scalar_bytes);
}

int mg_comm_barrier(mg_comm* c, void* cuda_stream) {
  if (!c || !c->connected) return cfail(MG_E_STATE, "mg_comm_barrier: not connected", nullptr);
  c->epoch++;
  k_xbarrier<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(c->peers, c->rank, c->world, c->epoch,
                                                     20ull * 1000000000ull, c->d_err);
  COMM_TRY(cudaGetLastError());
  return MG_OK;
}

int mg_comm_gather_scalars(mg_comm* c, int32_t buffer, void* dst_dev, void* cuda_stream) {
  if (!c || !c->connected || !dst_dev || buffer < 0 || buffer >= c->n_buf)
    return cfail(MG_E_INVALID, "mg_comm_gather_scalars: bad argument", nullptr);
  if (((uintptr_t)dst_dev & 15) != 0) return cfail(MG_E_INVALID, "mg_comm_gather_scalars: dst must be 16-byte aligned", nullptr);
  const int64_t n16 = c->scalar_bytes / 16;
  if (n16 == 0) return MG_OK;
  int bx = (int)((n16 + 255) / 256);
  if (bx > 64) bx = 64;
  k_peer_gather<<<dim3(bx, c->world), 256, 0, (cudaStream_t)cuda_stream>>>(
      c->peers, MG_COMM_FLAG_BYTES + c->buf_stride * buffer, (uint4*)dst_dev, n16);
  COMM_TRY(cudaGetLastError());
  return MG_OK;
}

int mg_comm_stack_push(mg_comm* c, int32_t buffer, int64_t view_offset, void* stacks_dev, const uint8_t* fresh_dev,
                       int64_t env_first, int64_t env_count, int64_t env_modulo, int32_t shard, int32_t res,
                       void* cuda_stream) {
  if (!c || !c->connected || !stacks_dev || buffer < 0 || buffer >= c->n_buf)
    return cfail(MG_E_INVALID, "mg_comm_stack_push: bad argument", nullptr);
  if (env_count <= 0) return MG_OK;
  const int gpe = res * res / 4;
  if (res <= 0 || gpe % 32 != 0 || shard <= 0 || env_first < 0 || (view_offset & 15) != 0 ||
      (env_modulo == 0 && (env_first + env_count + shard - 1) / shard > c->world) ||
      (env_modulo != 0 && (env_modulo != (int64_t)shard * c->world || env_first >= env_modulo || env_count > env_modulo)) ||
      ((uintptr_t)stacks_dev & 15) != 0)
    return cfail(MG_E_INVALID, "mg_comm_stack_push: bad range / alignment", nullptr);
  const long long total = env_count * gpe;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7FFFFFFFLL) return cfail(MG_E_INVALID, "mg_comm_stack_push: range too large", nullptr);
  k_stack_push_p2p<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(
      (uint8_t*)stacks_dev, c->peers,
      MG_COMM_FLAG_BYTES + c->buf_stride * buffer + align256(c->scalar_bytes) + view_offset, fresh_dev, env_first, total,
      gpe, shard, env_modulo);
  COMM_TRY(cudaGetLastError());
  return MG_OK;
}

int mg_comm_error(mg_comm* c, int32_t* out) {
  if (!c || !out) return cfail(MG_E_INVALID, "mg_comm_error: null argument", nullptr);
  uint32_t v = 0;
  COMM_TRY(cudaMemcpy(&v, c->d_err, sizeof(v), cudaMemcpyDeviceToHost));
  *out = (int32_t)v;
  return MG_OK;
}

int mg_comm_destroy(mg_comm* c) {
  if (!c) return MG_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; r++)
    if (r != c->rank && c->peers.p[r]) cudaIpcCloseMemHandle(c->peers.p[r]);
  cudaFree(c->local);
  cudaFree(c->d_err);
  delete c;
  return MG_OK;
}

} /* extern "C" */
