/*
 * mg_stack.cu — K4 `k_stack_push`: FlattenFrameStack for environments that were rendered on ANOTHER GPU.
 *
 * Multi-GPU layout: every rank renders its own shard straight into its slice of the global
 * observation tensor (k_raster) and additionally writes each environment's NEWEST frame (27 648 B) into a
 * send buffer; the send buffers are all-gathered (NCCL over NVLink), and this kernel folds the received
 * frames into the 4-frame stacks of the remote shards:
 *     stack <- stack[1:] + [frame]            (benchmarks/__init__.py:124-128, deque(maxlen=4).append)
 *     stack <- [frame] * 4  where `fresh`     (benchmarks/__init__.py:130-136, reset fills the deque)
 * Pure streaming work: per environment read 110 592 B (the 48-byte groups hold the dropped frame's bytes
 * in the same sectors as the surviving ones) + 27 648 B, write 110 592 B.  One thread = 4 pixels = 48 B of
 * stack (3 x 128-bit) + 12 B of frame.
 */
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256)
k_stack_push(uint8_t* __restrict__ stacks, const uint8_t* __restrict__ newest, const uint8_t* __restrict__ fresh,
             long long env_first, long long n_groups_total, int groups_per_env, int shard,
             long long rank_stride /* bytes between consecutive ranks' blocks in `newest` */) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups_total) return;
  const long long env = env_first + g / groups_per_env;
  const int gi = (int)(g % groups_per_env);
  const uint32_t* np_ = reinterpret_cast<const uint32_t*>(newest + (env / shard) * rank_stride +
                                                          ((env % shard) * (long long)groups_per_env + gi) * 12);
  const uint32_t n0 = __ldcs(np_), n1 = __ldcs(np_ + 1), n2 = __ldcs(np_ + 2);
  /* the 4 pixels' colours */
  const uint32_t c0 = n0 & 0xFFFFFFu, c1 = (n0 >> 24) | ((n1 & 0xFFFFu) << 8), c2 = (n1 >> 16) | ((n2 & 0xFFu) << 16),
                 c3 = n2 >> 8;
  uint4* sp = reinterpret_cast<uint4*>(stacks + (env * groups_per_env + gi) * 48);
  uint32_t w[12];
  const bool f = fresh != nullptr && fresh[env] != 0;
  if (!f) {
    const uint4 a = __ldcs(sp), b = __ldcs(sp + 1), c = __ldcs(sp + 2); /* evict-first: stream past the L2-resident physics state */
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
  }
  const uint32_t col[4] = {c0, c1, c2, c3};
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

cudaError_t mg_launch_stack_push(uint8_t* stacks, const uint8_t* newest, const uint8_t* fresh, long long env_first,
                                 long long env_count, int shard, long long rank_stride, int res, cudaStream_t stream) {
  if (env_count <= 0) return cudaSuccess;
  const int gpe = res * res / 4;
  const long long total = env_count * gpe;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7FFFFFFFLL) return cudaErrorInvalidValue;
  k_stack_push<<<(unsigned)blocks, 256, 0, stream>>>(stacks, newest, fresh, env_first, total, gpe, shard, rank_stride);
  return cudaGetLastError();
}
