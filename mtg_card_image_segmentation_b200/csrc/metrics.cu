// IoU / Dice / pixel-accuracy as ONE integer reduction: the 2x2 confusion counts
//   counts[t*2 + p] = #{pixels : target == t and argmax(logits) == p},  argmax ties -> class 0
// from which every ratio of train/utils.py:94-164 and the confusion matrix of train/evaluate.py:88
// follows exactly.  12 B/pixel with fp16/bf16 logits (2 logits + int64 target), pure HBM bandwidth.
#include <cuda_fp16.h>

#include "ops.h"

namespace mtgseg {
namespace {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256) metric_counts_kernel(const T* __restrict__ logits, const int64_t* __restrict__ targets,
                                                            unsigned long long* __restrict__ counts, long long batch, long long hw) {
  __shared__ unsigned long long sc[4];
  if (threadIdx.x < 4) sc[threadIdx.x] = 0ull;
  __syncthreads();
  unsigned int c[4] = {0u, 0u, 0u, 0u};
  const long long total = batch * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / hw, px = i - n * hw;
    const float z0 = to_f(logits[(n * 2) * hw + px]);
    const float z1 = to_f(logits[(n * 2 + 1) * hw + px]);
    const int t = static_cast<int>(targets[i]) & 1;
    c[t * 2 + (z1 > z0 ? 1 : 0)]++;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned int v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[k], static_cast<unsigned long long>(v));
  }
  __syncthreads();
  if (threadIdx.x < 4 && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}

}  // namespace

int launch_metric_counts(const void* logits, int dtype, const int64_t* targets, unsigned long long* counts4, long long batch,
                         long long hw, cudaStream_t st) {
  MTG_REQUIRE(logits && targets && counts4, MTG_ERR_ARG, "metric_counts: null pointer");
  MTG_REQUIRE(batch >= 0 && hw >= 0, MTG_ERR_ARG, "metric_counts: negative size");
  const long long total = batch * hw;
  if (total == 0) return MTG_OK;
  long long blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  const int g = static_cast<int>(blocks);
  if (dtype == LOGITS_F32) metric_counts_kernel<float><<<g, 256, 0, st>>>(static_cast<const float*>(logits), targets, counts4, batch, hw);
  else if (dtype == LOGITS_BF16) metric_counts_kernel<bf16><<<g, 256, 0, st>>>(static_cast<const bf16*>(logits), targets, counts4, batch, hw);
  else if (dtype == LOGITS_F16) metric_counts_kernel<__half><<<g, 256, 0, st>>>(static_cast<const __half*>(logits), targets, counts4, batch, hw);
  else MTG_REQUIRE(false, MTG_ERR_ARG, "metric_counts: unknown logits dtype %d", dtype);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
