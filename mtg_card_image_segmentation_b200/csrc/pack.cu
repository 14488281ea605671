// Weight packing: reference state_dict tensors (fp32, OIHW) -> the layouts the kernels read.
// Runs once per weight update; bandwidth is irrelevant (16 MB), so these are plain grid-stride kernels.
#include "ops.h"

namespace mtgseg {
namespace {

__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}
__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = in[i];
}
// in [O][I][T] -> out [O][T][I]
__global__ void oihw_to_otapi_kernel(const float* __restrict__ in, bf16* __restrict__ out, int O, int I, int T) {
  const size_t n = static_cast<size_t>(O) * I * T;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % I);
    const size_t r = i / I;
    const int t = static_cast<int>(r % T);
    const size_t o = r / T;
    out[i] = __float2bfloat16(in[(o * I + ci) * T + t]);
  }
}
// in [C][T] -> out [T][C] and (optional) the tap-reversed copy used by the stride-1 depthwise dgrad
__global__ void pack_dw_kernel(const float* __restrict__ in, bf16* __restrict__ out, bf16* __restrict__ out_flip, int C, int T) {
  const int n = C * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i % C, t = i / C;
    const bf16 v = __float2bfloat16(in[c * T + t]);
    out[i] = v;
    if (out_flip) out_flip[(T - 1 - t) * C + c] = v;
  }
}
// in [16][27] -> out [27][16]
__global__ void pack_stem_kernel(const float* __restrict__ in, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 27 * 16) out[i] = in[(i % 16) * 27 + i / 16];
}
__global__ void fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ m,
                               const float* __restrict__ v, float eps, float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float s = g[i] / sqrtf(v[i] + eps);
    scale[i] = s;
    shift[i] = b[i] - m[i] * s;
  }
}

// in [N][K] -> out [K][N]
__global__ void transpose_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int N, int K) {
  const size_t n = static_cast<size_t>(N) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int nn = static_cast<int>(i % N), kk = static_cast<int>(i / N);
    out[i] = __float2bfloat16(in[static_cast<size_t>(nn) * K + kk]);
  }
}
// in [O][I][9] -> out [I][9][O] with the taps flipped (dgrad of a stride-1 pad-1 3x3 conv is a 3x3 conv with w^T, rot180)
__global__ void dgrad3x3_kernel(const float* __restrict__ in, bf16* __restrict__ out, int O, int I) {
  const size_t n = static_cast<size_t>(O) * I * 9;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % O);
    const size_t r = i / O;
    const int t = static_cast<int>(r % 9);
    const size_t ci = r / 9;
    out[i] = __float2bfloat16(in[(static_cast<size_t>(o) * I + ci) * 9 + (8 - t)]);
  }
}

inline int blocks_for(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 1024) b = 1024;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

int launch_cast_bf16(const float* in, bf16* out, size_t n, cudaStream_t st) {
  cast_bf16_kernel<<<blocks_for(n), 256, 0, st>>>(in, out, n);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_copy_f32(const float* in, float* out, size_t n, cudaStream_t st) {
  copy_f32_kernel<<<blocks_for(n), 256, 0, st>>>(in, out, n);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_oihw_to_otapi(const float* in, bf16* out, int O, int I, int taps, cudaStream_t st) {
  oihw_to_otapi_kernel<<<blocks_for(static_cast<size_t>(O) * I * taps), 256, 0, st>>>(in, out, O, I, taps);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_dw(const float* in, bf16* out, bf16* out_flip, int C, int taps, cudaStream_t st) {
  pack_dw_kernel<<<blocks_for(static_cast<size_t>(C) * taps), 256, 0, st>>>(in, out, out_flip, C, taps);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_transpose_bf16(const float* in, bf16* out, int N, int K, cudaStream_t st) {
  transpose_bf16_kernel<<<blocks_for(static_cast<size_t>(N) * K), 256, 0, st>>>(in, out, N, K);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_dgrad3x3(const float* in, bf16* out, int O, int I, cudaStream_t st) {
  dgrad3x3_kernel<<<blocks_for(static_cast<size_t>(O) * I * 9), 256, 0, st>>>(in, out, O, I);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_stem(const float* in, float* out, cudaStream_t st) {
  pack_stem_kernel<<<2, 256, 0, st>>>(in, out);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                   float* shift, int C, cudaStream_t st) {
  fold_bn_kernel<<<ceil_div(C, 256), 256, 0, st>>>(gamma, beta, mean, var, eps, scale, shift, C);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
