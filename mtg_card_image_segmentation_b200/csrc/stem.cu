// Stem: dense 3x3 stride-2 convolution 3 -> 16 channels + folded BatchNorm + Hardswish, reading the
// reference's fp32 NCHW batch (train/dataset.py:84-88 contract) and writing NHWC bf16.
// K = 27 is far too small for the tensor cores to matter; the layer is bound by reading the image
// (12 B/pixel) and writing 32 B per output pixel, so it is a direct convolution: one thread per pair of output
// pixels, all 16 output channels of both in registers, weights broadcast from shared memory.
// Replaces features[0] of tv:models/mobilenetv3.py:160-170.
#include "ops.h"

namespace mtgseg {
namespace {

// kU8: the batch is raw uint8 HWC camera/dataset pixels; (v/255 - mean)/std (train/dataset.py:182-185) is applied on load,
// so the host never materialises the 4x larger normalised fp32 NCHW tensor.  Zero padding lives in the normalised domain.
template <bool kU8, bool kVec>
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ x, const uint8_t* __restrict__ xu8, float3 nmul,
                                                   float3 nadd, const float* __restrict__ w,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   bf16* __restrict__ out, int B, int H, int W, int Ho, int Wo, int act) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 sw[27 * 4];
  __shared__ float ssc[16], ssh[16];
  for (int i = threadIdx.x; i < 27 * 4; i += blockDim.x) sw[i] = reinterpret_cast<const float4*>(w)[i];
  if (threadIdx.x < 16) { ssc[threadIdx.x] = scale[threadIdx.x]; ssh[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  // one thread = two horizontally adjacent output pixels: the 3x3 stride-2 windows share a column (45 instead of 54
  // input values) and every weight quad read from shared memory feeds both pixels; the FMAs are issued as packed
  // fp32x2 over output-channel pairs (FFMA2), the input value broadcast into both halves.
  const int Wp = (Wo + 1) >> 1;  // pixel pairs per output row
  const long long total = static_cast<long long>(B) * Ho * Wp;
  // (whole warps stay in the loop: the wide-load path shuffles between lanes; lanes past the end recompute the last pair
  // and skip the store)
  const long long total_w = (total + 31) / 32 * 32;
  for (long long idx0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx0 < total_w;
       idx0 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool live = idx0 < total;
    const long long idx = live ? idx0 : total - 1;
    const int oxp = static_cast<int>(idx % Wp);
    const long long t = idx / Wp;
    const int oy = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    const int ox = oxp * 2;
    const bool second = ox + 1 < Wo;
    // gather the 45 taps first (no branches between the loads: all are in flight together)
    float xin[3][3][5];
    if (kVec) {
      // W % 4 == 0: a thread's window is columns 4*oxp - 1 .. 4*oxp + 3.  The four aligned columns come as ONE 16-byte load
      // (fp32) / three 4-byte words (uint8 RGB) per channel-row, the column to the left from the neighbouring lane (it holds
      // the previous pixel pair of the same row whenever oxp > 0; lane 0 loads it itself).  9 wide loads instead of 45 scalar
      // ones: the scalar version was bound by L1 wavefronts (each warp-wide scalar load touched 16 sectors for 128 useful bytes).
      const int lane = threadIdx.x & 31;
      const bool need_left = oxp > 0;
      if (kU8) {
        const uint8_t* xn = xu8 + static_cast<size_t>(n) * H * W * 3;
        const float mul[3] = {nmul.x, nmul.y, nmul.z}, add[3] = {nadd.x, nadd.y, nadd.z};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = oy * 2 - 1 + ky;
          const bool ok = iy >= 0 && iy < H;
          const uint8_t* row = xn + (static_cast<size_t>(ok ? iy : 0) * W + 4 * oxp) * 3;
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
          uint32_t w0 = __ldg(rw), w1 = __ldg(rw + 1), w2 = __ldg(rw + 2);
          uint32_t left = __shfl_up_sync(0xffffffffu, w2, 1) >> 8;  // bytes 9..11 of the neighbour = its last pixel
          if (lane == 0 && need_left) left = static_cast<uint32_t>(__ldg(row - 3)) | (static_cast<uint32_t>(__ldg(row - 2)) << 8) | (static_cast<uint32_t>(__ldg(row - 1)) << 16);
          const uint32_t by[5][3] = {{left & 255u, (left >> 8) & 255u, (left >> 16) & 255u},
                                     {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u},
                                     {w0 >> 24, w1 & 255u, (w1 >> 8) & 255u},
                                     {(w1 >> 16) & 255u, w1 >> 24, w2 & 255u},
                                     {(w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24}};
#pragma unroll
          for (int c = 0; c < 5; ++c)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
              xin[ci][ky][c] = (ok && (c > 0 || need_left)) ? fmaf(static_cast<float>(by[c][ci]), mul[ci], add[ci]) : 0.f;
        }
      } else {
        const float* xn = x + static_cast<size_t>(n) * 3 * H * W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - 1 + ky;
            const bool ok = iy >= 0 && iy < H;
            const float* row = xn + (static_cast<size_t>(ci) * H + (ok ? iy : 0)) * W + 4 * oxp;
            const float4 v = __ldg(reinterpret_cast<const float4*>(row));
            float left = __shfl_up_sync(0xffffffffu, v.w, 1);
            if (lane == 0 && need_left) left = __ldg(row - 1);
            xin[ci][ky][0] = (ok && need_left) ? left : 0.f;
            xin[ci][ky][1] = ok ? v.x : 0.f; xin[ci][ky][2] = ok ? v.y : 0.f;
            xin[ci][ky][3] = ok ? v.z : 0.f; xin[ci][ky][4] = ok ? v.w : 0.f;
          }
      }
    } else if (kU8) {
      const uint8_t* xn = xu8 + static_cast<size_t>(n) * H * W * 3;
      const float mul[3] = {nmul.x, nmul.y, nmul.z}, add[3] = {nadd.x, nadd.y, nadd.z};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + c;
          const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
          const uint8_t* px = xn + (static_cast<size_t>(ok ? iy : 0) * W + (ok ? ix : 0)) * 3;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) xin[ci][ky][c] = ok ? fmaf(static_cast<float>(__ldg(px + ci)), mul[ci], add[ci]) : 0.f;
        }
    } else {
      const float* xn = x + static_cast<size_t>(n) * 3 * H * W;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + c;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
            xin[ci][ky][c] = ok ? __ldg(xn + (static_cast<size_t>(ci) * H + iy) * W + ix) : 0.f;
          }
    }
    uint64_t acc[2][8];  // [pixel][channel pair] packed fp32x2
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = 0ull;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wp = &sw[(ci * 9 + ky * 3 + kx) * 4];
          uint64_t xx[2];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const float xv = xin[ci][ky][kx + 2 * p];
            asm("mov.b64 %0, {%1,%1};" : "=l"(xx[p]) : "f"(xv));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = wp[q];
            uint64_t w01, w23;
            asm("mov.b64 %0, {%1,%2};" : "=l"(w01) : "f"(wv.x), "f"(wv.y));
            asm("mov.b64 %0, {%1,%2};" : "=l"(w23) : "f"(wv.z), "f"(wv.w));
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[p][2 * q]) : "l"(xx[p]), "l"(w01));
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[p][2 * q + 1]) : "l"(xx[p]), "l"(w23));
            }
          }
        }
    // the activation is uniform over the launch: dispatch once per thread, not once per element
    auto emit = [&](auto actf) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        if (!live || (p == 1 && !second)) break;
        float o0[8], o1[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a0, a1, b0, b1;
          asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc[p][j]));
          asm("mov.b64 {%0,%1}, %2;" : "=f"(b0), "=f"(b1) : "l"(acc[p][4 + j]));
          o0[2 * j] = actf(fmaf(a0, ssc[2 * j], ssh[2 * j]));
          o0[2 * j + 1] = actf(fmaf(a1, ssc[2 * j + 1], ssh[2 * j + 1]));
          o1[2 * j] = actf(fmaf(b0, ssc[8 + 2 * j], ssh[8 + 2 * j]));
          o1[2 * j + 1] = actf(fmaf(b1, ssc[8 + 2 * j + 1], ssh[8 + 2 * j + 1]));
        }
        uint4* op = reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox + p) * 16);
        op[0] = pack8(o0);
        op[1] = pack8(o1);
      }
    };
    if (act == ACT_HSWISH) emit([](float z) { return z * __saturatef(fmaf(z, 1.f / 6.f, 0.5f)); });
    else emit([&](float z) { return apply_act(z, act); });
  }
}

}  // namespace

int launch_stem(const StemArgs& a, cudaStream_t st) {
  MTG_REQUIRE((a.x || a.x_u8) && a.w && a.scale && a.shift && a.out, MTG_ERR_ARG, "stem: null pointer");
  const int Ho = (a.H + 2 - 3) / 2 + 1, Wo = (a.W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(a.B) * Ho * ((Wo + 1) / 2);  // threads: one per pair of output pixels
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  // v/255 normalised: (v/255 - mean)/std = v * 1/(255 std) - mean/std
  const float3 nmul = make_float3(1.f / (255.f * a.std[0]), 1.f / (255.f * a.std[1]), 1.f / (255.f * a.std[2]));
  const float3 nadd = make_float3(-a.mean[0] / a.std[0], -a.mean[1] / a.std[1], -a.mean[2] / a.std[2]);
  // wide loads need 4-column alignment of every row (W % 4 == 0 and an aligned base pointer)
  const bool vec = a.W % 4 == 0 && (a.x_u8 ? (reinterpret_cast<uintptr_t>(a.x_u8) & 3) == 0 : (reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  const dim3 grid(static_cast<int>(blocks)), block(256);
  if (a.x_u8) {
    if (vec) MTG_CUDA(launch_pdl(stem_kernel<true, true>, grid, block, 0, st, nullptr, a.x_u8, nmul, nadd, a.w, a.scale, a.shift, a.out, a.B, a.H, a.W, Ho, Wo, a.act));
    else MTG_CUDA(launch_pdl(stem_kernel<true, false>, grid, block, 0, st, nullptr, a.x_u8, nmul, nadd, a.w, a.scale, a.shift, a.out, a.B, a.H, a.W, Ho, Wo, a.act));
  } else {
    if (vec) MTG_CUDA(launch_pdl(stem_kernel<false, true>, grid, block, 0, st, a.x, nullptr, nmul, nadd, a.w, a.scale, a.shift, a.out, a.B, a.H, a.W, Ho, Wo, a.act));
    else MTG_CUDA(launch_pdl(stem_kernel<false, false>, grid, block, 0, st, a.x, nullptr, nmul, nadd, a.w, a.scale, a.shift, a.out, a.B, a.H, a.W, Ho, Wo, a.act));
  }
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
