// Depthwise k x k convolution (k = 3 or 5; stride 1/2; dilation 1/2) on NHWC bf16 with the eval-mode
// BatchNorm scale/shift and activation fused, plus optional per-chunk channel sums of the output (the
// squeeze-excite global-average-pool, so the SE block never re-reads the tensor).
//
// Bandwidth-bound: every thread owns one 8-channel vector (one 128-bit load/store per tap / output) and
// walks output pixels; consecutive threads own consecutive vectors, so each warp touches contiguous
// 512-byte spans.  The block size is a multiple of the vectors-per-pixel count so a thread's channels
// (hence its weights and BN constants) never change.
//
// Replaces the depthwise Conv2dNormActivation of tv:models/mobilenetv3.py:83-95 (+ the AdaptiveAvgPool2d
// of tv:ops/misc.py:252-253).
#include "ops.h"

namespace mtgseg {
namespace {

struct DwP {
  const bf16* in; const bf16* w; bf16* out;
  const float* scale; const float* shift;
  float* gap;
  int act, H, W, C, Ho, Wo, stride, dil, pad, CV, PL, pix_per_chunk, chunks;
};

template <int KS>
__global__ void __launch_bounds__(256) dwconv_kernel(const DwP p) {
  __shared__ float red[256 * 8];
  const int tid = threadIdx.x;
  const int nthreads = p.CV * p.PL;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const bool active = tid < nthreads;
  const int v = active ? tid % p.CV : 0;
  const int pl = active ? tid / p.CV : 0;
  const int c0 = v * 8;
  float acc_gap[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_gap[j] = 0.f;

  if (active) {
    float sc[8], sh[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.scale + c0)), b = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + 4));
      const float4 c = __ldg(reinterpret_cast<const float4*>(p.shift + c0)), d = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
      sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
      sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
    }
    const int npix = p.Ho * p.Wo;
    const int p_begin = chunk * p.pix_per_chunk;
    const int p_end = min(npix, p_begin + p.pix_per_chunk);
    const bf16* in_n = p.in + static_cast<size_t>(n) * p.H * p.W * p.C + c0;
    bf16* out_n = p.out + static_cast<size_t>(n) * npix * p.C + c0;
    for (int pix = p_begin + pl; pix < p_end; pix += p.PL) {
      const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
      const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        const int iy = iy0 + ky * p.dil;
        if (iy < 0 || iy >= p.H) continue;
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int ix = ix0 + kx * p.dil;
          if (ix < 0 || ix >= p.W) continue;
          float xf[8], wf[8];
          unpack8(ldg16(in_n + (static_cast<size_t>(iy) * p.W + ix) * p.C), xf);
          unpack8(ldg16(p.w + (ky * KS + kx) * p.C + c0), wf);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(xf[j], wf[j], acc[j]);
        }
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = apply_act(fmaf(acc[j], sc[j], sh[j]), p.act);
      const uint4 packed = pack8(o);
      *reinterpret_cast<uint4*>(out_n + static_cast<size_t>(pix) * p.C) = packed;
      if (p.gap) {  // pool what the next layer will actually read (the bf16-rounded values)
        float rf[8];
        unpack8(packed, rf);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_gap[j] += rf[j];
      }
    }
  }
  if (p.gap) {
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(pl * p.CV + v) * 8 + j] = acc_gap[j];
    }
    __syncthreads();
    for (int c = tid; c < p.C; c += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < p.PL; ++l) s += red[l * p.C + c];  // fixed order -> deterministic
      p.gap[(static_cast<size_t>(n) * p.chunks + chunk) * p.C + c] = s;
    }
  }
}

}  // namespace

int dwconv_chunks(int Ho, int Wo, int C, bool need_gap) {
  const int CV = C / 8;
  const int PL = 256 / CV > 0 ? 256 / CV : 1;
  const int npix = Ho * Wo;
  const int per_thread = need_gap ? 16 : 8;  // output pixels per thread per CTA
  int chunks = ceil_div(npix, PL * per_thread);
  if (need_gap && chunks > 16) chunks = 16;
  if (chunks < 1) chunks = 1;
  return chunks;
}

int launch_dwconv(const DwConvArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.in && a.w && a.out && a.scale && a.shift, MTG_ERR_ARG, "dwconv: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && a.C >= 8 && a.C <= 2048, MTG_ERR_UNSUPPORTED, "dwconv: C=%d must be a multiple of 8 in [8,2048]", a.C);
  MTG_REQUIRE(a.k == 3 || a.k == 5, MTG_ERR_UNSUPPORTED, "dwconv: kernel size %d unsupported", a.k);
  DwP p{};
  p.in = a.in; p.w = a.w; p.out = a.out; p.scale = a.scale; p.shift = a.shift; p.gap = a.gap_partial; p.act = a.act;
  p.H = a.H; p.W = a.W; p.C = a.C; p.stride = a.stride; p.dil = a.dil;
  p.pad = (a.k - 1) / 2 * a.dil;
  p.Ho = (a.H + 2 * p.pad - a.dil * (a.k - 1) - 1) / a.stride + 1;
  p.Wo = (a.W + 2 * p.pad - a.dil * (a.k - 1) - 1) / a.stride + 1;
  p.CV = a.C / 8;
  MTG_REQUIRE(p.CV <= 256, MTG_ERR_UNSUPPORTED, "dwconv: C too large");
  p.PL = 256 / p.CV;
  p.chunks = a.chunks > 0 ? a.chunks : 1;
  p.pix_per_chunk = ceil_div(p.Ho * p.Wo, p.chunks);
  dim3 grid(p.chunks, a.B);
  if (a.k == 3) dwconv_kernel<3><<<grid, 256, 0, st>>>(p);
  else dwconv_kernel<5><<<grid, 256, 0, st>>>(p);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
