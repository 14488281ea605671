// Depthwise k x k convolution (k = 3 or 5; stride 1/2; dilation 1/2) on NHWC bf16 with the eval-mode
// BatchNorm scale/shift and activation fused, plus optional per-chunk channel sums of the output (the
// squeeze-excite global-average-pool, so the SE block never re-reads the tensor).
//
// Bandwidth-bound: every thread owns one 8-channel vector (128-bit loads/stores) and a horizontal strip of
// output pixels; consecutive threads own consecutive vectors, so each warp touches contiguous 512-byte
// spans.  The block size is a multiple of the vectors-per-pixel count so a thread's channels (hence its
// weights and BN constants) never change.
//
// Replaces the depthwise Conv2dNormActivation of tv:models/mobilenetv3.py:83-95 (+ the AdaptiveAvgPool2d
// of tv:ops/misc.py:252-253).
#include <stdlib.h>

#include "ops.h"

namespace mtgseg {
namespace {

struct DwP {
  const bf16* in; const bf16* w; bf16* out;
  const float* scale; const float* shift;
  float* gap;
  int act, H, W, C, Ho, Wo, pad, CV, CVc, PL, strips, items, items_per_chunk, chunks;
};

// One thread = one 8-channel vector x one horizontal strip of TW output pixels.  Per kernel row the strip's
// NI = (TW-1)*STRIDE + (KS-1)*DIL + 1 input vectors are streamed once and scattered into the TW accumulators
// (compile-time tap/offset matching), so the L1 traffic per output drops from KS*KS loads to ~KS*NI/TW.
template <int KS, int STRIDE, int DIL, int TW, int MINB>
__global__ void __launch_bounds__(256, MINB) dwconv_kernel(const DwP p) {
  __shared__ float red[256 * 8];
  constexpr int NI = (TW - 1) * STRIDE + (KS - 1) * DIL + 1;
  const int tid = threadIdx.x;
  const int n = blockIdx.z, chunk = blockIdx.x;
  // a CTA owns CVc channel vectors (<= 128 channels) of a band of rows: its input footprint fits L1
  const int vl = tid % p.CVc, pl_raw = tid / p.CVc;
  const int v_raw = blockIdx.y * p.CVc + vl;
  const bool active = pl_raw < p.PL && v_raw < p.CV;
  const int v = active ? v_raw : 0;
  const int pl = active ? pl_raw : 0;
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
    const int it_begin = chunk * p.items_per_chunk;
    const int it_end = min(p.items, it_begin + p.items_per_chunk);
    const bf16* in_n = p.in + static_cast<size_t>(n) * p.H * p.W * p.C + c0;
    bf16* out_n = p.out + static_cast<size_t>(n) * p.Ho * p.Wo * p.C + c0;
    for (int item = it_begin + pl; item < it_end; item += p.PL) {
      const int oy = item / p.strips, sx = item - oy * p.strips;
      const int ox0 = sx * TW;
      const int ix0 = ox0 * STRIDE - p.pad;
      float acc[TW][8];
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll 1  // keep the body (one kernel row) small enough for the instruction cache
      for (int ky = 0; ky < KS; ++ky) {
        const int iy = oy * STRIDE - p.pad + ky * DIL;
        if (iy < 0 || iy >= p.H) continue;
        float wv[KS][8];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) unpack8(ldg16(p.w + (ky * KS + kx) * p.C + c0), wv[kx]);
        const bf16* row = in_n + static_cast<size_t>(iy) * p.W * p.C;
#pragma unroll
        for (int xi = 0; xi < NI; ++xi) {
          const int ix = ix0 + xi;
          if (ix < 0 || ix >= p.W) continue;
          float xf[8];
          unpack8(ldg16(row + static_cast<size_t>(ix) * p.C), xf);
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            constexpr int dummy = 0;
            (void)dummy;
            const int t = xi - kx * DIL;  // compile-time after unrolling
            if (t >= 0 && t % STRIDE == 0 && t / STRIDE < TW) {
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[t / STRIDE][j] = fmaf(xf[j], wv[kx][j], acc[t / STRIDE][j]);
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < TW; ++t) {
        if (ox0 + t < p.Wo) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = apply_act(fmaf(acc[t][j], sc[j], sh[j]), p.act);
          const uint4 packed = pack8(o);
          *reinterpret_cast<uint4*>(out_n + (static_cast<size_t>(oy) * p.Wo + ox0 + t) * p.C) = packed;
          if (p.gap) {  // pool what the next layer will actually read (the bf16-rounded values)
            float rf[8];
            unpack8(packed, rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc_gap[j] += rf[j];
          }
        }
      }
    }
  }
  if (p.gap) {
    const int cw = p.CVc * 8;  // channels of this CTA
    if (pl_raw < p.PL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(pl_raw * p.CVc + vl) * 8 + j] = active ? acc_gap[j] : 0.f;
    }
    __syncthreads();
    for (int cl = tid; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < p.C) {
        float s = 0.f;
        for (int l = 0; l < p.PL; ++l) s += red[l * cw + cl];  // fixed order -> deterministic
        p.gap[(static_cast<size_t>(n) * p.chunks + chunk) * p.C + c] = s;
      }
    }
  }
}

// MTGSEG_DW_VARIANT=1 selects the wide-strip instantiations (tuning aid; the narrow strips are the default)
int dw_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTGSEG_DW_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}
inline int strip_width(int stride) { return dw_variant() == 0 ? (stride == 1 ? 4 : 2) : (stride == 1 ? 8 : 4); }

}  // namespace

// vectors per CTA: <= 16 (128 channels), chosen in [8,16] to waste the fewest lanes
int group_vectors(int CV) {
  if (CV <= 16) return CV;
  int best = 16, best_waste = 1 << 30;
  for (int c = 16; c >= 8; --c) {
    const int waste = ceil_div(CV, c) * c - CV;
    if (waste < best_waste) { best_waste = waste; best = c; }
  }
  return best;
}

int dwconv_chunks(int Ho, int Wo, int C, int stride, bool need_gap) {
  const int CV = C / 8;
  const int PL = 256 / group_vectors(CV);
  const int items = Ho * ceil_div(Wo, strip_width(stride));
  const int per_thread = need_gap ? 4 : 2;  // strips per thread per CTA
  int chunks = ceil_div(items, PL * per_thread);
  if (need_gap && chunks > 16) chunks = 16;
  if (chunks < 1) chunks = 1;
  return chunks;
}

int launch_dwconv(const DwConvArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.in && a.w && a.out && a.scale && a.shift, MTG_ERR_ARG, "dwconv: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && a.C >= 8 && a.C <= 2048, MTG_ERR_UNSUPPORTED, "dwconv: C=%d must be a multiple of 8 in [8,2048]", a.C);
  DwP p{};
  p.in = a.in; p.w = a.w; p.out = a.out; p.scale = a.scale; p.shift = a.shift; p.gap = a.gap_partial; p.act = a.act;
  p.H = a.H; p.W = a.W; p.C = a.C;
  p.pad = (a.k - 1) / 2 * a.dil;
  p.Ho = (a.H + 2 * p.pad - a.dil * (a.k - 1) - 1) / a.stride + 1;
  p.Wo = (a.W + 2 * p.pad - a.dil * (a.k - 1) - 1) / a.stride + 1;
  p.CV = a.C / 8;
  p.CVc = group_vectors(p.CV);
  p.PL = 256 / p.CVc;
  p.strips = ceil_div(p.Wo, strip_width(a.stride));
  p.items = p.Ho * p.strips;
  p.chunks = a.chunks > 0 ? a.chunks : 1;
  p.items_per_chunk = ceil_div(p.items, p.chunks);
  dim3 grid(p.chunks, ceil_div(p.CV, p.CVc), a.B);
  const int key = a.k * 100 + a.stride * 10 + a.dil;
  if (dw_variant() == 0) {  // default: narrow strips, <= 128 registers, two CTAs per SM (measured faster)
    switch (key) {
      case 311: dwconv_kernel<3, 1, 1, 4, 2><<<grid, 256, 0, st>>>(p); break;
      case 321: dwconv_kernel<3, 2, 1, 2, 2><<<grid, 256, 0, st>>>(p); break;
      case 511: dwconv_kernel<5, 1, 1, 4, 2><<<grid, 256, 0, st>>>(p); break;
      case 521: dwconv_kernel<5, 2, 1, 2, 2><<<grid, 256, 0, st>>>(p); break;
      case 512: dwconv_kernel<5, 1, 2, 4, 2><<<grid, 256, 0, st>>>(p); break;
      default:
        MTG_REQUIRE(false, MTG_ERR_UNSUPPORTED, "dwconv: (k=%d, stride=%d, dilation=%d) is not one of the MobileNetV3 shapes", a.k, a.stride, a.dil);
    }
  } else {
    switch (key) {
      case 311: dwconv_kernel<3, 1, 1, 8, 1><<<grid, 256, 0, st>>>(p); break;
      case 321: dwconv_kernel<3, 2, 1, 4, 1><<<grid, 256, 0, st>>>(p); break;
      case 511: dwconv_kernel<5, 1, 1, 8, 1><<<grid, 256, 0, st>>>(p); break;
      case 521: dwconv_kernel<5, 2, 1, 4, 1><<<grid, 256, 0, st>>>(p); break;
      case 512: dwconv_kernel<5, 1, 2, 8, 1><<<grid, 256, 0, st>>>(p); break;
      default:
        MTG_REQUIRE(false, MTG_ERR_UNSUPPORTED, "dwconv: (k=%d, stride=%d, dilation=%d) is not one of the MobileNetV3 shapes", a.k, a.stride, a.dil);
    }
  }
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
