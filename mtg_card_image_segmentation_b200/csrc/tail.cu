// Decoder tail of the LR-ASPP head (train/model.py:137-142, tv:models/segmentation/lraspp.py:44-46).
//
//   head_mix:      lowres[b,y,x,c] = b_high[c] + b_low[c] + sum_k w_low[c,k] low[b,y,x,k]
//                                    + up2( sum_i w_high[c,i] * s[b,i] * cbr[b,.,.,i] )[y,x]
//                  (the 1x1 high classifier is linear, so it is applied BEFORE the x2 bilinear: 2 channels
//                   are interpolated instead of 128; bilinear weights sum to 1 so the bias commutes too)
//   upsample_out:  logits = bilinear(lowres -> H x W), written NCHW in fp32/bf16/fp16, and/or the argmax
//                  mask (uint8) and/or the 2x2 confusion counts against int64 targets -- one pass, pure
//                  write bandwidth.
// Bilinear semantics are ATen's upsample_bilinear2d with align_corners=False and size given:
//   src = max(0, (dst + 0.5) * in/out - 0.5), i0 = floor(src), i1 = min(i0 + 1, in - 1), l = src - i0.
#include <cuda_fp16.h>

#include "ops.h"

namespace mtgseg {
namespace {

constexpr int MAX_NC = 8;

__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

struct MixP {
  const bf16* cbr; const float* s; const bf16* low;
  const float* w_high; const float* b_high; const float* w_low; const float* b_low;
  float* out;
  int Hh, Wh, Hl, Wl, IC, LC, NC;
};

template <int CMAX>
__global__ void __launch_bounds__(256) head_mix_kernel(const MixP p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* wsx = sm;                        // [NC][IC]  w_high * s[b]
  float* h2 = wsx + p.NC * p.IC;          // [Hh*Wh][NC]
  float* wl = h2 + p.Hh * p.Wh * p.NC;    // [NC][LC]
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < p.NC * p.IC; i += blockDim.x) wsx[i] = p.w_high[i] * p.s[static_cast<size_t>(b) * p.IC + i % p.IC];
  for (int i = threadIdx.x; i < p.NC * p.LC; i += blockDim.x) wl[i] = p.w_low[i];
  __syncthreads();
  // phase A: classifier at the high-level resolution, one thread per (pixel, class): a 128-channel dot product with the
  // gated weights in shared memory, no cross-lane reduction (the warp-per-pixel version was a serial chain of ~37
  // load -> FMA -> shuffle rounds per warp: 58 us for 44 MB)
  const int npix_h = p.Hh * p.Wh;
  for (int idx = threadIdx.x; idx < npix_h * p.NC; idx += blockDim.x) {
    const int pix = idx / p.NC, c = idx - pix * p.NC;
    const bf16* row = p.cbr + (static_cast<size_t>(b) * npix_h + pix) * p.IC;
    const float* wv = wsx + c * p.IC;
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < p.IC; i += 8) {
      float f[8];
      unpack8(ldg16(row + i), f);
      a0 = fmaf(f[0], wv[i], a0); a1 = fmaf(f[1], wv[i + 1], a1);
      a0 = fmaf(f[2], wv[i + 2], a0); a1 = fmaf(f[3], wv[i + 3], a1);
      a0 = fmaf(f[4], wv[i + 4], a0); a1 = fmaf(f[5], wv[i + 5], a1);
      a0 = fmaf(f[6], wv[i + 6], a0); a1 = fmaf(f[7], wv[i + 7], a1);
    }
    h2[idx] = a0 + a1;
  }
  __syncthreads();
  // phase B: one thread per low-level pixel
  const float sy = static_cast<float>(p.Hh) / p.Hl, sx = static_cast<float>(p.Wh) / p.Wl;
  const int npix_l = p.Hl * p.Wl;
  for (int pix = threadIdx.x; pix < npix_l; pix += blockDim.x) {
    const int y = pix / p.Wl, x = pix - y * p.Wl;
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(y, sy, p.Hh, y0, y1, ly);
    src_index(x, sx, p.Wh, x0, x1, lx);
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < p.NC) {
        const float v00 = h2[(y0 * p.Wh + x0) * p.NC + c], v01 = h2[(y0 * p.Wh + x1) * p.NC + c];
        const float v10 = h2[(y1 * p.Wh + x0) * p.NC + c], v11 = h2[(y1 * p.Wh + x1) * p.NC + c];
        acc[c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11) + p.b_high[c] + p.b_low[c];
      }
    const bf16* lrow = p.low + (static_cast<size_t>(b) * npix_l + pix) * p.LC;
    for (int k = 0; k < p.LC; k += 8) {
      float f[8];
      unpack8(ldg16(lrow + k), f);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < p.NC) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[c] = fmaf(f[e], wl[c * p.LC + k + e], acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < p.NC) p.out[(static_cast<size_t>(b) * npix_l + pix) * p.NC + c] = acc[c];
  }
}

struct UpP {
  const float* lowres; void* logits; int dtype; uint8_t* mask; const int64_t* targets; unsigned long long* counts;
  int Hl, Wl, H, W, NC, rows_per_cta;
};

// CMAX: compile-time class capacity (2 for the card/background network: no dead predicated code for 6 absent classes)
template <int CMAX>
__global__ void __launch_bounds__(256) upsample_out_kernel(const UpP p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float lo[];  // [Hl*Wl][NC]
  __shared__ unsigned long long scount[4];
  const int b = blockIdx.y;
  const int nlo = p.Hl * p.Wl * p.NC;
  for (int i = threadIdx.x; i < nlo; i += blockDim.x) lo[i] = p.lowres[static_cast<size_t>(b) * nlo + i];
  if (threadIdx.x < 4) scount[threadIdx.x] = 0ull;
  __syncthreads();
  const float sy = static_cast<float>(p.Hl) / p.H, sx = static_cast<float>(p.Wl) / p.W;
  const int W4 = (p.W + 3) / 4;
  const int y_begin = blockIdx.x * p.rows_per_cta;
  const int y_end = min(p.H, y_begin + p.rows_per_cta);
  const bool vec = (p.W % 4) == 0;
  unsigned int cnt[4] = {0u, 0u, 0u, 0u};
  const size_t plane = static_cast<size_t>(p.H) * p.W;
  for (int idx = threadIdx.x; idx < (y_end - y_begin) * W4; idx += blockDim.x) {
    const int y = y_begin + idx / W4, xb = (idx % W4) * 4;
    int y0, y1;
    float ly;
    src_index(y, sy, p.Hl, y0, y1, ly);
    float val[CMAX][4];
    int x0[4], x1[4];
    float lx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) src_index(min(xb + e, p.W - 1), sx, p.Wl, x0[e], x1[e], lx[e]);
    if (x0[0] == x0[3] && x1[0] == x1[3]) {
      // the four pixels interpolate inside the same source cell (always true for an aligned group of 4 at the x8 upsampling
      // of the network): fetch the four corners once, the per-pixel arithmetic is unchanged (bit-identical results)
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < p.NC) {
          const float v00 = lo[(y0 * p.Wl + x0[0]) * p.NC + c], v01 = lo[(y0 * p.Wl + x1[0]) * p.NC + c];
          const float v10 = lo[(y1 * p.Wl + x0[0]) * p.NC + c], v11 = lo[(y1 * p.Wl + x1[0]) * p.NC + c];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            val[c][e] = (1.f - ly) * ((1.f - lx[e]) * v00 + lx[e] * v01) + ly * ((1.f - lx[e]) * v10 + lx[e] * v11);
        }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < p.NC) {
            const float v00 = lo[(y0 * p.Wl + x0[e]) * p.NC + c], v01 = lo[(y0 * p.Wl + x1[e]) * p.NC + c];
            const float v10 = lo[(y1 * p.Wl + x0[e]) * p.NC + c], v11 = lo[(y1 * p.Wl + x1[e]) * p.NC + c];
            val[c][e] = (1.f - ly) * ((1.f - lx[e]) * v00 + lx[e] * v01) + ly * ((1.f - lx[e]) * v10 + lx[e] * v11);
          }
      }
    }
    const size_t pix0 = static_cast<size_t>(y) * p.W + xb;
    if (p.logits) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < p.NC) {
          const size_t off = (static_cast<size_t>(b) * p.NC + c) * plane + pix0;
          if (p.dtype == LOGITS_F32) {
            float* d = static_cast<float*>(p.logits) + off;
            if (vec) *reinterpret_cast<float4*>(d) = make_float4(val[c][0], val[c][1], val[c][2], val[c][3]);
            else for (int e = 0; e < 4 && xb + e < p.W; ++e) d[e] = val[c][e];
          } else if (p.dtype == LOGITS_BF16) {
            bf16* d = static_cast<bf16*>(p.logits) + off;
            if (vec) *reinterpret_cast<uint2*>(d) = make_uint2(pack2(val[c][0], val[c][1]), pack2(val[c][2], val[c][3]));
            else for (int e = 0; e < 4 && xb + e < p.W; ++e) d[e] = __float2bfloat16(val[c][e]);
          } else {
            __half* d = static_cast<__half*>(p.logits) + off;
            for (int e = 0; e < 4 && xb + e < p.W; ++e) d[e] = __float2half(val[c][e]);
          }
        }
    }
    if (p.mask || p.counts) {
      uint32_t m4 = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int best = 0;
        float bv = val[0][e];
#pragma unroll
        for (int c = 1; c < CMAX; ++c)
          if (c < p.NC && val[c][e] > bv) { bv = val[c][e]; best = c; }  // strict '>' : ties -> lowest class (torch.argmax)
        m4 |= static_cast<uint32_t>(best) << (8 * e);
        if (p.counts && xb + e < p.W) {
          const int t = static_cast<int>(p.targets[static_cast<size_t>(b) * plane + pix0 + e]);
          cnt[(t & 1) * 2 + (best & 1)]++;
        }
      }
      if (p.mask) {
        uint8_t* d = p.mask + static_cast<size_t>(b) * plane + pix0;
        if (vec) *reinterpret_cast<uint32_t*>(d) = m4;
        else for (int e = 0; e < 4 && xb + e < p.W; ++e) d[e] = static_cast<uint8_t>(m4 >> (8 * e));
      }
    }
  }
  if (p.counts) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned int v = cnt[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&scount[k], static_cast<unsigned long long>(v));
    }
    __syncthreads();
    if (threadIdx.x < 4 && scount[threadIdx.x]) atomicAdd(&p.counts[threadIdx.x], scount[threadIdx.x]);
  }
}

}  // namespace

int launch_head_mix(const HeadMixArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.cbr && a.s && a.low && a.w_high && a.b_high && a.w_low && a.b_low && a.out, MTG_ERR_ARG, "head_mix: null pointer");
  MTG_REQUIRE(a.NC >= 1 && a.NC <= MAX_NC, MTG_ERR_UNSUPPORTED, "head_mix: num_classes %d not in [1,%d]", a.NC, MAX_NC);
  MTG_REQUIRE(a.IC % 8 == 0 && a.LC % 8 == 0, MTG_ERR_UNSUPPORTED, "head_mix: channel counts must be multiples of 8");
  MixP p{a.cbr, a.s, a.low, a.w_high, a.b_high, a.w_low, a.b_low, a.out, a.Hh, a.Wh, a.Hl, a.Wl, a.IC, a.LC, a.NC};
  const size_t smem = sizeof(float) * (static_cast<size_t>(a.NC) * a.IC + static_cast<size_t>(a.Hh) * a.Wh * a.NC + static_cast<size_t>(a.NC) * a.LC);
  MTG_REQUIRE(smem <= 200 * 1024, MTG_ERR_UNSUPPORTED, "head_mix: feature map too large for one CTA (%zu B)", smem);
  static bool configured = false;
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(head_mix_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MTG_CUDA(cudaFuncSetAttribute(head_mix_kernel<MAX_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (a.NC <= 2) MTG_CUDA(launch_pdl(head_mix_kernel<2>, dim3(a.B), dim3(256), smem, st, p));
  else MTG_CUDA(launch_pdl(head_mix_kernel<MAX_NC>, dim3(a.B), dim3(256), smem, st, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_upsample_out(const UpsampleOutArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.lowres && (a.logits || a.mask || a.counts), MTG_ERR_ARG, "upsample_out: nothing to do");
  MTG_REQUIRE(a.NC >= 1 && a.NC <= MAX_NC, MTG_ERR_UNSUPPORTED, "upsample_out: num_classes %d not in [1,%d]", a.NC, MAX_NC);
  MTG_REQUIRE(!a.counts || (a.targets && a.NC == 2), MTG_ERR_UNSUPPORTED, "upsample_out: counts need targets and num_classes == 2");
  UpP p{a.lowres, a.logits, a.logits_dtype, a.mask, a.targets, a.counts, a.Hl, a.Wl, a.H, a.W, a.NC, 0};
  const size_t smem = sizeof(float) * static_cast<size_t>(a.Hl) * a.Wl * a.NC;
  MTG_REQUIRE(smem <= 200 * 1024, MTG_ERR_UNSUPPORTED, "upsample_out: low-res map too large (%zu B)", smem);
  static bool configured = false;
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(upsample_out_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MTG_CUDA(cudaFuncSetAttribute(upsample_out_kernel<MAX_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  // ~8 CTAs per image keeps the smem staging of the low-res map (<= 10 KB) negligible vs the rows written
  int row_blocks = ceil_div(a.H, 40);
  p.rows_per_cta = ceil_div(a.H, row_blocks);
  dim3 grid(row_blocks, a.B);
  if (a.NC <= 2) MTG_CUDA(launch_pdl(upsample_out_kernel<2>, dim3(grid), dim3(256), smem, st, p));
  else MTG_CUDA(launch_pdl(upsample_out_kernel<MAX_NC>, dim3(grid), dim3(256), smem, st, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
