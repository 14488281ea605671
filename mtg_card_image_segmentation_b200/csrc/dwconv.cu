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
#include <cuda.h>
#include <stdlib.h>

#include "ops.h"
#include "ptx.cuh"

namespace mtgseg {

int make_tma_map_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, int kbox);  // gemm_tc.cu

// column-strip kernel (dwcol.cu)
struct DwColPlan { bool ok; int P, TH, pad, Ho, Wo, band, bands, R, Wp, groups; size_t stage_bytes; };
DwColPlan dw_col_plan(int H, int W, int C, int k, int stride, int dil, bool need_gap);
int launch_dwconv_col(const DwConvArgs& a, const DwColPlan& q, cudaStream_t st);

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

// ---------------------------------------------------------------------------------------------------------
// Shared-memory staged variant (default): a CTA owns CVc channel vectors of a band of output rows of one image.
// Phase 1 streams the band's input rows (zero padded in x and y, so phase 2 needs no bounds checks) and the
// k*k weight vectors into shared memory with 16-byte cp.async (many loads in flight, no registers held);
// phase 2 is the same register-strip arithmetic as above, fed by LDS.128 instead of L1/L2 round trips.
// Every input byte leaves L2 exactly once per band (+ halo rows).
// ---------------------------------------------------------------------------------------------------------
struct DwS {
  const bf16* in; const bf16* w; bf16* out;
  const float* scale; const float* shift;
  float* gap;
  int act, H, W, C, Ho, Wo, pad, CV, CVc, PL, strips, band, bands, R, Wp;
  int xoff;  // offset of the input tile inside the dynamic buffer, in uint4 (weights first, padded to 128 bytes for TMA)
  int dbg;   // MTGSEG_DW_PHASE (timing experiments only): 1 = fill phase only, 2 = compute phase only (reads an unfilled tile)
  double* stat;  // training (kernels instantiated with ST): [2][C] fp64, += per-channel sum / sum of squares of the stored outputs
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;  // src-size 0 -> the 16 bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}

// acc[0..7] += x[0..7] * w[0..7]; F2 issues the four channel pairs as packed fp32x2 FMAs (FFMA2 on sm_100).
template <bool F2>
__device__ __forceinline__ void fma8(float (&acc)[8], const float (&x)[8], const float (&w)[8]) {
  if constexpr (F2) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      uint64_t a, xv, wv;
      asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(acc[j]), "f"(acc[j + 1]));
      asm("mov.b64 %0, {%1,%2};" : "=l"(xv) : "f"(x[j]), "f"(x[j + 1]));
      asm("mov.b64 %0, {%1,%2};" : "=l"(wv) : "f"(w[j]), "f"(w[j + 1]));
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(xv), "l"(wv));
      asm("mov.b64 {%0,%1}, %2;" : "=f"(acc[j]), "=f"(acc[j + 1]) : "l"(a));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(x[j], w[j], acc[j]);
  }
}

// acc[0..7] += x[0..7] * w[0..7] on packed bf16 operands: eight FHFMA.BF16 with half-register selectors
__device__ __forceinline__ void fh8(float (&acc)[8], const uint4& x, const uint4& w) {
  const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
        "mov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
        "fma.rn.f32.bf16 %0, xl, wl, %0;\n\tfma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
        : "+f"(acc[2 * i]), "+f"(acc[2 * i + 1]) : "r"(xs[i]), "r"(ws[i]));
}

// TMA: phase 1 is ONE 4-D box load of the band's input rows (out-of-bounds rows / columns / channels are zero filled by
// the TMA unit = the convolution padding) plus one 2-D box load of the weights, issued by thread 0 and awaited on an
// mbarrier; no per-thread address arithmetic.  !TMA: the same tile gathered with 16-byte cp.async (kept for A/B).
// ST: training variant -- instead of the SE pool partials the epilogue accumulates the BatchNorm statistics (sum, sum of squares
// of the bf16-rounded outputs) and adds them to p.stat with one fp64 atomic per channel and CTA.
template <int KS, int STRIDE, int DIL, int TW, bool TMA, bool F2, bool ST = false>
__global__ void __launch_bounds__(256, 2) dwconv_smem_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                                                              const DwS p) {
  extern __shared__ __align__(128) uint4 dsm[];
  __shared__ uint64_t bar;
  pdl_trigger();
  constexpr int NI = (TW - 1) * STRIDE + (KS - 1) * DIL + 1;
  uint4* sw = dsm;              // [KS*KS][CVc]
  uint4* sx = dsm + p.xoff;     // [R][Wp][CVc]
  float* red = reinterpret_cast<float*>(dsm);  // SE pool partials reuse the buffer after phase 2 (<= 8 KB)
  const int tid = threadIdx.x;
  const int n = blockIdx.z, band = blockIdx.x;
  const int v0 = blockIdx.y * p.CVc;
  const int nv = min(p.CVc, p.CV - v0);  // vectors of this group that exist
  const int oy0 = band * p.band, oy1 = min(p.Ho, oy0 + p.band);
  const int iy_base = oy0 * STRIDE - p.pad;
  // ---- phase 1: stage ----
  if constexpr (TMA) {
    if (p.dbg != 2) {
      if (tid == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_mbar_init();
      }
      __syncthreads();
      pdl_wait();
      if (tid == 0) {
        ptx::mbar_arrive_expect_tx(&bar, static_cast<uint32_t>((KS * KS + p.R * p.Wp) * p.CVc) * 16u);
        ptx::tma_load_2d(sw, &tmw, &bar, v0 * 8, 0);
        ptx::tma_load_4d(sx, &tmx, &bar, v0 * 8, -p.pad, iy_base, n);
      }
      ptx::mbar_wait(&bar, 0);
    }
    if (p.dbg == 1) return;
  } else {
    pdl_wait();
    const int rows = (oy1 - oy0 - 1) * STRIDE + (KS - 1) * DIL + 1;
    const bf16* in_n = p.in + static_cast<size_t>(n) * p.H * p.W * p.C + v0 * 8;
    const int per_row = p.Wp * p.CVc;
    for (int i = tid; i < rows * per_row; i += 256) {
      const int r = i / per_row, rem = i - r * per_row;
      const int xp = rem / p.CVc, vl = rem - xp * p.CVc;
      const int iy = iy_base + r, ix = xp - p.pad;
      const bool ok = vl < nv && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      const bf16* src = ok ? in_n + (static_cast<size_t>(iy) * p.W + ix) * p.C + vl * 8 : p.in;
      cp_async16(sx + i, src, ok);
    }
    for (int i = tid; i < KS * KS * p.CVc; i += 256) {
      const int t = i / p.CVc, vl = i - t * p.CVc;
      const bool ok = vl < nv;
      cp_async16(sw + i, ok ? p.w + t * p.C + (v0 + vl) * 8 : p.w, ok);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  // ---- phase 2: compute ----
  const int vl = tid % p.CVc, pl = tid / p.CVc;
  const bool active = pl < p.PL && vl < nv;
  const int c0 = (v0 + (active ? vl : 0)) * 8;
  float acc_gap[8], acc_sq[ST ? 8 : 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_gap[j] = 0.f;
#pragma unroll
  for (int j = 0; j < (ST ? 8 : 1); ++j) acc_sq[j] = 0.f;
  if (active) {
    float sc[8], sh[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.scale + c0)), b = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + 4));
      const float4 c = __ldg(reinterpret_cast<const float4*>(p.shift + c0)), d = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
      sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
      sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
    }
    bf16* out_n = p.out + static_cast<size_t>(n) * p.Ho * p.Wo * p.C + c0;
    const int items = (oy1 - oy0) * p.strips;
    for (int item = pl; item < items; item += p.PL) {
      const int ry = item / p.strips, sxi = item - ry * p.strips;
      const int ox0 = sxi * TW;
      float acc[TW][8];
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll 1
      for (int ky = 0; ky < KS; ++ky) {
        const uint4* row = sx + (static_cast<size_t>(ry * STRIDE + ky * DIL) * p.Wp + ox0 * STRIDE) * p.CVc + vl;
        if constexpr (F2) {
          // FHFMA.BF16 (bf16 x bf16 + fp32, see dwcol.cu): weights and activations are used as loaded, no unpack instructions;
          // bit-identical to the fp32 FMA on widened operands
          uint4 wq[KS];
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) wq[kx] = sw[(ky * KS + kx) * p.CVc + vl];
#pragma unroll
          for (int xi = 0; xi < NI; ++xi) {
            const uint4 xq = row[xi * p.CVc];
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              const int t = xi - kx * DIL;  // compile-time after unrolling
              if (t >= 0 && t % STRIDE == 0 && t / STRIDE < TW) fh8(acc[t / STRIDE], xq, wq[kx]);
            }
          }
        } else {
          float wv[KS][8];
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) unpack8(sw[(ky * KS + kx) * p.CVc + vl], wv[kx]);
#pragma unroll
          for (int xi = 0; xi < NI; ++xi) {
            float xf[8];
            unpack8(row[xi * p.CVc], xf);
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              const int t = xi - kx * DIL;  // compile-time after unrolling
              if (t >= 0 && t % STRIDE == 0 && t / STRIDE < TW) fma8<false>(acc[t / STRIDE], xf, wv[kx]);
            }
          }
        }
      }
      // the activation is uniform over the launch: pick it once per strip, not once per element
      auto finish = [&](auto actf) {
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          if (ox0 + t < p.Wo) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = actf(fmaf(acc[t][j], sc[j], sh[j]));
            const uint4 packed = pack8(o);
            *reinterpret_cast<uint4*>(out_n + (static_cast<size_t>(oy0 + ry) * p.Wo + ox0 + t) * p.C) = packed;
            if (ST) {
              float rf[8];
              unpack8(packed, rf);
#pragma unroll
              for (int j = 0; j < 8; ++j) { acc_gap[j] += rf[j]; acc_sq[j] = fmaf(rf[j], rf[j], acc_sq[j]); }
            } else if (p.gap) {
              float rf[8];
              unpack8(packed, rf);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc_gap[j] += rf[j];
            }
          }
        }
      };
      if (p.act == ACT_HSWISH) finish([](float v) { return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f); });
      else if (p.act == ACT_RELU) finish([](float v) { return fmaxf(v, 0.f); });
      else if (p.act == ACT_NONE) finish([](float v) { return v; });
      else finish([&](float v) { return apply_act(v, p.act); });
    }
  }
  if (ST) {
    const int cw = p.CVc * 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      __syncthreads();  // (k = 0: every thread is done with the staged tile: its first 8 KB become the reduction buffer)
      if (pl < p.PL) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(pl * p.CVc + vl) * 8 + j] = active ? (k == 0 ? acc_gap[j] : acc_sq[j]) : 0.f;
      }
      __syncthreads();
      for (int cl = tid; cl < cw; cl += blockDim.x) {
        const int c = blockIdx.y * cw + cl;
        if (c < p.C) {
          float s = 0.f;
          for (int l = 0; l < p.PL; ++l) s += red[l * cw + cl];
          atomicAdd(p.stat + k * p.C + c, static_cast<double>(s));
        }
      }
    }
  } else if (p.gap) {
    const int cw = p.CVc * 8;
    __syncthreads();  // every thread is done with the staged tile: its first 8 KB become the reduction buffer
    if (pl < p.PL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(pl * p.CVc + vl) * 8 + j] = active ? acc_gap[j] : 0.f;
    }
    __syncthreads();
    for (int cl = tid; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < p.C) {
        float s = 0.f;
        for (int l = 0; l < p.PL; ++l) s += red[l * cw + cl];
        p.gap[(static_cast<size_t>(n) * p.bands + band) * p.C + c] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Half-vector variant for the stride-1 5x5 layers: a thread owns FOUR channels (8-byte LDS / STG) and a strip of EIGHT
// output pixels.  Same 32 accumulators as 8 channels x 4 pixels, but every staged input value now feeds up to 5 taps of 8
// outputs: per kernel row 16 (dilation 2) or 12 (dilation 1) 8-byte loads and 4-element unpacks for 160 FMAs per lane,
// instead of 12 / 8 16-byte loads and 8-element unpacks: ~38 % fewer shared-memory bytes and unpack instructions per FMA
// (the 8-channel kernel sits where shared-memory bandwidth and issue bandwidth meet; 8 pixels x 8 channels spills).
// Same tile layout and TMA fill as dwconv_smem_kernel; the plan is made with TW = 8.
// ---------------------------------------------------------------------------------------------------------
template <int KS, int DIL, bool ST = false>
__global__ void __launch_bounds__(256, 2) dwconv_half_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                                                              const DwS p) {
  constexpr int TW = 8, NI = (TW - 1) + (KS - 1) * DIL + 1;
  extern __shared__ __align__(128) uint4 dsm[];
  __shared__ uint64_t bar;
  const uint2* sw = reinterpret_cast<const uint2*>(dsm);            // [KS*KS][HV]
  const uint2* sx = reinterpret_cast<const uint2*>(dsm + p.xoff);   // [R][Wp][HV]
  float* red = reinterpret_cast<float*>(dsm);
  const int tid = threadIdx.x;
  const int n = blockIdx.z, band = blockIdx.x;
  const int v0 = blockIdx.y * p.CVc;
  const int nv = min(p.CVc, p.CV - v0);
  const int HV = p.CVc * 2, nvh = nv * 2, PLh = 256 / HV;  // half-vectors per pixel of this group, pixel lanes
  const int oy0 = band * p.band, oy1 = min(p.Ho, oy0 + p.band);
  pdl_trigger();
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(&bar, static_cast<uint32_t>((KS * KS + p.R * p.Wp) * p.CVc) * 16u);
    ptx::tma_load_2d(dsm, &tmw, &bar, v0 * 8, 0);
    ptx::tma_load_4d(dsm + p.xoff, &tmx, &bar, v0 * 8, -p.pad, oy0 - p.pad, n);
  }
  ptx::mbar_wait(&bar, 0);
  const int vl = tid % HV, pl = tid / HV;
  const bool active = pl < PLh && vl < nvh;
  const int c0 = v0 * 8 + (active ? vl : 0) * 4;
  float acc_gap[4] = {0.f, 0.f, 0.f, 0.f}, acc_sq[4] = {0.f, 0.f, 0.f, 0.f};
  // one bf16x2 word -> (even channel, odd channel) as packed fp32x2
  auto widen = [](uint32_t w) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(w << 16), "r"(w & 0xFFFF0000u));
    return r;
  };
  if (active) {
    const float4 sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + c0)), sh4 = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
    bf16* out_n = p.out + static_cast<size_t>(n) * p.Ho * p.Wo * p.C + c0;
    const int items = (oy1 - oy0) * p.strips;
    for (int item = pl; item < items; item += PLh) {
      const int ry = item / p.strips, sxi = item - ry * p.strips;
      const int ox0 = sxi * TW;
      uint64_t acc[TW][2];
#pragma unroll
      for (int t = 0; t < TW; ++t) acc[t][0] = acc[t][1] = 0ull;
#pragma unroll 1
      for (int ky = 0; ky < KS; ++ky) {
        uint64_t wv[KS][2];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const uint2 q = sw[(ky * KS + kx) * HV + vl];
          wv[kx][0] = widen(q.x); wv[kx][1] = widen(q.y);
        }
        const uint2* row = sx + (static_cast<size_t>(ry + ky * DIL) * p.Wp + ox0) * HV + vl;
#pragma unroll
        for (int xi = 0; xi < NI; ++xi) {
          const uint2 q = row[xi * HV];
          const uint64_t x0 = widen(q.x), x1 = widen(q.y);
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const int t = xi - kx * DIL;  // compile-time after unrolling
            if (t >= 0 && t < TW) {
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[t][0]) : "l"(x0), "l"(wv[kx][0]));
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[t][1]) : "l"(x1), "l"(wv[kx][1]));
            }
          }
        }
      }
      auto finish = [&](auto actf) {
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          if (ox0 + t < p.Wo) {
            float a[4];
            asm("mov.b64 {%0,%1}, %2;" : "=f"(a[0]), "=f"(a[1]) : "l"(acc[t][0]));
            asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2]), "=f"(a[3]) : "l"(acc[t][1]));
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = actf(fmaf(a[j], sc[j], sh[j]));
            const uint2 packed = make_uint2(pack2(o[0], o[1]), pack2(o[2], o[3]));
            *reinterpret_cast<uint2*>(out_n + (static_cast<size_t>(oy0 + ry) * p.Wo + ox0 + t) * p.C) = packed;
            if (ST || p.gap) {
              const float r[4] = {__uint_as_float(packed.x << 16), __uint_as_float(packed.x & 0xFFFF0000u),
                                  __uint_as_float(packed.y << 16), __uint_as_float(packed.y & 0xFFFF0000u)};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc_gap[j] += r[j];
                if (ST) acc_sq[j] = fmaf(r[j], r[j], acc_sq[j]);
              }
            }
          }
        }
      };
      if (p.act == ACT_HSWISH) finish([](float v) { return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f); });
      else if (p.act == ACT_RELU) finish([](float v) { return fmaxf(v, 0.f); });
      else if (p.act == ACT_NONE) finish([](float v) { return v; });
      else finish([&](float v) { return apply_act(v, p.act); });
    }
  }
  if (ST) {
    const int cw = p.CVc * 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      __syncthreads();
      if (pl < PLh) {
#pragma unroll
        for (int j = 0; j < 4; ++j) red[(pl * HV + vl) * 4 + j] = active ? (k == 0 ? acc_gap[j] : acc_sq[j]) : 0.f;
      }
      __syncthreads();
      for (int cl = tid; cl < cw; cl += blockDim.x) {
        const int c = blockIdx.y * cw + cl;
        if (c < p.C) {
          float s = 0.f;
          for (int l = 0; l < PLh; ++l) s += red[l * cw + cl];
          atomicAdd(p.stat + k * p.C + c, static_cast<double>(s));
        }
      }
    }
  } else if (p.gap) {
    const int cw = p.CVc * 8;
    __syncthreads();  // every thread is done with the staged tile: its first bytes become the reduction buffer (PLh * cw floats <= 4 KB)
    if (pl < PLh) {
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(pl * HV + vl) * 4 + j] = active ? acc_gap[j] : 0.f;
    }
    __syncthreads();
    for (int cl = tid; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < p.C) {
        float s = 0.f;
        for (int l = 0; l < PLh; ++l) s += red[l * cw + cl];
        p.gap[(static_cast<size_t>(n) * p.bands + band) * p.C + c] = s;
      }
    }
  }
}

int dw_phase() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTGSEG_DW_PHASE");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// MTGSEG_DW_VARIANT (A/B switch): 0 (default) column-strip kernel or staged kernel per layer (use_col), TMA box fill, FHFMA.BF16; 4 = staged kernel with unpack + scalar fp32 FMAs; 3 staged kernel,
// cp.async gather fill, scalar FMAs; 5 = 3 with packed FMAs (all four bit-identical; B=256 family time 1.57 / 1.58 / 2.03 / 2.03 ms); 2 legacy direct kernel, narrow strips; 1 legacy direct kernel, wide strips
int dw_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTGSEG_DW_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}
// 8-wide strips at 8 channels per thread were measured and rejected: 128 registers spill (b14 213 -> 317 us).
// Half-vector kernel + 8-vector channel groups: default for the dilated stride-1 5x5 layers (b13-b15).  Measured per layer at
// B=256 (tools/dw_probe.py, outputs bit-identical): 8 ch/thread with 15-vector groups 159 / 214 / 214 us (31 % of the
// shared-memory load wavefronts are bank-conflict replays: a warp's 16-byte lanes straddle 240-byte runs), half vectors with
// 8-vector groups (every half-warp reads one aligned 128-byte run) 154 / 196 / 197 us; the undilated 5x5 layers (b5, b6) are
// faster with the 8-channel kernel (105 vs 113 us).  Variant 7 forces the half kernel on every stride-1 5x5 layer.
inline bool use_half(int k, int stride, int dil) {
  if (k != 5 || stride != 1) return false;
  return dw_variant() == 7 || (dw_variant() == 0 && dil == 2);
}
inline int strip_width(int stride, int k = 3, int dil = 1) {
  if (use_half(k, stride, dil)) return 8;
  return dw_variant() != 1 ? (stride == 1 ? 4 : 2) : (stride == 1 ? 8 : 4);
}

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

struct DwPlan { int pad, Ho, Wo, CV, CVc, PL, TW, strips, band, bands, R, Wp, xoff; size_t smem; bool ok; };

DwPlan dw_plan(int H, int W, int C, int k, int stride, int dil, bool need_gap) {
  DwPlan q{};
  q.pad = (k - 1) / 2 * dil;
  q.Ho = (H + 2 * q.pad - dil * (k - 1) - 1) / stride + 1;
  q.Wo = (W + 2 * q.pad - dil * (k - 1) - 1) / stride + 1;
  q.CV = C / 8; q.CVc = group_vectors(q.CV);
  if (use_half(k, stride, dil) && q.CVc > 8) q.CVc = 8;  // 16 half-vectors = one aligned 128-byte run per half-warp
  q.PL = 256 / q.CVc;
  q.TW = strip_width(stride, k, dil);
  q.strips = ceil_div(q.Wo, q.TW);
  // padded row: every strip (including the ragged last one) may read NI inputs from its first column
  const int NI = (q.TW - 1) * stride + (k - 1) * dil + 1;
  q.Wp = (q.strips - 1) * q.TW * stride + NI;
  if (q.Wp < W + 2 * q.pad) q.Wp = W + 2 * q.pad;
  // two CTAs per SM: 2 x (108 KB + 1 KB reserved) fits the 227 KB of an SM (the kernel has no static shared memory
  // besides one mbarrier; the SE-pool reduction reuses the tile)
  const size_t budget = 108 * 1024;
  q.xoff = (k * k * q.CVc + 7) / 8 * 8;  // weights first, the input tile starts 128-byte aligned (TMA destination)
  const size_t wbytes = static_cast<size_t>(q.xoff) * 16;
  auto bytes = [&](int band) { return (static_cast<size_t>((band - 1) * stride + (k - 1) * dil + 1) * q.Wp * q.CVc) * 16 + wbytes; };
  q.ok = bytes(1) <= budget;
  int band = 1;
  while (band < q.Ho && bytes(band + 1) <= budget) ++band;
  // keep at least ~2 strips per thread-lane of work and, for the SE pool, at most 16 partials per image
  int bands = ceil_div(q.Ho, band);
  if (need_gap && bands > 16) q.ok = false;
  band = ceil_div(q.Ho, bands);  // even bands
  q.band = band; q.bands = ceil_div(q.Ho, band);
  q.R = (band - 1) * stride + (k - 1) * dil + 1;
  q.smem = bytes(band) < 8192 ? 8192 : bytes(band);  // >= the 8 KB reduction buffer
  return q;
}

// Column-strip kernel (dwcol.cu) or the staged 8-channel kernels above?  Measured per layer at B=256 on one box (tools/dw_probe.py,
// column / staged, us): b1 80/129, b4 82/116, b5 72/111, b7 46/53, b8 26/34, b11 53/65, b12 67/89, b13 97/152, b14 144/196; but
// b2 204/138 and b3 154/109: 3x3 layers on large maps whose channel groups are narrower than a pixel (32 / 48-byte TMA box
// rows out of 128 / 144-byte pixels) stay on the staged kernel, which reads whole contiguous rows.
// MTGSEG_DW_VARIANT=8 forces the column-strip kernel wherever its tiling fits; 7 / 4 / 5 / 3 / 2 / 1 never use it.
static bool use_col(int H, int W, int C, int k, int stride, int dil, bool need_gap, DwColPlan* plan) {
  const int v = dw_variant();
  if (v != 0 && v != 8) return false;
  const DwColPlan q = dw_col_plan(H, W, C, k, stride, dil, need_gap);
  if (!q.ok) return false;
  if (v == 0 && k == 3 && q.groups > 1 && H * W >= 4096) return false;
  if (plan) *plan = q;
  return true;
}

int dwconv_chunks(int H, int W, int C, int k, int stride, int dil, bool need_gap) {
  {
    DwColPlan q;
    if (use_col(H, W, C, k, stride, dil, need_gap, &q)) return q.bands;
  }
  if (dw_variant() == 2) {  // legacy direct-from-L1 kernel
    const int pad = (k - 1) / 2 * dil;
    const int Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
    const int PL = 256 / group_vectors(C / 8);
    const int items = Ho * ceil_div(Wo, strip_width(stride));
    int chunks = ceil_div(items, PL * (need_gap ? 4 : 2));
    if (need_gap && chunks > 16) chunks = 16;
    return chunks < 1 ? 1 : chunks;
  }
  return dw_plan(H, W, C, k, stride, dil, need_gap).bands;
}

template <int KS, int STRIDE, int DIL, int TW, bool TMA, bool F2, bool ST = false>
int launch_smem2(const CUtensorMap& tmx, const CUtensorMap& tmw, const DwS& p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(dwconv_smem_kernel<KS, STRIDE, DIL, TW, TMA, F2, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  MTG_CUDA(launch_pdl(dwconv_smem_kernel<KS, STRIDE, DIL, TW, TMA, F2, ST>, grid, dim3(256), smem, st, tmx, tmw, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

template <int KS, int STRIDE, int DIL, int TW>
int launch_smem(const DwConvArgs& a, const DwS& p, dim3 grid, size_t smem, cudaStream_t st) {
  const int v = dw_variant();
  CUtensorMap tmx{}, tmw{};
  if (v == 0 || v == 4 || v == 7) {
    // input [B][H][W][C] bf16 as a 4-D tensor, box = (group channels, padded row, band rows + halo, 1 image);
    // weights [k*k][C] as a 2-D tensor, box = (group channels, all taps).  No swizzle: the tile is read as stored.
    const unsigned long long xd[4] = {static_cast<unsigned long long>(a.C), static_cast<unsigned long long>(a.W),
                                      static_cast<unsigned long long>(a.H), static_cast<unsigned long long>(a.B)};
    const unsigned long long xs[3] = {xd[0] * 2, xd[0] * xd[1] * 2, xd[0] * xd[1] * xd[2] * 2};
    const unsigned xb[4] = {static_cast<unsigned>(p.CVc * 8), static_cast<unsigned>(p.Wp), static_cast<unsigned>(p.R), 1u};
    MTG_REQUIRE(xb[1] <= 256 && xb[2] <= 256, MTG_ERR_UNSUPPORTED, "dwconv: tile %ux%u exceeds the TMA box limit", xb[1], xb[2]);
    int rc = make_tma_map_bf16(&tmx, a.in, 4, xd, xs, xb, 0);
    if (rc != MTG_OK) return rc;
    const unsigned long long wd[2] = {static_cast<unsigned long long>(a.C), static_cast<unsigned long long>(KS * KS)};
    const unsigned long long ws[1] = {wd[0] * 2};
    const unsigned wb[2] = {static_cast<unsigned>(p.CVc * 8), static_cast<unsigned>(KS * KS)};
    rc = make_tma_map_bf16(&tmw, a.w, 2, wd, ws, wb, 0);
    if (rc != MTG_OK) return rc;
  }
  if (p.stat) {
    MTG_REQUIRE(v == 0 || v == 7, MTG_ERR_UNSUPPORTED, "dwconv: BatchNorm statistics need the default kernel variant (MTGSEG_DW_VARIANT=0)");
    return launch_smem2<KS, STRIDE, DIL, TW, true, true, true>(tmx, tmw, p, grid, smem, st);
  }
  switch (v) {
    case 0: case 7: return launch_smem2<KS, STRIDE, DIL, TW, true, true>(tmx, tmw, p, grid, smem, st);
    case 4: return launch_smem2<KS, STRIDE, DIL, TW, true, false>(tmx, tmw, p, grid, smem, st);
    case 5: return launch_smem2<KS, STRIDE, DIL, TW, false, true>(tmx, tmw, p, grid, smem, st);
    default: return launch_smem2<KS, STRIDE, DIL, TW, false, false>(tmx, tmw, p, grid, smem, st);
  }
}

template <int KS, int DIL>
int launch_half(const DwConvArgs& a, const DwS& p, dim3 grid, size_t smem, cudaStream_t st) {
  CUtensorMap tmx{}, tmw{};
  const unsigned long long xd[4] = {static_cast<unsigned long long>(a.C), static_cast<unsigned long long>(a.W),
                                    static_cast<unsigned long long>(a.H), static_cast<unsigned long long>(a.B)};
  const unsigned long long xs[3] = {xd[0] * 2, xd[0] * xd[1] * 2, xd[0] * xd[1] * xd[2] * 2};
  const unsigned xb[4] = {static_cast<unsigned>(p.CVc * 8), static_cast<unsigned>(p.Wp), static_cast<unsigned>(p.R), 1u};
  MTG_REQUIRE(xb[1] <= 256 && xb[2] <= 256, MTG_ERR_UNSUPPORTED, "dwconv: tile %ux%u exceeds the TMA box limit", xb[1], xb[2]);
  int rc = make_tma_map_bf16(&tmx, a.in, 4, xd, xs, xb, 0);
  if (rc != MTG_OK) return rc;
  const unsigned long long wd[2] = {static_cast<unsigned long long>(a.C), static_cast<unsigned long long>(KS * KS)};
  const unsigned long long ws[1] = {wd[0] * 2};
  const unsigned wb[2] = {static_cast<unsigned>(p.CVc * 8), static_cast<unsigned>(KS * KS)};
  rc = make_tma_map_bf16(&tmw, a.w, 2, wd, ws, wb, 0);
  if (rc != MTG_OK) return rc;
  static bool configured = false;
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(dwconv_half_kernel<KS, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    MTG_CUDA(cudaFuncSetAttribute(dwconv_half_kernel<KS, DIL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  if (p.stat) MTG_CUDA(launch_pdl(dwconv_half_kernel<KS, DIL, true>, grid, dim3(256), smem, st, tmx, tmw, p));
  else MTG_CUDA(launch_pdl(dwconv_half_kernel<KS, DIL, false>, grid, dim3(256), smem, st, tmx, tmw, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_dwconv(const DwConvArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.in && a.w && a.out && a.scale && a.shift, MTG_ERR_ARG, "dwconv: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && a.C >= 8 && a.C <= 2048, MTG_ERR_UNSUPPORTED, "dwconv: C=%d must be a multiple of 8 in [8,2048]", a.C);
  {
    DwColPlan q;
    if (use_col(a.H, a.W, a.C, a.k, a.stride, a.dil, a.gap_partial != nullptr, &q)) return launch_dwconv_col(a, q, st);
  }
  if (dw_variant() != 2) {
    const DwPlan q = dw_plan(a.H, a.W, a.C, a.k, a.stride, a.dil, a.gap_partial != nullptr);
    MTG_REQUIRE(q.ok, MTG_ERR_UNSUPPORTED, "dwconv: feature map %dx%d (C=%d, k=%d) does not fit the shared-memory tiling", a.H, a.W, a.C, a.k);
    MTG_REQUIRE(!a.gap_partial || a.chunks == q.bands, MTG_ERR_ARG, "dwconv: gap_partial must have mtgseg_dwconv_chunks() = %d chunks, got %d", q.bands, a.chunks);
    MTG_REQUIRE(!(a.stat && a.gap_partial), MTG_ERR_ARG, "dwconv: statistics and pool partials are exclusive");
    DwS s{a.in, a.w, a.out, a.scale, a.shift, a.gap_partial, a.act, a.H, a.W, a.C, q.Ho, q.Wo, q.pad, q.CV, q.CVc, q.PL, q.strips,
          q.band, q.bands, q.R, q.Wp, q.xoff, dw_phase(), a.stat};
    dim3 grid(q.bands, ceil_div(q.CV, q.CVc), a.B);
    switch (a.k * 100 + a.stride * 10 + a.dil) {
      case 311: return launch_smem<3, 1, 1, 4>(a, s, grid, q.smem, st);
      case 321: return launch_smem<3, 2, 1, 2>(a, s, grid, q.smem, st);
      case 511: return use_half(5, 1, 1) ? launch_half<5, 1>(a, s, grid, q.smem, st) : launch_smem<5, 1, 1, 4>(a, s, grid, q.smem, st);
      case 521: return launch_smem<5, 2, 1, 2>(a, s, grid, q.smem, st);
      case 512: return use_half(5, 1, 2) ? launch_half<5, 2>(a, s, grid, q.smem, st) : launch_smem<5, 1, 2, 4>(a, s, grid, q.smem, st);
      default:
        MTG_REQUIRE(false, MTG_ERR_UNSUPPORTED, "dwconv: (k=%d, stride=%d, dilation=%d) is not one of the MobileNetV3 shapes", a.k, a.stride, a.dil);
    }
  }
  MTG_REQUIRE(!a.stat, MTG_ERR_UNSUPPORTED, "dwconv: BatchNorm statistics need the default kernel variant (MTGSEG_DW_VARIANT=0)");
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
  if (dw_variant() != 1) {  // narrow strips, <= 128 registers, two CTAs per SM
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
