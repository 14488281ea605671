// Depthwise k x k convolution, column-strip kernel (default for the layers use_col() in dwconv.cu names).
//
// What changed against dwconv_smem_kernel (dwconv.cu), and why (profiles/r2_forward_full.md: the compute phase of the staged
// kernel, not its fill, was the cost: ALU pipe 60 % of peak on bf16 -> fp32 unpacks, issue 61 %, shared-memory pipe 72 %):
//  * the multiply-add is `fma.rn.f32.bf16` (SASS FHFMA.BF16: bf16 x bf16 + fp32 -> fp32, half-register operand selectors):
//    a packed bf16 pair is used as loaded, there is no unpack instruction at all.  Measured issue rate = FFMA
//    (tools/ubench/fma_rate.cu: 3.3 warp instructions / clock / SM for both).  The product of two bf16 is exact in fp32 and the
//    addition rounds once, so the result is bit-identical to fmaf(float(x), float(w), acc) of the older kernels;
//  * a thread owns ONE channel pair and a vertical strip of TH output rows at one output column.  Consecutive threads own
//    consecutive (column, pair) words of the staged tile, so every LDS.32 of a warp reads 32 consecutive words (no bank
//    conflicts for any channel count) and every store of a warp is one 128-byte run.  The k*k packed weights of the pair live
//    in registers (k*k <= 25) for the life of the CTA: the inner loop loads only activations, k per staged row, each
//    feeding up to k output rows x 2 channels;
//  * a CTA owns one channel group and walks (image, band) tiles through 2-4 shared-memory stages: the 4-D TMA boxes of the
//    next tiles are in flight while tile i is computed (the staged kernel filled, waited, computed once per CTA);
//  * hardswish is z * sat(z/6 + 1/2) (FFMA.SAT + FMUL, the form of the GEMM epilogues), the squeeze-excite pool and the
//    training statistics are accumulated with FHFMA from the packed result (x * 1.0, x * x): 13 instructions per stored pair.
//
// Accumulation order per output is (ky, kx) ascending, the same as the older kernels: the convolution sums are bit-identical
// to theirs (the hardswish form differs in the last bit on some values); the squeeze-excite pool partials are per (image, band) in a fixed order, so results
// do not depend on the batch size or on which CTA computed a tile.
//
// Replaces the depthwise Conv2dNormActivation of tv:models/mobilenetv3.py:83-95 (+ the AdaptiveAvgPool2d of tv:ops/misc.py:252-253).
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "ops.h"
#include "ptx.cuh"

namespace mtgseg {

int make_tma_map_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, int kbox);  // gemm_tc.cu

namespace {

struct DwC {
  const bf16* w; bf16* out;
  const float* scale; const float* shift;
  float* gap;     // mode 1: [B][bands][C] per-band channel sums of the stored outputs
  double* stat;   // mode 2: [2][C] fp64, += sum / sum of squares of the stored outputs
  int mode, act, C, Ho, Wo, pad, band, bands, R, Wp, tiles, stage_words, stages;
};

// acc0 += x.lo * w.lo ; acc1 += x.hi * w.hi   (two FHFMA.BF16, operands selected as register halves)
__device__ __forceinline__ void fhfma2(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, wl, %0;\n\tfma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1) : "r"(x), "r"(w));
}

constexpr int kMaxStages = 4;

template <int KS, int S, int D, int P, int TH>
__global__ void __launch_bounds__(256, 2) dw_col_kernel(const __grid_constant__ CUtensorMap tmx, const DwC p) {
  constexpr int RI = (TH - 1) * S + (KS - 1) * D + 1;  // staged rows one strip reads
  constexpr int XL = 256 / P;                          // column lanes
  extern __shared__ __align__(128) uint32_t dyn[];
  __shared__ uint64_t full[kMaxStages];
  __shared__ float red[2 * XL * P];
  const int tid = threadIdx.x;
  const int pr = tid % P, xl = tid / P;
  const bool lane_ok = xl < XL;
  pdl_trigger();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kMaxStages; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  // CTA (x, g) owns channel group g and the (image, band) tiles x, x + gridDim.x, ...: the CTAs of one x walk the same
  // tiles at the same pace, so the sibling groups of a pixel are fetched while its line is still in L2 (a group-major order
  // was measured: every group pass re-read the whole tensor from DRAM, 4x traffic on b2)
  const int g = blockIdx.y;
  const int first = blockIdx.x, step = gridDim.x;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.R * p.Wp * P) * 4u;
  auto issue = [&](int t, int stage) {  // thread 0 only
    const int n = t / p.bands, band = t - n * p.bands;
    ptx::mbar_arrive_expect_tx(&full[stage], stage_bytes);
    ptx::tma_load_4d(dyn + stage * p.stage_words, &tmx, &full[stage], g * (2 * P), -p.pad, band * p.band * S - p.pad, n);
  };
  pdl_wait();
  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i)
      if (first + i * step < p.tiles) issue(first + i * step, i);
  }
  uint32_t wreg[KS * KS];
  float sc0 = 0.f, sc1 = 0.f, sh0 = 0.f, sh1 = 0.f;
  float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;  // mode 2: running sums over this CTA's tiles of one group
  const int c0 = g * (2 * P) + 2 * pr;
  const bool ch_ok = lane_ok && c0 < p.C;
  {
    const int cc = ch_ok ? c0 : 0;
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p.w + static_cast<size_t>(i) * p.C + cc));
      wreg[i] = ch_ok ? v : 0u;
    }
    const float2 a = __ldg(reinterpret_cast<const float2*>(p.scale + cc)), b = __ldg(reinterpret_cast<const float2*>(p.shift + cc));
    sc0 = a.x; sc1 = a.y; sh0 = b.x; sh1 = b.y;
  }
  const int row_words = p.Wp * P;

  auto flush_stats = [&]() {  // all threads; adds the CTA's sums to p.stat
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      __syncthreads();
      if (lane_ok) {
        red[(xl * P + pr) * 2] = k == 0 ? st_s0 : st_q0;
        red[(xl * P + pr) * 2 + 1] = k == 0 ? st_s1 : st_q1;
      }
      __syncthreads();
      if (tid < 2 * P) {
        const int c = g * (2 * P) + tid;
        if (c < p.C) {
          float s = 0.f;
          for (int l = 0; l < XL; ++l) s += red[l * 2 * P + tid];
          atomicAdd(p.stat + k * p.C + c, static_cast<double>(s));
        }
      }
    }
  };

  for (int t = first, stage = 0, phase = 0; t < p.tiles; t += step) {
    const int n = t / p.bands, band = t - n * p.bands;
    ptx::mbar_wait(&full[stage], phase);
    const uint32_t* tile = dyn + stage * p.stage_words;
    const int oy0 = band * p.band;
    const int nstrip = min(p.band, p.Ho - oy0 + TH - 1) / TH;  // strips of this band that hold at least one real row
    const int Q = nstrip * p.Wo;
    float g0 = 0.f, g1 = 0.f;
    if (ch_ok) {
      for (int q = xl; q < Q; q += XL) {
        const int s = q / p.Wo, x = q - s * p.Wo;
        const uint32_t* src = tile + (static_cast<size_t>(s * TH * S) * p.Wp + x * S) * P + pr;
        float acc[TH][2];
#pragma unroll
        for (int tt = 0; tt < TH; ++tt) acc[tt][0] = acc[tt][1] = 0.f;
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          uint32_t xin[KS];
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) xin[kx] = src[kx * D * P];
          src += row_words;
#pragma unroll
          for (int ky = 0; ky < KS; ++ky) {
            const int d = r - ky * D;  // compile-time after unrolling
            if (d >= 0 && d % S == 0 && d / S < TH) {
#pragma unroll
              for (int kx = 0; kx < KS; ++kx) fhfma2(acc[d / S][0], acc[d / S][1], xin[kx], wreg[ky * KS + kx]);
            }
          }
        }
        const int oy_s = oy0 + s * TH;
        bf16* dst = p.out + ((static_cast<size_t>(n) * p.Ho + oy_s) * p.Wo + x) * p.C + c0;
        const size_t out_row = static_cast<size_t>(p.Wo) * p.C;
        const bool whole = oy_s + TH <= p.Ho;  // uniform per strip: the common case stores without per-row predicates
        auto finish = [&](auto actf) {
          auto rows = [&](auto full) {
#pragma unroll
            for (int tt = 0; tt < TH; ++tt) {
              if (decltype(full)::value || oy_s + tt < p.Ho) {
                const uint32_t packed = pack2(actf(fmaf(acc[tt][0], sc0, sh0)), actf(fmaf(acc[tt][1], sc1, sh1)));
                *reinterpret_cast<uint32_t*>(dst + tt * out_row) = packed;
                // pool / statistics of what the next layer reads (the bf16-rounded values), again without unpacking:
                // g += packed * 1.0, q += packed * packed
                if (p.mode) fhfma2(g0, g1, packed, 0x3F803F80u);
                if (p.mode == 2) fhfma2(st_q0, st_q1, packed, packed);
              }
            }
          };
          if (whole) rows(std::true_type{}); else rows(std::false_type{});
        };
        // hardswish as z * sat(z/6 + 1/2) (one FFMA.SAT + one FMUL), the form of the GEMM epilogues (gemm_tc.cu)
        if (p.act == ACT_HSWISH) finish([](float v) { return v * __saturatef(fmaf(v, 1.f / 6.f, 0.5f)); });
        else if (p.act == ACT_RELU) finish([](float v) { return fmaxf(v, 0.f); });
        else if (p.act == ACT_NONE) finish([](float v) { return v; });
        else finish([&](float v) { return apply_act(v, p.act); });
      }
    }
    if (p.mode == 2) { st_s0 += g0; st_s1 += g1; }
    if (p.mode == 1) {
      if (lane_ok) { red[(xl * P + pr) * 2] = g0; red[(xl * P + pr) * 2 + 1] = g1; }
      __syncthreads();
      if (tid < 2 * P) {
        const int c = g * (2 * P) + tid;
        if (c < p.C) {
          float s = 0.f;
          for (int l = 0; l < XL; ++l) s += red[l * 2 * P + tid];  // fixed order -> deterministic
          p.gap[(static_cast<size_t>(n) * p.bands + band) * p.C + c] = s;
        }
      }
    }
    __syncthreads();  // every thread is done with this stage (and with `red`)
    if (tid == 0 && t + p.stages * step < p.tiles) issue(t + p.stages * step, stage);
    if (++stage == p.stages) { stage = 0; phase ^= 1; }
  }
  if (p.mode == 2) flush_stats();
}

constexpr size_t kDynBudget = 106 * 1024;  // dynamic shared memory per CTA at two CTAs per SM
constexpr int kPairSet[] = {8, 12, 16, 20, 24, 32};
// MTGSEG_DWCOL_STAGE_KB (A/B): stage size the planner may fill (default 52: two stages, two CTAs per SM; smaller = more stages)
size_t stage_budget() {
  static int kb = -1;
  if (kb < 0) {
    const char* e = getenv("MTGSEG_DWCOL_STAGE_KB");
    kb = e ? atoi(e) : 52;
    if (kb < 8 || kb > 52) kb = 52;
  }
  return static_cast<size_t>(kb) * 1024;
}

int col_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}
// MTGSEG_DWCOL_STAGES (A/B): caps the shared-memory stages per CTA (default: as many as fit, 2..4)
int col_stage_pin() { static int v = env_int("MTGSEG_DWCOL_STAGES"); return v; }
// MTGSEG_DWCOL_TH (A/B): pins the strip height of the stride-1 kernels to 10 or 20 (default: the planner's choice)
int col_th_pin() { static int v = env_int("MTGSEG_DWCOL_TH"); return v; }

}  // namespace

struct DwColPlan { bool ok; int P, TH, pad, Ho, Wo, band, bands, R, Wp, groups; size_t stage_bytes; };

// Channel pairs per CTA group, strip height and rows per band: the candidate with the best product of lane use, column balance,
// (at half weight: halo rows cost L2 bandwidth, not issue slots) useful staged rows and strip efficiency (a strip re-reads
// (k-1)*dil halo rows from shared memory and pays its fixed costs once), under the stage budget.
DwColPlan dw_col_plan(int H, int W, int C, int k, int stride, int dil, bool need_gap) {
  DwColPlan best{};
  best.ok = false;
  if (C % 8 != 0 || C < 8) return best;
  const int pad = (k - 1) / 2 * dil;
  const int Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  if (Ho < 1 || Wo < 1) return best;
  const int Wp = (Wo - 1) * stride + (k - 1) * dil + 1;
  if (Wp > 256) return best;  // TMA box limit
  const int pairs = C / 2;
  const size_t budget = stage_budget();
  double best_score = -1.0;
  for (int TH : {10, 20}) {
    if (stride == 2 && TH == 20) continue;
    const int th = stride == 2 ? 5 : TH;
    if (stride == 1 && col_th_pin() && th != col_th_pin()) continue;
    for (int P : kPairSet) {
      const int groups = ceil_div(pairs, P);
      auto rows = [&](int band) { return (band - 1) * stride + (k - 1) * dil + 1; };
      auto bytes = [&](int band) { return static_cast<size_t>(rows(band)) * Wp * P * 4; };
      if (bytes(th) > budget || rows(th) > 256) continue;
      const int ho_up = ceil_div(Ho, th) * th;
      int band = th;
      while (band + th <= ho_up && bytes(band + th) <= budget && rows(band + th) <= 256) band += th;
      int bands = ceil_div(Ho, band);
      band = ceil_div(ceil_div(Ho, bands), th) * th;  // even bands, whole strips
      bands = ceil_div(Ho, band);
      if (need_gap && bands > 16) continue;
      const int XL = 256 / P;
      const int Q = band / th * Wo;
      const double lane = static_cast<double>(pairs) / (groups * P) * (XL * P) / 256.0;
      const double col = static_cast<double>(Q) / (ceil_div(Q, XL) * XL);
      const double halo = static_cast<double>(band * stride) / rows(band);
      const double box = P >= 16 ? 1.0 : (P == 12 ? 0.97 : 0.94);  // narrow TMA box rows (48 / 32 bytes) cost request rate
      const double rows_used = static_cast<double>(Ho) / (bands * band);  // strips hanging over the last row compute unused outputs
      const double strip = static_cast<double>(th * stride) / ((th - 1) * stride + (k - 1) * dil + 1 + 2);  // + ~2 rows of fixed cost
      const double score = lane * col * (0.5 + 0.5 * halo) * box * rows_used * (0.6 + 0.4 * strip);
      if (score > best_score) {
        best_score = score;
        best = DwColPlan{true, P, th, pad, Ho, Wo, band, bands, rows(band), Wp, groups, align_up(bytes(band), 128)};
      }
    }
  }
  return best;
}

namespace {

template <int KS, int S, int D, int P, int TH>
int launch_col2(const DwConvArgs& a, const DwColPlan& q, cudaStream_t st) {
  CUtensorMap tmx{};
  // input [B][H][W][C] bf16 as a 4-D tensor, box = (group channels, padded row, band rows + halo, 1 image); out-of-bounds
  // rows / columns / channels are zero filled by the TMA unit = the convolution padding.  No swizzle: read as stored.
  const unsigned long long xd[4] = {static_cast<unsigned long long>(a.C), static_cast<unsigned long long>(a.W),
                                    static_cast<unsigned long long>(a.H), static_cast<unsigned long long>(a.B)};
  const unsigned long long xs[3] = {xd[0] * 2, xd[0] * xd[1] * 2, xd[0] * xd[1] * xd[2] * 2};
  const unsigned xb[4] = {static_cast<unsigned>(2 * P), static_cast<unsigned>(q.Wp), static_cast<unsigned>(q.R), 1u};
  int rc = make_tma_map_bf16(&tmx, a.in, 4, xd, xs, xb, 0);
  if (rc != MTG_OK) return rc;
  DwC p{};
  p.w = a.w; p.out = a.out; p.scale = a.scale; p.shift = a.shift; p.gap = a.gap_partial; p.stat = a.stat;
  p.mode = a.stat ? 2 : (a.gap_partial ? 1 : 0);
  p.act = a.act; p.C = a.C; p.Ho = q.Ho; p.Wo = q.Wo; p.pad = q.pad; p.band = q.band; p.bands = q.bands; p.R = q.R; p.Wp = q.Wp;
  p.tiles = a.B * q.bands;
  p.stage_words = static_cast<int>(q.stage_bytes / 4);
  const int slots = 2 * col_sms();
  int gx = slots / q.groups;
  if (gx < 1) gx = 1;
  if (gx > p.tiles) gx = p.tiles;
  gx = ceil_div(p.tiles, ceil_div(p.tiles, gx));  // same tiles per CTA, fewer CTAs
  const dim3 grid(gx, q.groups);
  // as many stages as fit (2..4) (measured: more than two change nothing, the kernel is issue bound, not latency bound)
  int stages = static_cast<int>(kDynBudget / q.stage_bytes);
  stages = stages < 2 ? 2 : (stages > kMaxStages ? kMaxStages : stages);
  if (col_stage_pin() >= 2 && col_stage_pin() <= stages) stages = col_stage_pin();
  p.stages = stages;
  const size_t smem = stages * q.stage_bytes;
  static bool configured = false;  // per instantiation
  if (!configured) {
    MTG_CUDA(cudaFuncSetAttribute(dw_col_kernel<KS, S, D, P, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDynBudget)));
    configured = true;
  }
  MTG_CUDA(launch_pdl(dw_col_kernel<KS, S, D, P, TH>, grid, dim3(256), smem, st, tmx, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

template <int KS, int S, int D, int TH>
int launch_col1(const DwConvArgs& a, const DwColPlan& q, cudaStream_t st) {
  switch (q.P) {
    case 8: return launch_col2<KS, S, D, 8, TH>(a, q, st);
    case 12: return launch_col2<KS, S, D, 12, TH>(a, q, st);
    case 16: return launch_col2<KS, S, D, 16, TH>(a, q, st);
    case 20: return launch_col2<KS, S, D, 20, TH>(a, q, st);
    case 24: return launch_col2<KS, S, D, 24, TH>(a, q, st);
    case 32: return launch_col2<KS, S, D, 32, TH>(a, q, st);
    default: break;
  }
  MTG_REQUIRE(false, MTG_ERR_UNSUPPORTED, "dwconv: no column-strip instantiation for %d channel pairs per group", q.P);
}

template <int KS, int S, int D>
int launch_col0(const DwConvArgs& a, const DwColPlan& q, cudaStream_t st) {
  if constexpr (S == 2) return launch_col1<KS, S, D, 5>(a, q, st);
  else return q.TH == 20 ? launch_col1<KS, S, D, 20>(a, q, st) : launch_col1<KS, S, D, 10>(a, q, st);
}

}  // namespace

int launch_dwconv_col(const DwConvArgs& a, const DwColPlan& q, cudaStream_t st) {
  MTG_REQUIRE(q.ok, MTG_ERR_UNSUPPORTED, "dwconv: feature map %dx%d (C=%d, k=%d) does not fit the column-strip tiling", a.H, a.W, a.C, a.k);
  MTG_REQUIRE(!(a.stat && a.gap_partial), MTG_ERR_ARG, "dwconv: statistics and pool partials are exclusive");
  MTG_REQUIRE(!a.gap_partial || a.chunks == q.bands, MTG_ERR_ARG, "dwconv: gap_partial must have mtgseg_dwconv_chunks() = %d chunks, got %d", q.bands, a.chunks);
  switch (a.k * 100 + a.stride * 10 + a.dil) {
    case 311: return launch_col0<3, 1, 1>(a, q, st);
    case 321: return launch_col0<3, 2, 1>(a, q, st);
    case 511: return launch_col0<5, 1, 1>(a, q, st);
    case 521: return launch_col0<5, 2, 1>(a, q, st);
    case 512: return launch_col0<5, 1, 2>(a, q, st);
    default: break;
  }
  MTG_REQUIRE(false, MTG_ERR_UNSUPPORTED, "dwconv: (k=%d, stride=%d, dilation=%d) is not one of the MobileNetV3 shapes", a.k, a.stride, a.dil);
}

}  // namespace mtgseg
