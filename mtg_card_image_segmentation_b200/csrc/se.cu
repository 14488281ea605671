// Pooled-feature micro kernels: global average pool (channel sums) and the per-image MLP used by the
// squeeze-excite blocks (tv:ops/misc.py:252-261: FC+bias -> ReLU -> FC+bias -> Hardsigmoid) and by the
// head's scale branch (train/model.py:115-119: GAP -> 1x1 conv (no bias) -> Sigmoid).
// These are latency-bound (a few hundred kFLOP per image); the design goal is few launches and no
// re-read of the big activation: the depthwise kernel already produced the channel sums.
#include <stdlib.h>

#include "ops.h"

namespace mtgseg {
namespace {

// in [B][HW][C] bf16 -> out [B][C] fp32 sums.  grid (ceil(CV/8), B); block 256 = 8 vectors x 32 pixel lanes.
__global__ void __launch_bounds__(256) gap_kernel(const bf16* __restrict__ in, float* __restrict__ out, int HW, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32][64 + 1];
  const int CV = C / 8;
  const int vl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int v = blockIdx.x * 8 + vl;
  const int n = blockIdx.y;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (v < CV) {
    const bf16* base = in + static_cast<size_t>(n) * HW * C + v * 8;
    for (int p = pl; p < HW; p += 32) {
      float f[8];
      unpack8(ldg16(base + static_cast<size_t>(p) * C), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[pl][vl * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < C) {
      float s = 0.f;
      for (int l = 0; l < 32; ++l) s += red[l][threadIdx.x];
      out[static_cast<size_t>(n) * C + c] = s;
    }
  }
}

constexpr int FC_IMGS = 8;   // images per CTA: each weight element is loaded once per 8 images
constexpr int FC_OUTS = 32;  // outputs per CTA (4 per warp)

struct FcP {
  const float* in; int chunks; float in_scale;  // in [B][chunks][C] (summed over chunks, times in_scale)
  int B, C, O;
  const bf16* w; const float* bias; int act;
  float* out;  // [B][O]
};

// out[n][o] = act( sum_c w[o][c] * (in_scale * sum_k in[n][k][c]) + bias[o] )
// grid (ceil(O/32), ceil(B/8)); each warp owns 4 outputs x 8 images, lanes stride over C in 8-wide vectors.
__global__ void __launch_bounds__(256) fc_batched_kernel(const FcP p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float xin[];  // [FC_IMGS][C]
  const int n0 = blockIdx.y * FC_IMGS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < FC_IMGS * p.C; idx += blockDim.x) {
    const int i = idx / p.C, c = idx - i * p.C;
    float s = 0.f;
    if (n0 + i < p.B)
      for (int k = 0; k < p.chunks; ++k) s += p.in[(static_cast<size_t>(n0 + i) * p.chunks + k) * p.C + c];
    xin[idx] = s * p.in_scale;
  }
  __syncthreads();
  const int o0 = blockIdx.x * FC_OUTS + warp * 4;
  if (o0 >= p.O) return;
  float acc[4][FC_IMGS];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < FC_IMGS; ++i) acc[a][i] = 0.f;
  for (int c = lane * 8; c < p.C; c += 256) {
    float wf[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (o0 + a < p.O) unpack8(ldg16(p.w + static_cast<size_t>(o0 + a) * p.C + c), wf[a]);
      else
#pragma unroll
        for (int e = 0; e < 8; ++e) wf[a][e] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < FC_IMGS; ++i) {
      const float4 x0 = *reinterpret_cast<const float4*>(xin + i * p.C + c);
      const float4 x1 = *reinterpret_cast<const float4*>(xin + i * p.C + c + 4);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[a][i] = fmaf(wf[a][e], xv[e], acc[a][i]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < FC_IMGS; ++i)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[a][i] += __shfl_xor_sync(0xffffffffu, acc[a][i], o);
  // lane (a*8 + i) writes output a of image i
  const int a_sel = lane >> 3, i_sel = lane & 7;
  float v = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < FC_IMGS; ++i)
      if (a == a_sel && i == i_sel) v = acc[a][i];
  if (o0 + a_sel < p.O && n0 + i_sel < p.B) {
    const float b = p.bias ? p.bias[o0 + a_sel] : 0.f;
    p.out[static_cast<size_t>(n0 + i_sel) * p.O + o0 + a_sel] = apply_act(v + b, p.act);
  }
}


// ---------------------------------------------------------------------------------------------------------
// Fused squeeze-excite MLP: one launch per SE block.  A CTA owns SE_IMGS images: it sums the depthwise kernel's pool
// partials into shared memory, computes hidden = act1(W1 mean + b1) into shared memory and gate = act2(W2 hidden + b2)
// into global memory.  Warp-per-output-group with coalesced 16-byte weight loads, as in fc_batched_kernel; the
// weights (<= 460 KB per layer) are served by L2 to the B/SE_IMGS CTAs.
// ---------------------------------------------------------------------------------------------------------
constexpr int SE_IMGS = 2;   // images per CTA: B/2 CTAs cover the whole GPU at B = 256
constexpr int SE_OUTS = 8;   // outputs per warp pass: 8 independent 16-byte weight loads in flight per lane
constexpr int SE_THREADS = 512;
static_assert(SE_IMGS * SE_OUTS == 16, "the transposing reduction below is written for 16 values per lane");

struct SeP {
  const float* sums; int chunks; float in_scale;  // [B][chunks][C]
  int B, C, SQ;
  const bf16* w1; const float* b1; int act1;  // [SQ][C]
  const bf16* w2; const float* b2; int act2;  // [C][SQ] or null (single-layer form: out is [B][SQ])
  float* out; float* hidden;                  // hidden [B][SQ] optional copy for the training backward
};

// y[i][o] = act(sum_c w[o][c] * x[i][c] + bias[o]) for the CTA's SE_IMGS images; x in shared memory [SE_IMGS][I].
// A warp owns SE_OUTS outputs per pass, lanes stride over the input dimension in 8-wide vectors.  Products are
// accumulated as packed fp32x2 (even / odd input channel), and the 16 per-lane partial sums are reduced over the warp
// with a transposing butterfly: 8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5 (measured: shuffles were 35 % of the
// instructions of the first version of this kernel).
template <typename Store>
__device__ __forceinline__ void se_fc(const float* x, int I, const bf16* __restrict__ w, const float* __restrict__ bias, int O, int act,
                                      Store store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = SE_THREADS / 32;
  for (int o0 = warp * SE_OUTS; o0 < O; o0 += nwarps * SE_OUTS) {
    uint64_t acc[SE_OUTS][SE_IMGS];
#pragma unroll
    for (int a = 0; a < SE_OUTS; ++a)
#pragma unroll
      for (int i = 0; i < SE_IMGS; ++i) acc[a][i] = 0ull;
    for (int c = lane * 8; c < I; c += 256) {
      uint4 wq[SE_OUTS];
#pragma unroll
      for (int a = 0; a < SE_OUTS; ++a) wq[a] = o0 + a < O ? ldg16(w + static_cast<size_t>(o0 + a) * I + c) : make_uint4(0u, 0u, 0u, 0u);
      uint64_t xx[SE_IMGS][4];
#pragma unroll
      for (int i = 0; i < SE_IMGS; ++i) {
        const float4 x0 = *reinterpret_cast<const float4*>(x + i * I + c);
        const float4 x1 = *reinterpret_cast<const float4*>(x + i * I + c + 4);
        asm("mov.b64 %0, {%1,%2};" : "=l"(xx[i][0]) : "f"(x0.x), "f"(x0.y));
        asm("mov.b64 %0, {%1,%2};" : "=l"(xx[i][1]) : "f"(x0.z), "f"(x0.w));
        asm("mov.b64 %0, {%1,%2};" : "=l"(xx[i][2]) : "f"(x1.x), "f"(x1.y));
        asm("mov.b64 %0, {%1,%2};" : "=l"(xx[i][3]) : "f"(x1.z), "f"(x1.w));
      }
#pragma unroll
      for (int a = 0; a < SE_OUTS; ++a) {
        const uint32_t ww[4] = {wq[a].x, wq[a].y, wq[a].z, wq[a].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint64_t wp;  // (even channel, odd channel) of one bf16x2 word as fp32x2
          asm("mov.b64 %0, {%1,%2};" : "=l"(wp) : "r"(ww[h] << 16), "r"(ww[h] & 0xFFFF0000u));
#pragma unroll
          for (int i = 0; i < SE_IMGS; ++i) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[a][i]) : "l"(wp), "l"(xx[i][h]));
        }
      }
    }
    float v[16];  // value index a * SE_IMGS + i
#pragma unroll
    for (int a = 0; a < SE_OUTS; ++a)
#pragma unroll
      for (int i = 0; i < SE_IMGS; ++i) {
        float lo, hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[a][i]));
        v[a * SE_IMGS + i] = lo + hi;
      }
    // transposing butterfly: after the step with offset d a lane keeps half of its values; lane L ends with value L >> 1
#pragma unroll
    for (int n = 8, d = 16; n >= 1; n >>= 1, d >>= 1) {
      const bool up = (lane & d) != 0;
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const float keep = up ? v[j + n] : v[j];
        const float send = up ? v[j] : v[j + n];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    const int idx = lane >> 1, a_sel = idx / SE_IMGS, i_sel = idx % SE_IMGS;
    if ((lane & 1) == 0 && o0 + a_sel < O) store(i_sel, o0 + a_sel, apply_act(v[0] + (bias ? bias[o0 + a_sel] : 0.f), act));
  }
}

__global__ void __launch_bounds__(SE_THREADS) se_fused_kernel(const SeP p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float se_smem[];
  float* xin = se_smem;                   // [SE_IMGS][C]
  float* hid = se_smem + SE_IMGS * p.C;   // [SE_IMGS][SQ]
  const int n0 = blockIdx.x * SE_IMGS;
  for (int i = 0; i < SE_IMGS; ++i) {
    const bool ok = n0 + i < p.B;
    const float* src = p.sums + static_cast<size_t>(ok ? n0 + i : 0) * p.chunks * p.C;
    for (int c = threadIdx.x; c < p.C; c += SE_THREADS) {
      float s = 0.f;
      if (ok) {
        int k = 0;
        for (; k + 4 <= p.chunks; k += 4) {  // four independent loads in flight, summed in the original order
          const float a0 = src[k * p.C + c], a1 = src[(k + 1) * p.C + c], a2 = src[(k + 2) * p.C + c], a3 = src[(k + 3) * p.C + c];
          s += a0; s += a1; s += a2; s += a3;
        }
        for (; k < p.chunks; ++k) s += src[k * p.C + c];
      }
      xin[i * p.C + c] = s * p.in_scale;
    }
  }
  __syncthreads();
  if (p.w2) {
    se_fc(xin, p.C, p.w1, p.b1, p.SQ, p.act1, [&](int i, int o, float v) {
      hid[i * p.SQ + o] = v;
      if (p.hidden && n0 + i < p.B) p.hidden[static_cast<size_t>(n0 + i) * p.SQ + o] = v;
    });
    __syncthreads();
    se_fc(hid, p.SQ, p.w2, p.b2, p.C, p.act2, [&](int i, int o, float v) {
      if (n0 + i < p.B) p.out[static_cast<size_t>(n0 + i) * p.C + o] = v;
    });
  } else {
    se_fc(xin, p.C, p.w1, p.b1, p.SQ, p.act1, [&](int i, int o, float v) {
      if (n0 + i < p.B) p.out[static_cast<size_t>(n0 + i) * p.SQ + o] = v;
    });
  }
}

}  // namespace

int launch_gap(const bf16* in, float* out, int B, int HW, int C, cudaStream_t st) {
  MTG_REQUIRE(in && out && C % 8 == 0, MTG_ERR_ARG, "gap: bad arguments");
  dim3 grid(ceil_div(C / 8, 8), B);
  MTG_CUDA(launch_pdl(gap_kernel, dim3(grid), dim3(256), 0, st, in, out, HW, C));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_se_mlp(const SeMlpArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.sums && a.w1 && a.out && a.HW > 0, MTG_ERR_ARG, "se_mlp: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && a.SQ % 8 == 0, MTG_ERR_UNSUPPORTED, "se_mlp: C=%d SQ=%d must be multiples of 8", a.C, a.SQ);
  MTG_REQUIRE(!a.w2 || a.hidden, MTG_ERR_ARG, "se_mlp: the two-layer form needs a hidden scratch buffer [B][SQ]");
  const int cmax = a.C > a.SQ ? a.C : a.SQ;
  MTG_REQUIRE(static_cast<size_t>(FC_IMGS) * cmax * sizeof(float) <= 48 * 1024, MTG_ERR_UNSUPPORTED, "se_mlp: C too large");
  static const bool two_launch = getenv("MTGSEG_SE_VARIANT") && atoi(getenv("MTGSEG_SE_VARIANT")) == 1;  // A/B switch
  const size_t fused_smem = static_cast<size_t>(SE_IMGS) * (a.C + a.SQ) * sizeof(float);
  if (!two_launch && fused_smem <= 48 * 1024) {
    SeP sp{a.sums, a.chunks, 1.0f / static_cast<float>(a.HW), a.B, a.C, a.SQ, a.w1, a.b1, a.act1, a.w2, a.b2, a.act2, a.out, a.w2 ? a.hidden : nullptr};
    MTG_CUDA(launch_pdl(se_fused_kernel, dim3(ceil_div(a.B, SE_IMGS)), dim3(SE_THREADS), fused_smem, st, sp));
    MTG_LAUNCH_CHECK();
    return MTG_OK;
  }
  FcP l1{a.sums, a.chunks, 1.0f / static_cast<float>(a.HW), a.B, a.C, a.SQ, a.w1, a.b1, a.act1, a.w2 ? a.hidden : a.out};
  dim3 g1(ceil_div(a.SQ, FC_OUTS), ceil_div(a.B, FC_IMGS));
  MTG_CUDA(launch_pdl(fc_batched_kernel, dim3(g1), dim3(256), static_cast<size_t>(FC_IMGS) * a.C * sizeof(float), st, l1));
  MTG_LAUNCH_CHECK();
  if (a.w2) {
    FcP l2{a.hidden, 1, 1.0f, a.B, a.SQ, a.C, a.w2, a.b2, a.act2, a.out};
    dim3 g2(ceil_div(a.C, FC_OUTS), ceil_div(a.B, FC_IMGS));
    MTG_CUDA(launch_pdl(fc_batched_kernel, dim3(g2), dim3(256), static_cast<size_t>(FC_IMGS) * a.SQ * sizeof(float), st, l2));
    MTG_LAUNCH_CHECK();
  }
  return MTG_OK;
}

}  // namespace mtgseg
