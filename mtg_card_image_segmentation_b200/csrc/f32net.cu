// fp32-exact inference path: the network of train/model.py as train/evaluate.py:66 runs it -- `self.model(images)` in
// float32 with NO autocast -- for callers that hold the output to the reference's own fp32 tolerance (1e-4, the bar
// train/export.py:159 puts on its exported graph).
//
// Everything stays IEEE fp32 with round-to-nearest FMA accumulation on the CUDA cores: activations are NHWC fp32 in the
// workspace, weights are read straight from the fp32 OIHW master parameters (no packed copy, nothing rounded to bf16), the
// eval-mode BatchNorm is applied in the epilogue from the running statistics.  The tensor-core path (gemm_tc.cu) stores
// bf16 and accumulates in the tcgen05 datapath; split-bf16 / 3xTF32 emulation on that datapath inherits its truncating
// accumulation, whose error grows with K (8640 for the head's 3x3), so the accuracy mode does not use it.
//
//   f32_conv_kernel<TAPS>   1x1 (TAPS=1) and 3x3 pad-1 (TAPS=9) convolutions as a register-tiled SGEMM: 128x64 tile,
//                           8-deep k slices double-buffered through shared memory, 8x4 outputs per thread; squeeze-excite
//                           gate applied while the A slice is loaded; BN + activation + residual in the epilogue
//   f32_dw_kernel           depthwise k x k (stride 1/2, dilation 1/2), one thread = 4 channels of one output pixel
//   f32_stem_kernel         3x3 stride-2 stem on the NCHW input
//   f32_pool / f32_mlp      global average pool and the squeeze-excite / head-scale MLPs
//   f32_head_cls / f32_head_low   the head's classifiers + x2 bilinear (train/model.py:137-142), then tail.cu's
//                           upsample_out (already fp32) writes logits / argmax mask / confusion counts
#include "net.h"

namespace mtgseg {

namespace {

struct F32Bn {  // eval-mode BatchNorm of one layer, straight from the state_dict tensors (nullptr gamma: identity)
  const float* gamma; const float* beta; const float* mean; const float* var; float eps;
};

__device__ __forceinline__ void bn_coeff(const F32Bn& bn, int c, float& sc, float& sh) {
  if (bn.gamma == nullptr) { sc = 1.f; sh = 0.f; return; }
  // same operation order as ATen's batch_norm in eval mode: invstd = 1/sqrt(var + eps); y = (x - mean) * invstd * w + b
  const float invstd = 1.f / sqrtf(__ldg(bn.var + c) + bn.eps);
  sc = invstd * __ldg(bn.gamma + c);
  sh = __ldg(bn.beta + c) - __ldg(bn.mean + c) * sc;
}

__device__ __forceinline__ float act_f32(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_HSWISH: return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) / 6.f;
    case ACT_HSIGMOID: return fminf(fmaxf(x + 3.f, 0.f), 6.f) / 6.f;
    case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}

// ---------------------------------------------------------------------------------------------------------
// convolution as SGEMM
// ---------------------------------------------------------------------------------------------------------
constexpr int FBM = 128, FBN = 64, FBK = 8;

struct F32ConvP {
  const float* a;   // NHWC activations [M][K]
  const float* w;   // TAPS == 1: OIHW [N][K] ; TAPS == 9: [N][9][K] (f32_pack3x3_kernel)
  float* out;       // [M][N]
  int M, N, K, H, W;
  F32Bn bn; int act;
  const float* residual;  // [M][N] or nullptr
  const float* gate; int hw;  // squeeze-excite multiplier [B][K] on the A rows of image m / hw, or nullptr
};

template <int TAPS>
__global__ void __launch_bounds__(256) f32_conv_kernel(const F32ConvP p) {
  __shared__ __align__(16) float As[2][FBK][FBM + 4];
  __shared__ __align__(16) float Bs[2][FBK][FBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * FBM, n0 = blockIdx.y * FBN;
  // loader roles: A slice = 128 rows x 8 k -> one float4 per thread; B slice = 64 rows x 8 k -> threads 0..127
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const int am = m0 + lrow;
  const bool a_ok = am < p.M;
  int ab = 0, ay = 0, ax = 0;
  if (TAPS == 9 && a_ok) {
    ab = am / (p.H * p.W);
    const int r = am - ab * p.H * p.W;
    ay = r / p.W; ax = r - ay * p.W;
  }
  const float* gate_row = (p.gate && a_ok) ? p.gate + static_cast<size_t>(am / p.hw) * p.K : nullptr;
  const int bn_row = n0 + lrow;
  const bool b_ok = tid < 128 && bn_row < p.N;
  const int kslices = p.K / FBK;  // K % 8 == 0 (checked by the launcher)
  const int total = kslices * TAPS;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  auto fetch = [&](int it, float4& av, float4& bv) {
    const int tap = TAPS == 9 ? it / kslices : 0;
    const int k0 = (it - tap * kslices) * FBK + lk;
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_ok) {
      if (TAPS == 9) {
        const int yy = ay + tap / 3 - 1, xx = ax + tap % 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
          av = __ldg(reinterpret_cast<const float4*>(p.a + (static_cast<size_t>(ab * p.H + yy) * p.W + xx) * p.K + k0));
      } else {
        av = __ldg(reinterpret_cast<const float4*>(p.a + static_cast<size_t>(am) * p.K + k0));
        if (gate_row) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(gate_row + k0));
          av.x *= g.x; av.y *= g.y; av.z *= g.z; av.w *= g.w;
        }
      }
    }
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_ok) bv = __ldg(reinterpret_cast<const float4*>(p.w + (static_cast<size_t>(bn_row) * TAPS + tap) * p.K + k0));
  };
  auto stash = [&](int buf, const float4& av, const float4& bv) {
    As[buf][lk + 0][lrow] = av.x; As[buf][lk + 1][lrow] = av.y; As[buf][lk + 2][lrow] = av.z; As[buf][lk + 3][lrow] = av.w;
    if (tid < 128) {
      Bs[buf][lk + 0][lrow] = bv.x; Bs[buf][lk + 1][lrow] = bv.y; Bs[buf][lk + 2][lrow] = bv.z; Bs[buf][lk + 3][lrow] = bv.w;
    }
  };

  float4 av, bv;
  fetch(0, av, bv);
  stash(0, av, bv);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) fetch(it + 1, av, bv);
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < total) stash(buf ^ 1, av, bv);
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= p.N) return;  // N % 4 == 0
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bn_coeff(p.bn, n + j, sc[j], sh[j]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) break;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = act_f32(fmaf(acc[i][j], sc[j], sh[j]), p.act);
    const size_t off = static_cast<size_t>(m) * p.N + n;
    if (p.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + off));
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    *reinterpret_cast<float4*>(p.out + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// OIHW [N][K][3][3] -> [N][9][K]
__global__ void f32_pack3x3_kernel(const float* __restrict__ w, float* __restrict__ out, int N, int K) {
  const size_t total = static_cast<size_t>(N) * K * 9;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const size_t r = i / K;
    const int t = static_cast<int>(r % 9);
    const size_t n = r / 9;
    out[i] = w[(n * K + k) * 9 + t];
  }
}

// ---------------------------------------------------------------------------------------------------------
// stem, depthwise
// ---------------------------------------------------------------------------------------------------------
// x NCHW [B][3][H][W] -> out NHWC [B][Ho][Wo][16]; w OIHW [16][3][3][3]; one thread = one output pixel, 16 channels
__global__ void __launch_bounds__(128) f32_stem_kernel(const float* __restrict__ x, const float* __restrict__ w, F32Bn bn,
                                                       float* __restrict__ out, int B, int H, int W, int Ho, int Wo) {
  __shared__ float sw[27][16];
  __shared__ float ssc[16], ssh[16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) {
    const int co = i / 27, r = i - co * 27;  // OIHW: w[co][ci][ky][kx], r = ci*9 + ky*3 + kx
    sw[r][co] = w[i];
  }
  if (threadIdx.x < 16) bn_coeff(bn, threadIdx.x, ssc[threadIdx.x], ssh[threadIdx.x]);
  __syncthreads();
  const long long total = static_cast<long long>(B) * Ho * Wo;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(idx % Wo);
    const long long t = idx / Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    const float* xn = x + static_cast<size_t>(n) * 3 * H * W;
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + kx;
          if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
          const float v = __ldg(xn + (static_cast<size_t>(ci) * H + iy) * W + ix);
          const float* wr = sw[ci * 9 + ky * 3 + kx];
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
        }
    float* o = out + static_cast<size_t>(idx) * 16;
#pragma unroll
    for (int c = 0; c < 16; c += 4)
      *reinterpret_cast<float4*>(o + c) = make_float4(act_f32(fmaf(acc[c], ssc[c], ssh[c]), ACT_HSWISH),
                                                      act_f32(fmaf(acc[c + 1], ssc[c + 1], ssh[c + 1]), ACT_HSWISH),
                                                      act_f32(fmaf(acc[c + 2], ssc[c + 2], ssh[c + 2]), ACT_HSWISH),
                                                      act_f32(fmaf(acc[c + 3], ssc[c + 3], ssh[c + 3]), ACT_HSWISH));
  }
}

struct F32DwP {
  const float* in; const float* w; float* out;  // w: OIHW depthwise [C][1][k][k]
  int B, H, W, C, Ho, Wo, k, stride, dil, pad;
  F32Bn bn; int act;
};
__global__ void __launch_bounds__(256) f32_dw_kernel(const F32DwP p) {
  const int C4 = p.C >> 2;
  const long long total = static_cast<long long>(p.B) * p.Ho * p.Wo * C4;
  const int kk = p.k * p.k;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C4) * 4;
    long long t = idx / C4;
    const int ox = static_cast<int>(t % p.Wo); t /= p.Wo;
    const int oy = static_cast<int>(t % p.Ho);
    const int n = static_cast<int>(t / p.Ho);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* in_n = p.in + static_cast<size_t>(n) * p.H * p.W * p.C + c;
    for (int ky = 0; ky < p.k; ++ky) {
      const int iy = oy * p.stride - p.pad + ky * p.dil;
      if (iy < 0 || iy >= p.H) continue;
      for (int kx = 0; kx < p.k; ++kx) {
        const int ix = ox * p.stride - p.pad + kx * p.dil;
        if (ix < 0 || ix >= p.W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(in_n + (static_cast<size_t>(iy) * p.W + ix) * p.C));
        const int wi = ky * p.k + kx;
        acc[0] = fmaf(v.x, __ldg(p.w + (c + 0) * kk + wi), acc[0]);
        acc[1] = fmaf(v.y, __ldg(p.w + (c + 1) * kk + wi), acc[1]);
        acc[2] = fmaf(v.z, __ldg(p.w + (c + 2) * kk + wi), acc[2]);
        acc[3] = fmaf(v.w, __ldg(p.w + (c + 3) * kk + wi), acc[3]);
      }
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sc, sh;
      bn_coeff(p.bn, c + j, sc, sh);
      o[j] = act_f32(fmaf(acc[j], sc, sh), p.act);
    }
    *reinterpret_cast<float4*>(p.out + ((static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox) * p.C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// pooling + small MLPs
// ---------------------------------------------------------------------------------------------------------
// mean[b][c] = (1/HW) sum_pix x[b][pix][c]; grid (ceil(C/32), B), 256 threads = 32 channels x 8 pixel lanes, fixed order
__global__ void __launch_bounds__(256) f32_pool_kernel(const float* __restrict__ x, float* __restrict__ mean, int HW, int C) {
  __shared__ float red[8][32];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, b = blockIdx.y;
  float s = 0.f;
  if (c < C) {
    const float* base = x + static_cast<size_t>(b) * HW * C + c;
    for (int r = pl; r < HW; r += 8) s += __ldg(base + static_cast<size_t>(r) * C);
  }
  red[pl][cl] = s;
  __syncthreads();
  if (pl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cl];
    mean[static_cast<size_t>(b) * C + c] = t / static_cast<float>(HW);
  }
}

// per image: h = act1(W1 mean + b1) [SQ] ; out = act2(W2 h + b2) [C] (out = h when W2 == nullptr).  W1 [SQ][C], W2 [C][SQ] fp32
// (the 1x1 convs fc1 / fc2 of tv:ops/misc.py:225-261, and the bias-free head scale conv of train/model.py:115-119)
__global__ void __launch_bounds__(256) f32_mlp_kernel(const float* __restrict__ mean, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, int act1, const float* __restrict__ w2,
                                                      const float* __restrict__ b2, int act2, float* __restrict__ out, int C, int SQ) {
  extern __shared__ float sm[];  // [C] mean, [SQ] hidden
  float* smean = sm;
  float* shid = sm + C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) smean[c] = mean[static_cast<size_t>(b) * C + c];
  __syncthreads();
  for (int j = warp; j < SQ; j += 8) {
    const float* wr = w1 + static_cast<size_t>(j) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(__ldg(wr + c), smean[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float v = act_f32(s + (b1 ? __ldg(b1 + j) : 0.f), act1);
      shid[j] = v;
      if (!w2) out[static_cast<size_t>(b) * SQ + j] = v;
    }
  }
  if (!w2) return;
  __syncthreads();
  for (int c = warp; c < C; c += 8) {
    const float* wr = w2 + static_cast<size_t>(c) * SQ;
    float s = 0.f;
    for (int j = lane; j < SQ; j += 32) s = fmaf(__ldg(wr + j), shid[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[static_cast<size_t>(b) * C + c] = act_f32(s + (b2 ? __ldg(b2 + c) : 0.f), act2);
  }
}

// ---------------------------------------------------------------------------------------------------------
// head tail: high classifier on (cbr * s) at 20x15 (linear, so it commutes with the x2 bilinear and the bias),
// then low classifier + biases + x2 bilinear at 40x30
// ---------------------------------------------------------------------------------------------------------
// h2[b][pix][c] = sum_i w_high[c][i] * s[b][i] * cbr[b][pix][i]; one warp per (b, pix)
__global__ void __launch_bounds__(256) f32_head_cls_kernel(const float* __restrict__ cbr, const float* __restrict__ s,
                                                           const float* __restrict__ w_high, float* __restrict__ h2, int B, int HWh,
                                                           int IC, int NC) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (wid >= static_cast<long long>(B) * HWh) return;
  const int b = static_cast<int>(wid / HWh);
  const float* row = cbr + static_cast<size_t>(wid) * IC;
  const float* sb = s + static_cast<size_t>(b) * IC;
  for (int c = 0; c < NC; ++c) {
    float a = 0.f;
    for (int i = lane; i < IC; i += 32) a = fmaf(__ldg(row + i) * __ldg(sb + i), __ldg(w_high + c * IC + i), a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) h2[static_cast<size_t>(wid) * NC + c] = a;
  }
}

__device__ __forceinline__ void src_index_f32(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

// lowres[b][y][x][c] = b_low[c] + sum_k w_low[c][k] low[b][y][x][k] + b_high[c] + bilinear(h2)[y][x][c]; one thread per (b, y, x)
__global__ void __launch_bounds__(128) f32_head_low_kernel(const float* __restrict__ h2, const float* __restrict__ low,
                                                           const float* __restrict__ w_low, const float* __restrict__ b_low,
                                                           const float* __restrict__ b_high, float* __restrict__ lowres, int B, int Hh,
                                                           int Wh, int Hl, int Wl, int LC, int NC) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * Hl * Wl) return;
  const int x = static_cast<int>(idx % Wl);
  const long long t = idx / Wl;
  const int y = static_cast<int>(t % Hl), b = static_cast<int>(t / Hl);
  int y0, y1, x0, x1;
  float ly, lx;
  src_index_f32(y, static_cast<float>(Hh) / Hl, Hh, y0, y1, ly);
  src_index_f32(x, static_cast<float>(Wh) / Wl, Wh, x0, x1, lx);
  const float* hb = h2 + static_cast<size_t>(b) * Hh * Wh * NC;
  const float* lrow = low + static_cast<size_t>(idx) * LC;
  for (int c = 0; c < NC; ++c) {
    const float v00 = hb[(y0 * Wh + x0) * NC + c], v01 = hb[(y0 * Wh + x1) * NC + c];
    const float v10 = hb[(y1 * Wh + x0) * NC + c], v11 = hb[(y1 * Wh + x1) * NC + c];
    float hi = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11) + __ldg(b_high + c);
    float lo = __ldg(b_low + c);
    for (int k = 0; k < LC; ++k) lo = fmaf(__ldg(lrow + k), __ldg(w_low + c * LC + k), lo);
    lowres[static_cast<size_t>(idx) * NC + c] = lo + hi;
  }
}

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

inline int conv_out(int in, int k, int stride, int dil) {
  const int pad = (k - 1) / 2 * dil;
  return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1;
}

int launch_f32_conv(const F32ConvP& p, int taps, cudaStream_t st) {
  MTG_REQUIRE(p.K % 8 == 0 && p.N % 4 == 0, MTG_ERR_UNSUPPORTED, "f32 conv: K %% 8 / N %% 4 (K=%d N=%d)", p.K, p.N);
  dim3 grid(ceil_div(p.M, FBM), ceil_div(p.N, FBN));
  if (taps == 9) f32_conv_kernel<9><<<grid, 256, 0, st>>>(p);
  else f32_conv_kernel<1><<<grid, 256, 0, st>>>(p);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

inline int blocks_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return static_cast<int>(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// One fp32-exact inference forward.  ws == nullptr: dry run that only sizes the workspace.
int run_infer_f32(const NetPlan& P, const InferF32IO& io, uint8_t* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st) {
  const bool dry = ws == nullptr;
  Bump bump;
  const int B = io.batch;
  auto buf = [&](size_t elems) { return reinterpret_cast<float*>(ws + bump.take(elems * sizeof(float))); };
  auto prm = [&](int idx) { return static_cast<const float*>(io.params[idx]); };
  auto bn_of = [&](const ConvBnPlan& c) { return F32Bn{prm(c.gamma), prm(c.beta), prm(c.mean), prm(c.var), c.eps}; };
#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)
  int H = conv_out(P.desc.in_h, 3, 2, 1), W = conv_out(P.desc.in_w, 3, 2, 1);
  float* t = buf(static_cast<size_t>(B) * H * W * 16);
  if (!dry) {
    const long long total = static_cast<long long>(B) * H * W;
    f32_stem_kernel<<<blocks_for(total, 128), 128, 0, st>>>(io.x, prm(P.stem.w_idx), bn_of(P.stem), t, B, P.desc.in_h, P.desc.in_w, H, W);
    MTG_LAUNCH_CHECK();
  }
  const float* low = nullptr;
  int Hl = 0, Wl = 0;
  for (int i = 0; i < kNumBlocks; ++i) {
    const BlockPlan& b = P.blocks[i];
    const BlockCfg& c = b.cfg;
    const float* inp = t;
    const float* e = t;
    if (b.has_expand) {
      float* eb = buf(static_cast<size_t>(B) * H * W * c.cexp);
      if (!dry) {
        F32ConvP g{};
        g.a = t; g.w = prm(b.expand.w_idx); g.out = eb; g.M = B * H * W; g.N = c.cexp; g.K = c.cin; g.bn = bn_of(b.expand); g.act = c.act;
        RC(launch_f32_conv(g, 1, st));
      }
      e = eb;
    }
    const int stride = c.dil > 1 ? 1 : c.stride;
    const int Ho = conv_out(H, c.k, stride, c.dil), Wo = conv_out(W, c.k, stride, c.dil);
    float* dwo = buf(static_cast<size_t>(B) * Ho * Wo * c.cexp);
    float* mean = c.se ? buf(static_cast<size_t>(B) * c.cexp) : nullptr;
    float* gate = c.se ? buf(static_cast<size_t>(B) * c.cexp) : nullptr;
    if (!dry) {
      F32DwP d{};
      d.in = e; d.w = prm(b.dw.w_idx); d.out = dwo; d.B = B; d.H = H; d.W = W; d.C = c.cexp; d.Ho = Ho; d.Wo = Wo;
      d.k = c.k; d.stride = stride; d.dil = c.dil; d.pad = (c.k - 1) / 2 * c.dil; d.bn = bn_of(b.dw); d.act = c.act;
      f32_dw_kernel<<<blocks_for(static_cast<long long>(B) * Ho * Wo * (c.cexp / 4), 256), 256, 0, st>>>(d);
      MTG_LAUNCH_CHECK();
      if (c.se) {
        f32_pool_kernel<<<dim3(ceil_div(c.cexp, 32), B), 256, 0, st>>>(dwo, mean, Ho * Wo, c.cexp);
        MTG_LAUNCH_CHECK();
        f32_mlp_kernel<<<B, 256, (c.cexp + b.sq) * sizeof(float), st>>>(mean, prm(b.fc1_w), prm(b.fc1_b), ACT_RELU, prm(b.fc2_w),
                                                                          prm(b.fc2_b), ACT_HSIGMOID, gate, c.cexp, b.sq);
        MTG_LAUNCH_CHECK();
      }
    }
    H = Ho; W = Wo;
    float* o = buf(static_cast<size_t>(B) * H * W * c.cout);
    if (!dry) {
      F32ConvP g{};
      g.a = dwo; g.w = prm(b.project.w_idx); g.out = o; g.M = B * H * W; g.N = c.cout; g.K = c.cexp; g.bn = bn_of(b.project);
      g.act = ACT_NONE; g.residual = (c.stride == 1 && c.cin == c.cout) ? inp : nullptr; g.gate = gate; g.hw = H * W;
      RC(launch_f32_conv(g, 1, st));
    }
    t = o;
    if (i == 3) { low = o; Hl = H; Wl = W; }
  }
  const int ic = P.desc.inter_channels, nc = P.desc.num_classes;
  float* high = buf(static_cast<size_t>(B) * H * W * 960);
  float* cbr = buf(static_cast<size_t>(B) * H * W * ic);
  float* w3 = buf(static_cast<size_t>(ic) * 960 * 9);
  float* hmean = buf(static_cast<size_t>(B) * 960);
  float* hscale = buf(static_cast<size_t>(B) * ic);
  float* h2 = buf(static_cast<size_t>(B) * H * W * nc);
  float* lowres = buf(static_cast<size_t>(B) * Hl * Wl * nc);
  if (ws_needed) *ws_needed = bump.off;
  if (dry) return MTG_OK;
  MTG_REQUIRE(bump.off <= ws_bytes, MTG_ERR_WORKSPACE, "forward_infer_f32: workspace too small: need %zu bytes, got %zu", bump.off, ws_bytes);
  {
    F32ConvP g{};
    g.a = t; g.w = prm(P.last.w_idx); g.out = high; g.M = B * H * W; g.N = 960; g.K = 160; g.bn = bn_of(P.last); g.act = ACT_HSWISH;
    RC(launch_f32_conv(g, 1, st));
    f32_pack3x3_kernel<<<blocks_for(static_cast<long long>(ic) * 960 * 9, 256), 256, 0, st>>>(prm(P.cbr.w_idx), w3, ic, 960);
    MTG_LAUNCH_CHECK();
    F32ConvP h{};
    h.a = high; h.w = w3; h.out = cbr; h.M = B * H * W; h.N = ic; h.K = 960; h.H = H; h.W = W; h.bn = bn_of(P.cbr); h.act = ACT_RELU;
    RC(launch_f32_conv(h, 9, st));
    f32_pool_kernel<<<dim3(ceil_div(960, 32), B), 256, 0, st>>>(high, hmean, H * W, 960);
    MTG_LAUNCH_CHECK();
    f32_mlp_kernel<<<B, 256, (960 + ic) * sizeof(float), st>>>(hmean, prm(P.scale_w), nullptr, ACT_SIGMOID, nullptr, nullptr, ACT_NONE,
                                                                hscale, 960, ic);
    MTG_LAUNCH_CHECK();
    f32_head_cls_kernel<<<ceil_div(B * H * W * 32, 256), 256, 0, st>>>(cbr, hscale, prm(P.high_w), h2, B, H * W, ic, nc);
    MTG_LAUNCH_CHECK();
    f32_head_low_kernel<<<ceil_div(B * Hl * Wl, 128), 128, 0, st>>>(h2, low, prm(P.low_w), prm(P.low_b), prm(P.high_b), lowres, B, H, W,
                                                                     Hl, Wl, 40, nc);
    MTG_LAUNCH_CHECK();
    UpsampleOutArgs u;
    u.lowres = lowres; u.logits = io.logits; u.logits_dtype = io.logits_dtype; u.mask = io.mask; u.targets = io.targets;
    u.counts = reinterpret_cast<unsigned long long*>(io.counts4);
    u.B = B; u.Hl = Hl; u.Wl = Wl; u.H = P.desc.in_h; u.W = P.desc.in_w; u.NC = nc;
    RC(launch_upsample_out(u, st));
  }
#undef RC
  return MTG_OK;
}

}  // namespace mtgseg
