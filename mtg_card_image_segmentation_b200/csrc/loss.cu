// CombinedLoss = dice_w * DiceLoss + ce_w * CrossEntropy (train/utils.py:15-92) and its gradient, fused.
//
//   p = softmax(z) over classes;  S = sum_pixels p[target];  N = #pixels
//   dice = 1 - (2S + eps) / (2N + eps)      (ONE global scalar: sum p == N, sum one_hot == N)
//   ce   = -(1/N) sum log p[target]
//   dL/dz_c = ce_w (p_c - y_c)/N - dice_w * 2/(2N+eps) * p_t (y_c - p_c)
//
// The gradient needs no global quantity, so loss partial sums and dlogits come out of ONE pass over the
// logits (the reference runs ~12 elementwise/reduction kernels and materialises a one-hot tensor).
// Partial sums are written per block and reduced in fixed order by a one-block kernel: deterministic.
#include <cuda_fp16.h>

#include "ops.h"

namespace mtgseg {
namespace {

constexpr int MAX_NC = 8;
constexpr int LOSS_BLOCKS = 148 * 4;

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half(v); }

// T: logits element type; G: gradient element type (T, or float: fp16 logits get fp32 gradients, the unscaled values
// (~1e-7 at B=32, 320x240) are fp16 subnormals and must not be narrowed before the GradScaler factor is applied)
template <typename T, typename G>
__global__ void __launch_bounds__(256) loss_kernel(const T* __restrict__ logits, const int64_t* __restrict__ targets,
                                                   G* __restrict__ dlogits, float* __restrict__ partials, long long batch,
                                                   long long hw, int nc, float g_ce, float g_dice) {
  __shared__ float red[2][8];
  float s_pt = 0.f, s_log = 0.f;
  const long long total = batch * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / hw, px = i - n * hw;
    const T* zp = logits + n * nc * hw + px;
    float z[MAX_NC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c)
      if (c < nc) { z[c] = ldf(zp + c * hw); m = fmaxf(m, z[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c)
      if (c < nc) { z[c] = expf(z[c] - m); sum += z[c]; }
    const float inv = 1.f / sum;
    const int t = static_cast<int>(targets[i]);
    float pt = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c)
      if (c < nc) { z[c] *= inv; if (c == t) pt = z[c]; }
    s_pt += pt;
    s_log += logf(pt);
    if (dlogits) {
      G* gp = dlogits + n * nc * hw + px;
#pragma unroll
      for (int c = 0; c < MAX_NC; ++c)
        if (c < nc) {
          const float y = (c == t) ? 1.f : 0.f;
          stf(gp + c * hw, g_ce * (z[c] - y) - g_dice * pt * (y - z[c]));
        }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_pt += __shfl_xor_sync(0xffffffffu, s_pt, o);
    s_log += __shfl_xor_sync(0xffffffffu, s_log, o);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][warp] = s_pt; red[1][warp] = s_log; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

__global__ void loss_finalize_kernel(const float* __restrict__ partials, int blocks, double n, float dice_w, float ce_w,
                                     float smooth, float* __restrict__ out) {
  __shared__ double sa[256], sb[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 256) { a += partials[2 * i]; b += partials[2 * i + 1]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double dice = 1.0 - (2.0 * sa[0] + smooth) / (2.0 * n + smooth);
    const double ce = -sb[0] / n;
    out[0] = static_cast<float>(dice_w * dice + ce_w * ce);
    out[1] = static_cast<float>(dice);
    out[2] = static_cast<float>(ce);
  }
}

// ---------------------------------------------------------------------------------------------------------
// The same loss for the training step, computed from the LOW-RESOLUTION logits (40x30) the head produces: the x8 bilinear
// upsample (tv:models/segmentation/lraspp.py:46) is linear, so the full-resolution logits are recomputed on the fly (same
// arithmetic as upsample_out_kernel), the per-pixel softmax / CE / Dice terms are summed, and the gradient is pulled back
// to the 40x30 grid by the transposed interpolation -- no full-resolution logits or dlogits tensor exists in the step.
// Gather form: one thread per low-resolution pixel walks the fine pixels whose interpolation touches it (fixed order: this
// gradient feeds the whole backward chain and must be reproducible); every fine pixel is visited by its (up to) four corner
// owners, and counted for the loss sums by the owner of its top-left corner only.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index_l(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

struct LowLossP {
  const float* lowres; const int64_t* targets; float* d_lowres; float* partials;
  int Hl, Wl, H, W, NC; float g_ce, g_dice;
};
constexpr int LL_CP = 12;     // low-resolution pixels per CTA
constexpr int LL_SLOTS = 20;  // fine-row slots per low-resolution pixel (a x8 window spans <= 18 fine rows)
// grid (ceil(Hl*Wl / LL_CP), B), LL_CP * LL_SLOTS threads: thread (slot, cp) walks ONE fine row of the window of low-resolution
// pixel cp; the slots' partial sums meet in shared memory and are added in slot order (deterministic).
__global__ void __launch_bounds__(LL_CP * LL_SLOTS) lowres_loss_kernel(const LowLossP p) {
  extern __shared__ float lo[];  // [Hl*Wl][NC] of this image
  __shared__ float part[LL_SLOTS][LL_CP][MAX_NC + 2];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int nlo = p.Hl * p.Wl * p.NC;
  for (int i = threadIdx.x; i < nlo; i += blockDim.x) lo[i] = p.lowres[static_cast<size_t>(b) * nlo + i];
  __syncthreads();
  const int cp = threadIdx.x % LL_CP, slot = threadIdx.x / LL_CP;
  const int q = blockIdx.x * LL_CP + cp;
  float acc[MAX_NC];
#pragma unroll
  for (int c = 0; c < MAX_NC; ++c) acc[c] = 0.f;
  float s_pt = 0.f, s_log = 0.f;
  if (q < p.Hl * p.Wl) {
    const int qy = q / p.Wl, qx = q - qy * p.Wl;
    const float sy = static_cast<float>(p.Hl) / p.H, sx = static_cast<float>(p.Wl) / p.W;
    // fine rows / columns whose source interval can touch (qy, qx): src in (q - 1, q + 1)
    const int y_lo = max(0, static_cast<int>(floorf((qy - 1 + 0.5f) / sy - 0.5f)) - 1);
    const int y_hi = min(p.H - 1, static_cast<int>(ceilf((qy + 1 + 0.5f) / sy - 0.5f)) + 1);
    const int x_lo = max(0, static_cast<int>(floorf((qx - 1 + 0.5f) / sx - 0.5f)) - 1);
    const int x_hi = min(p.W - 1, static_cast<int>(ceilf((qx + 1 + 0.5f) / sx - 0.5f)) + 1);
    const int64_t* tb = p.targets + static_cast<size_t>(b) * p.H * p.W;
    for (int y = y_lo + slot; y <= y_hi; y += LL_SLOTS) {  // one row per thread for x8 (more only for larger factors)
      int y0, y1; float ly;
      src_index_l(y, sy, p.Hl, y0, y1, ly);
      const float wy = (y0 == qy ? 1.f - ly : 0.f) + (y1 == qy ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x = x_lo; x <= x_hi; ++x) {
        int x0, x1; float lx;
        src_index_l(x, sx, p.Wl, x0, x1, lx);
        const float wx = (x0 == qx ? 1.f - lx : 0.f) + (x1 == qx ? lx : 0.f);
        if (wx == 0.f) continue;
        // full-resolution logits of pixel (y, x): the arithmetic of upsample_out_kernel
        float z[MAX_NC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAX_NC; ++c)
          if (c < p.NC) {
            const float v00 = lo[(y0 * p.Wl + x0) * p.NC + c], v01 = lo[(y0 * p.Wl + x1) * p.NC + c];
            const float v10 = lo[(y1 * p.Wl + x0) * p.NC + c], v11 = lo[(y1 * p.Wl + x1) * p.NC + c];
            z[c] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
            m = fmaxf(m, z[c]);
          }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < MAX_NC; ++c)
          if (c < p.NC) { z[c] = expf(z[c] - m); sum += z[c]; }
        const float inv = 1.f / sum;
        const int t = static_cast<int>(tb[static_cast<size_t>(y) * p.W + x]);
        float pt = 0.f;
#pragma unroll
        for (int c = 0; c < MAX_NC; ++c)
          if (c < p.NC) { z[c] *= inv; if (c == t) pt = z[c]; }
        if (y0 == qy && x0 == qx) { s_pt += pt; s_log += logf(pt); }  // counted once, by its top-left corner's owner
        const float w = wy * wx;
#pragma unroll
        for (int c = 0; c < MAX_NC; ++c)
          if (c < p.NC) {
            const float yv = (c == t) ? 1.f : 0.f;
            acc[c] = fmaf(w, p.g_ce * (z[c] - yv) - p.g_dice * pt * (yv - z[c]), acc[c]);
          }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MAX_NC; ++c) part[slot][cp][c] = acc[c];
  part[slot][cp][MAX_NC] = s_pt;
  part[slot][cp][MAX_NC + 1] = s_log;
  __syncthreads();
  if (slot == 0 && q < p.Hl * p.Wl) {
    for (int c = 0; c < p.NC; ++c) {
      float t = 0.f;
      for (int k = 0; k < LL_SLOTS; ++k) t += part[k][cp][c];
      p.d_lowres[(static_cast<size_t>(b) * p.Hl * p.Wl + q) * p.NC + c] = t;
    }
  }
  if (threadIdx.x == 0) {
    float a = 0.f, bb = 0.f;
    for (int k = 0; k < LL_SLOTS; ++k)
      for (int j = 0; j < LL_CP; ++j) { a += part[k][j][MAX_NC]; bb += part[k][j][MAX_NC + 1]; }
    const int s = blockIdx.y * gridDim.x + blockIdx.x;
    p.partials[2 * s] = a;
    p.partials[2 * s + 1] = bb;
  }
}

}  // namespace

size_t loss_scratch_bytes() { return sizeof(float) * 2 * LOSS_BLOCKS; }
size_t lowres_loss_scratch_floats(int B, int Hl, int Wl) { return static_cast<size_t>(2) * B * ceil_div(Hl * Wl, LL_CP); }

int launch_lowres_loss(const float* lowres, const int64_t* targets, float* d_lowres, float* scratch, float* loss3, int B, int Hl, int Wl,
                       int H, int W, int nc, float dice_w, float ce_w, float smooth, cudaStream_t st) {
  MTG_REQUIRE(lowres && targets && d_lowres && scratch && loss3, MTG_ERR_ARG, "lowres_loss: null pointer");
  MTG_REQUIRE(nc >= 2 && nc <= MAX_NC, MTG_ERR_UNSUPPORTED, "lowres_loss: num_classes %d not in [2,%d]", nc, MAX_NC);
  const double n = static_cast<double>(B) * H * W;
  LowLossP p{lowres, targets, d_lowres, scratch, Hl, Wl, H, W, nc, static_cast<float>(ce_w / n),
             static_cast<float>(dice_w * 2.0 / (2.0 * n + smooth))};
  const dim3 grid(ceil_div(Hl * Wl, LL_CP), B);
  const size_t smem = sizeof(float) * static_cast<size_t>(Hl) * Wl * nc;
  MTG_REQUIRE(smem <= 40 * 1024, MTG_ERR_UNSUPPORTED, "lowres_loss: low-resolution map too large (%zu B)", smem);
  MTG_CUDA(launch_pdl(lowres_loss_kernel, grid, dim3(LL_CP * LL_SLOTS), smem, st, p));
  MTG_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 256, 0, st>>>(scratch, static_cast<int>(grid.x * grid.y), n, dice_w, ce_w, smooth, loss3);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_loss(const void* logits, int dtype, const int64_t* targets, void* dlogits, int dlogits_dtype, float* scratch, float* loss3,
                long long batch, long long hw, int nc, float dice_w, float ce_w, float smooth, cudaStream_t st) {
  MTG_REQUIRE(logits && targets && scratch && loss3, MTG_ERR_ARG, "loss: null pointer");
  MTG_REQUIRE(nc >= 2 && nc <= MAX_NC, MTG_ERR_UNSUPPORTED, "loss: num_classes %d not in [2,%d]", nc, MAX_NC);
  MTG_REQUIRE(batch > 0 && hw > 0, MTG_ERR_ARG, "loss: empty batch");
  const double n = static_cast<double>(batch) * static_cast<double>(hw);
  const float g_ce = static_cast<float>(ce_w / n);
  const float g_dice = static_cast<float>(dice_w * 2.0 / (2.0 * n + smooth));
  long long blocks = (batch * hw + 256 * 4 - 1) / (256 * 4);
  if (blocks > LOSS_BLOCKS) blocks = LOSS_BLOCKS;
  const int g = static_cast<int>(blocks);
  if (!dlogits) dlogits_dtype = dtype;
  MTG_REQUIRE(dlogits_dtype == dtype || dlogits_dtype == LOGITS_F32, MTG_ERR_UNSUPPORTED,
              "loss: dlogits must have the logits' dtype or be float32 (got %d for logits %d)", dlogits_dtype, dtype);
  const bool gf = dlogits_dtype == LOGITS_F32;
#define MTG_LOSS_LAUNCH(T, G) \
  loss_kernel<T, G><<<g, 256, 0, st>>>(static_cast<const T*>(logits), targets, static_cast<G*>(dlogits), scratch, batch, hw, nc, g_ce, g_dice)
  if (dtype == LOGITS_F32) MTG_LOSS_LAUNCH(float, float);
  else if (dtype == LOGITS_BF16) { if (gf) MTG_LOSS_LAUNCH(bf16, float); else MTG_LOSS_LAUNCH(bf16, bf16); }
  else if (dtype == LOGITS_F16) { if (gf) MTG_LOSS_LAUNCH(__half, float); else MTG_LOSS_LAUNCH(__half, __half); }
  else MTG_REQUIRE(false, MTG_ERR_ARG, "loss: unknown logits dtype %d", dtype);
#undef MTG_LOSS_LAUNCH
  MTG_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 256, 0, st>>>(scratch, g, n, dice_w, ce_w, smooth, loss3);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
