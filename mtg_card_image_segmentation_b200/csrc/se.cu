// Pooled-feature micro kernels: global average pool (channel sums) and the per-image MLP used by the
// squeeze-excite blocks (tv:ops/misc.py:252-261: FC+bias -> ReLU -> FC+bias -> Hardsigmoid) and by the
// head's scale branch (train/model.py:115-119: GAP -> 1x1 conv (no bias) -> Sigmoid).
// These are latency-bound (a few hundred kFLOP per image); the design goal is few launches and no
// re-read of the big activation: the depthwise kernel already produced the channel sums.
#include "ops.h"

namespace mtgseg {
namespace {

// in [B][HW][C] bf16 -> out [B][C] fp32 sums.  grid (ceil(CV/8), B); block 256 = 8 vectors x 32 pixel lanes.
__global__ void __launch_bounds__(256) gap_kernel(const bf16* __restrict__ in, float* __restrict__ out, int HW, int C) {
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

constexpr int IPC = 2;  // images per CTA (each weight element is loaded once per IPC images)

struct SeP {
  const float* sums; int chunks, B, C, SQ; float inv_hw;
  const bf16* w1; const float* b1; int act1;
  const bf16* w2; const float* b2; int act2;
  float* out;
};

__global__ void __launch_bounds__(256) se_mlp_kernel(const SeP p) {
  extern __shared__ float sm[];
  float* mean = sm;                   // [IPC][C]
  float* hid = sm + IPC * p.C;        // [IPC][SQ]
  const int n0 = blockIdx.x * IPC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < IPC; ++i) {
    const int n = n0 + i;
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      float s = 0.f;
      if (n < p.B)
        for (int k = 0; k < p.chunks; ++k) s += p.sums[(static_cast<size_t>(n) * p.chunks + k) * p.C + c];
      mean[i * p.C + c] = s * p.inv_hw;
    }
  }
  __syncthreads();
  // layer 1: one warp per hidden unit, lanes stride over C in 8-channel vectors
  for (int j = warp; j < p.SQ; j += 8) {
    float acc[IPC];
#pragma unroll
    for (int i = 0; i < IPC; ++i) acc[i] = 0.f;
    const bf16* wr = p.w1 + static_cast<size_t>(j) * p.C;
    for (int c = lane * 8; c < p.C; c += 256) {
      float wf[8];
      unpack8(ldg16(wr + c), wf);
#pragma unroll
      for (int i = 0; i < IPC; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i] = fmaf(wf[e], mean[i * p.C + c + e], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < IPC; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane == 0) {
      const float b = p.b1 ? p.b1[j] : 0.f;
#pragma unroll
      for (int i = 0; i < IPC; ++i) {
        const float h = apply_act(acc[i] + b, p.act1);
        hid[i * p.SQ + j] = h;
        if (!p.w2 && n0 + i < p.B) p.out[static_cast<size_t>(n0 + i) * p.SQ + j] = h;
      }
    }
  }
  if (!p.w2) return;
  __syncthreads();
  // layer 2: one warp per output channel, lanes stride over SQ (SQ % 8 == 0)
  for (int c = warp; c < p.C; c += 8) {
    float acc[IPC];
#pragma unroll
    for (int i = 0; i < IPC; ++i) acc[i] = 0.f;
    const bf16* wr = p.w2 + static_cast<size_t>(c) * p.SQ;
    for (int j = lane * 8; j < p.SQ; j += 256) {
      float wf[8];
      unpack8(ldg16(wr + j), wf);
#pragma unroll
      for (int i = 0; i < IPC; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i] = fmaf(wf[e], hid[i * p.SQ + j + e], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < IPC; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane == 0) {
      const float b = p.b2 ? p.b2[c] : 0.f;
#pragma unroll
      for (int i = 0; i < IPC; ++i)
        if (n0 + i < p.B) p.out[static_cast<size_t>(n0 + i) * p.C + c] = apply_act(acc[i] + b, p.act2);
    }
  }
}

}  // namespace

int launch_gap(const bf16* in, float* out, int B, int HW, int C, cudaStream_t st) {
  MTG_REQUIRE(in && out && C % 8 == 0, MTG_ERR_ARG, "gap: bad arguments");
  dim3 grid(ceil_div(C / 8, 8), B);
  gap_kernel<<<grid, 256, 0, st>>>(in, out, HW, C);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_se_mlp(const SeMlpArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.sums && a.w1 && a.out && a.HW > 0, MTG_ERR_ARG, "se_mlp: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && a.SQ % 8 == 0, MTG_ERR_UNSUPPORTED, "se_mlp: C=%d SQ=%d must be multiples of 8", a.C, a.SQ);
  SeP p{a.sums, a.chunks, a.B, a.C, a.SQ, 1.0f / static_cast<float>(a.HW), a.w1, a.b1, a.act1, a.w2, a.b2, a.act2, a.out};
  const size_t smem = static_cast<size_t>(IPC) * (a.C + a.SQ) * sizeof(float);
  MTG_REQUIRE(smem <= 48 * 1024, MTG_ERR_UNSUPPORTED, "se_mlp: C too large");
  se_mlp_kernel<<<ceil_div(a.B, IPC), 256, smem, st>>>(p);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
