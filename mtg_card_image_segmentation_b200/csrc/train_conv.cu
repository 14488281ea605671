// Backward kernels of the convolutions that are not (yet) on the tensor cores:
//   wgrad of the 1x1 / 3x3 convolutions (CUDA-core tiled GEMM over the pixel dimension, split-K + fp32 atomics),
//   depthwise dgrad and wgrad, stem wgrad.
// dgrad of the 1x1 / 3x3 convolutions reuses the tcgen05 implicit-GEMM kernel (gemm_tc.cu) with transposed weights.
// Gradients of parameters are accumulated in fp32 directly in the reference's OIHW layout.
#include "ops.h"

namespace mtgseg {

int group_vectors(int CV);  // dwconv.cu

namespace {

// ---------------------------------------------------------------------------------------------------------
// dW[n][k][tap] += sum_m dz[m][n] * x[shift_tap(m)][k] (* a_scale[img(m)][k])
// CTA tile 64 (n) x 64 (k), 256 threads x (4 x 4) outputs, 16 pixels per smem step.
// ---------------------------------------------------------------------------------------------------------
struct WgP {
  const bf16* dz; const bf16* x; float* dw; const float* a_scale;
  long long M; int N, K, taps, H, W, hw;  // taps = 1 or 9 (3x3 pad 1)
  long long rows_per_chunk;
};

__global__ void __launch_bounds__(256) wgrad_kernel(const WgP p) {
  __shared__ __align__(16) float sdz[16][64];
  __shared__ __align__(16) float sx[16][64];
  const int k_tiles = (p.K + 63) / 64;
  const int n0 = (blockIdx.x / k_tiles) * 64, k0 = (blockIdx.x % k_tiles) * 64;
  const int tap = blockIdx.y;
  const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
  const long long m_begin = blockIdx.z * p.rows_per_chunk;
  const long long m_end = min(p.M, m_begin + p.rows_per_chunk);
  const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;
  // loader role: threads 0..127 load dz (16 px x 8 vectors), 128..255 load x
  const int lrole = threadIdx.x >> 7, lt = threadIdx.x & 127, lp = lt >> 3, lv = lt & 7;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long m0 = m_begin; m0 < m_end; m0 += 16) {
    const long long m = m0 + lp;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (m < m_end) {
      if (lrole == 0) {
        const int n = n0 + lv * 8;
        if (n < p.N) unpack8(ldg16(p.dz + m * p.N + n), f);
      } else {
        const int k = k0 + lv * 8;
        if (k < p.K) {
          long long ms = m;
          bool ok = true;
          if (p.taps == 9) {
            const int px = static_cast<int>(m % p.W), py = static_cast<int>((m / p.W) % p.H);
            ok = (px + dx >= 0) && (px + dx < p.W) && (py + dy >= 0) && (py + dy < p.H);
            ms = m + dy * p.W + dx;
          }
          if (ok) {
            unpack8(ldg16(p.x + ms * p.K + k), f);
            if (p.a_scale) {
              const float* sp = p.a_scale + (m / p.hw) * p.K + k;
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j] * sp[j]));  // what the forward MMA consumed
            }
          }
        }
      }
    }
    float* dst = lrole == 0 ? &sdz[lp][lv * 8] : &sx[lp][lv * 8];
    *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < 16; ++pp) {
      const float4 a = *reinterpret_cast<const float4*>(&sdz[pp][tn * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sx[pp][tk * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + j;
      if (n < p.N && k < p.K) atomicAdd(p.dw + (static_cast<size_t>(n) * p.K + k) * p.taps + tap, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// depthwise
// ---------------------------------------------------------------------------------------------------------
struct DwBwdP {
  const bf16* dz; const bf16* x; const bf16* w; bf16* dx; float* dw;
  int H, W, C, Ho, Wo, k, stride, dil, pad, CV, CVc, PL, rows_per_chunk, chunks;
  int B, imgs_per_cta;  // wgrad: a CTA accumulates over several images before its block reduction + atomics
};

// dx[n,iy,ix,c] = sum_taps w[tap][c] * dz[n,(iy+pad-ky*d)/s,(ix+pad-kx*d)/s,c]   (grid: chunks, groups, B)
template <int KS>
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const DwBwdP p) {
  pdl_trigger();
  pdl_wait();
  const int vl = threadIdx.x % p.CVc, pl = threadIdx.x / p.CVc;
  const int v = blockIdx.y * p.CVc + vl;
  if (pl >= p.PL || v >= p.CV) return;
  const int c0 = v * 8, n = blockIdx.z;
  const int npix = p.H * p.W;
  const int r0 = blockIdx.x * p.rows_per_chunk, r1 = min(npix, r0 + p.rows_per_chunk);
  const bf16* dz_n = p.dz + static_cast<size_t>(n) * p.Ho * p.Wo * p.C + c0;
  bf16* dx_n = p.dx + static_cast<size_t>(n) * npix * p.C + c0;
  for (int pix = r0 + pl; pix < r1; pix += p.PL) {
    const int iy = pix / p.W, ix = pix - iy * p.W;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const int ty = iy + p.pad - ky * p.dil;
      if (ty < 0 || ty % p.stride) continue;
      const int oy = ty / p.stride;
      if (oy >= p.Ho) continue;
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const int tx = ix + p.pad - kx * p.dil;
        if (tx < 0 || tx % p.stride) continue;
        const int ox = tx / p.stride;
        if (ox >= p.Wo) continue;
        float g[8], wf[8];
        unpack8(ldg16(dz_n + (static_cast<size_t>(oy) * p.Wo + ox) * p.C), g);
        unpack8(ldg16(p.w + (ky * KS + kx) * p.C + c0), wf);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], wf[j], acc[j]);
      }
    }
    *reinterpret_cast<uint4*>(dx_n + static_cast<size_t>(pix) * p.C) = pack8(acc);
  }
}

// dW[c][ky][kx] += sum dz[n,oy,ox,c] * x[n,oy*s-pad+ky*d,ox*s-pad+kx*d,c]; one kernel row (ky) per CTA:
// grid (chunks*KS, groups, ceil(B / imgs_per_cta)): the per-CTA cost is the block reduction + cw*KS atomics at the end, so a
// CTA walks imgs_per_cta images of its (pixel chunk, kernel row) first (B=32, 20x15, C=960: 3840 CTAs of <= 8 pixels per
// thread before, ~600 now)
template <int KS>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const DwBwdP p) {
  __shared__ float red[256 * 8];
  const int vl = threadIdx.x % p.CVc, pl = threadIdx.x / p.CVc;
  const int v = blockIdx.y * p.CVc + vl;
  const bool active = pl < p.PL && v < p.CV;
  const int c0 = (active ? v : 0) * 8;
  const int chunk = blockIdx.x / KS, ky = blockIdx.x % KS;
  float acc[KS][8];
#pragma unroll
  for (int kx = 0; kx < KS; ++kx)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[kx][j] = 0.f;
  if (active) {
    const int npix = p.Ho * p.Wo;
    const int r0 = chunk * p.rows_per_chunk, r1 = min(npix, r0 + p.rows_per_chunk);
    const int n_begin = blockIdx.z * p.imgs_per_cta, n_end = min(p.B, n_begin + p.imgs_per_cta);
    for (int n = n_begin; n < n_end; ++n) {
      const bf16* dz_n = p.dz + static_cast<size_t>(n) * npix * p.C + c0;
      const bf16* x_n = p.x + static_cast<size_t>(n) * p.H * p.W * p.C + c0;
      for (int pix = r0 + pl; pix < r1; pix += p.PL) {
        const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
        const int iy = oy * p.stride - p.pad + ky * p.dil;
        if (iy < 0 || iy >= p.H) continue;
        float g[8];
        unpack8(ldg16(dz_n + static_cast<size_t>(pix) * p.C), g);
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int ix = ox * p.stride - p.pad + kx * p.dil;
          if (ix < 0 || ix >= p.W) continue;
          float xf[8];
          unpack8(ldg16(x_n + (static_cast<size_t>(iy) * p.W + ix) * p.C), xf);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[kx][j] = fmaf(g[j], xf[j], acc[kx][j]);
        }
      }
    }
  }
  const int cw = p.CVc * 8;
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) {
    __syncthreads();
    if (pl < p.PL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(pl * p.CVc + vl) * 8 + j] = active ? acc[kx][j] : 0.f;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < p.C) {
        float s = 0.f;
        for (int r = 0; r < p.PL; ++r) s += red[r * cw + cl];
        atomicAdd(p.dw + static_cast<size_t>(c) * KS * KS + ky * KS + kx, s);
      }
    }
  }
}

// Row-walking form of the same sum (the default): a thread owns one (image, output row, ky) and walks ox with the input row's
// window in registers, so every x vector is loaded ONCE per (row, ky) instead of once per tap (the per-pixel form above
// re-reads x KS times per pixel through L1/L2: 648 us for the 20x15x960 5x5 layers at B=256, 7 % of HBM speed).
// Ring of R = WIN + S packed vectors, WIN = (KS-1)*DIL + 1: window position j of step u (ix = ox*S - pad + j) lives in slot
// (u*S + j) % R; the S vectors entering the window at step u+1 and the dz vector of step u+1 are loaded at the top of step u.
// The ox loop is unrolled by R so that all slot indices are compile-time.  grid (ctas*KS, groups).
template <int KS, int S, int DIL>
__global__ void __launch_bounds__(256) dw_wgrad_rows_kernel(const DwBwdP p, int tasks_per_cta) {
  pdl_trigger();
  pdl_wait();
  constexpr int WIN = (KS - 1) * DIL + 1, R = WIN + S;
  __shared__ float red[256 * 8];
  const int vl = threadIdx.x % p.CVc, pl = threadIdx.x / p.CVc;
  const int v = blockIdx.y * p.CVc + vl;
  const bool active = pl < p.PL && v < p.CV;
  const int c0 = (active ? v : 0) * 8;
  const int cta = blockIdx.x / KS, ky = blockIdx.x % KS;
  float acc[KS][8];
#pragma unroll
  for (int kx = 0; kx < KS; ++kx)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[kx][j] = 0.f;
  if (active) {
    const int total = p.B * p.Ho;
    const int t0 = cta * tasks_per_cta, t1 = min(total, t0 + tasks_per_cta);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int t = t0 + pl; t < t1; t += p.PL) {
      const int n = t / p.Ho, oy = t - n * p.Ho;
      const int iy = oy * S - p.pad + ky * DIL;
      if (iy < 0 || iy >= p.H) continue;
      const bf16* xr = p.x + (static_cast<size_t>(n) * p.H + iy) * p.W * p.C + c0;
      const bf16* gr = p.dz + (static_cast<size_t>(n) * p.Ho + oy) * p.Wo * p.C + c0;
      uint4 win[R];
#pragma unroll
      for (int j = 0; j < WIN; ++j) {
        const int ix = j - p.pad;
        win[j] = (ix >= 0 && ix < p.W) ? ldg16(xr + static_cast<size_t>(ix) * p.C) : zero;
      }
      uint4 gq = ldg16(gr);
      for (int ox0 = 0; ox0 < p.Wo; ox0 += R) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const int ox = ox0 + u;
          if (ox < p.Wo) {
            const uint4 gcur = gq;
            if (ox + 1 < p.Wo) gq = ldg16(gr + static_cast<size_t>(ox + 1) * p.C);
#pragma unroll
            for (int sft = 0; sft < S; ++sft) {  // positions WIN .. WIN+S-1: needed from the next step on
              const int ix = ox * S - p.pad + WIN + sft;
              win[(u * S + WIN + sft) % R] = (ix >= 0 && ix < p.W) ? ldg16(xr + static_cast<size_t>(ix) * p.C) : zero;
            }
            float g[8];
            unpack8(gcur, g);
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              float xf[8];
              unpack8(win[(u * S + kx * DIL) % R], xf);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[kx][j] = fmaf(g[j], xf[j], acc[kx][j]);
            }
          }
        }
      }
    }
  }
  const int cw = p.CVc * 8;
#pragma unroll
  for (int kx = 0; kx < KS; ++kx) {
    __syncthreads();
    if (pl < p.PL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(pl * p.CVc + vl) * 8 + j] = active ? acc[kx][j] : 0.f;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < p.C) {
        float sum = 0.f;
        for (int r = 0; r < p.PL; ++r) sum += red[r * cw + cl];
        atomicAdd(p.dw + static_cast<size_t>(c) * KS * KS + ky * KS + kx, sum);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// stem wgrad: dW[o][ci][ky][kx] += sum dz[n,oy,ox,o] * x[n,ci,2oy-1+ky,2ox-1+kx]   (432 outputs)
// CTA = 9 warps, warp w owns (ci, ky) = (w / 3, w % 3): 3 kx x 16 o = 48 accumulators per lane in registers.  The lanes walk the
// output pixels of a row (ox = lane + 32 s), rows (n, oy) are dealt to the CTAs grid-stride: per pixel a lane loads its 16
// gradients (32 contiguous bytes) and two input values (the third tap is the neighbouring lane's, by shuffle) and issues 48 FMAs.
// (The first version staged 64-pixel patches in shared memory and spent two shared-memory loads per FMA plus an index
// decomposition per gathered element: 1.2 ms at B = 256 for 4.2 GFLOP.)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dz,
                                                         float* __restrict__ dw, int B, int H, int W, int Ho, int Wo) {
  pdl_trigger();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci = warp / 3, ky = warp - ci * 3;
  float acc[3][16];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[k][o] = 0.f;
  const int rows = B * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, oy = row - n * Ho;
    const int iy = 2 * oy - 1 + ky;
    if (iy < 0 || iy >= H) continue;  // zero padding row (warp-uniform)
    const float* xr = x + ((static_cast<size_t>(n) * 3 + ci) * H + iy) * W;
    const bf16* dr = dz + static_cast<size_t>(row) * Wo * 16;
#pragma unroll 2
    for (int ox0 = 0; ox0 < Wo; ox0 += 32) {
      const int ox = ox0 + lane;
      const bool live = ox < Wo;
      const int c = 2 * ox;
      const float x0 = (live && c < W) ? __ldg(xr + c) : 0.f;
      const float x1 = (live && c + 1 < W) ? __ldg(xr + c + 1) : 0.f;
      uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;
      if (live) {
        g0 = ldg16(dr + static_cast<size_t>(ox) * 16);
        g1 = ldg16(dr + static_cast<size_t>(ox) * 16 + 8);
      }
      float xm = __shfl_up_sync(0xffffffffu, x1, 1);  // column 2*ox - 1 is the previous pixel's column 2*(ox-1) + 1
      if (lane == 0) xm = (live && ox > 0) ? __ldg(xr + c - 1) : 0.f;
      float g[16];
      {
        float lo[8], hi[8];
        unpack8(g0, lo);
        unpack8(g1, hi);
#pragma unroll
        for (int o = 0; o < 8; ++o) { g[o] = lo[o]; g[8 + o] = hi[o]; }
      }
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        acc[0][o] = fmaf(g[o], xm, acc[0][o]);
        acc[1][o] = fmaf(g[o], x0, acc[1][o]);
        acc[2][o] = fmaf(g[o], x1, acc[2][o]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      float v = acc[k][o];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) atomicAdd(dw + o * 27 + ci * 9 + ky * 3 + k, v);
    }
}

}  // namespace

int launch_wgrad(const WgradArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.dz && a.x && a.dw, MTG_ERR_ARG, "wgrad: null pointer");
  MTG_REQUIRE(a.N % 8 == 0 && a.K % 8 == 0, MTG_ERR_UNSUPPORTED, "wgrad: channels must be multiples of 8");
  MTG_REQUIRE(a.taps == 1 || a.taps == 9, MTG_ERR_UNSUPPORTED, "wgrad: taps must be 1 or 9");
  WgP p{a.dz, a.x, a.dw, a.a_scale, a.M, a.N, a.K, a.taps, a.H, a.W, a.hw, 0};
  const int tiles = ceil_div(a.N, 64) * ceil_div(a.K, 64);
  // enough pixel chunks to fill the machine (~4 CTAs per SM), at least 256 pixels per chunk
  long long chunks = (148 * 4 + tiles * a.taps - 1) / (tiles * a.taps);
  const long long max_chunks = (a.M + 255) / 256;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  p.rows_per_chunk = (((a.M + chunks - 1) / chunks) + 15) / 16 * 16;
  chunks = (a.M + p.rows_per_chunk - 1) / p.rows_per_chunk;
  dim3 grid(tiles, a.taps, static_cast<unsigned>(chunks));
  wgrad_kernel<<<grid, 256, 0, st>>>(p);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

static int fill_dw(const DwBwdArgs& a, DwBwdP& p, int rows) {
  p.dz = a.dz; p.x = a.x; p.w = a.w; p.dx = a.dx; p.dw = a.dw;
  p.H = a.H; p.W = a.W; p.C = a.C; p.k = a.k; p.stride = a.stride; p.dil = a.dil;
  p.pad = (a.k - 1) / 2 * a.dil;
  p.Ho = (a.H + 2 * p.pad - a.dil * (a.k - 1) - 1) / a.stride + 1;
  p.Wo = (a.W + 2 * p.pad - a.dil * (a.k - 1) - 1) / a.stride + 1;
  p.CV = a.C / 8; p.CVc = group_vectors(p.CV); p.PL = 256 / p.CVc;
  const int npix = rows == 0 ? a.H * a.W : p.Ho * p.Wo;
  p.chunks = ceil_div(npix, p.PL * 8);
  if (p.chunks > 64) p.chunks = 64;
  p.rows_per_chunk = ceil_div(npix, p.chunks);
  return MTG_OK;
}

int launch_dw_dgrad(const DwBwdArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.dz && a.w && a.dx, MTG_ERR_ARG, "dw_dgrad: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && (a.k == 3 || a.k == 5), MTG_ERR_UNSUPPORTED, "dw_dgrad: unsupported shape");
  DwBwdP p{};
  fill_dw(a, p, 0);
  dim3 grid(p.chunks, ceil_div(p.CV, p.CVc), a.B);
  if (a.k == 3) MTG_CUDA(launch_pdl(dw_dgrad_kernel<3>, dim3(grid), dim3(256), 0, st, p));
  else MTG_CUDA(launch_pdl(dw_dgrad_kernel<5>, dim3(grid), dim3(256), 0, st, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_dw_wgrad(const DwBwdArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.dz && a.x && a.dw, MTG_ERR_ARG, "dw_wgrad: null pointer");
  MTG_REQUIRE(a.C % 8 == 0 && (a.k == 3 || a.k == 5), MTG_ERR_UNSUPPORTED, "dw_wgrad: unsupported shape");
  DwBwdP p{};
  fill_dw(a, p, 1);
  p.B = a.B;
  const int groups = ceil_div(p.CV, p.CVc);
  p.imgs_per_cta = ceil_div(p.chunks * a.k * groups * a.B, 148 * 4);  // ~4 CTAs per SM in total
  if (p.imgs_per_cta < 1) p.imgs_per_cta = 1;
  if (p.imgs_per_cta > a.B) p.imgs_per_cta = a.B;
  dim3 grid(p.chunks * a.k, groups, ceil_div(a.B, p.imgs_per_cta));
  static const bool per_pixel = [] { const char* e = getenv("MTGSEG_DW_WGRAD"); return e && e[0] == 'p'; }();  // A/B: the older form
  const int combo = per_pixel ? 0 : a.k * 100 + a.stride * 10 + a.dil;
  if (combo == 311 || combo == 321 || combo == 511 || combo == 521 || combo == 512) {
    // ~6 CTAs per SM in total, at least one (image, row) task per row lane
    const int total = a.B * p.Ho;
    int ctas = ceil_div(148 * 6, a.k * groups);
    if (ctas > ceil_div(total, p.PL)) ctas = ceil_div(total, p.PL);
    if (ctas < 1) ctas = 1;
    const int tpc = ceil_div(total, ctas);
    ctas = ceil_div(total, tpc);
    const dim3 rgrid(ctas * a.k, groups);
    switch (combo) {
      case 311: MTG_CUDA(launch_pdl(dw_wgrad_rows_kernel<3, 1, 1>, dim3(rgrid), dim3(256), 0, st, p, tpc)); break;
      case 321: MTG_CUDA(launch_pdl(dw_wgrad_rows_kernel<3, 2, 1>, dim3(rgrid), dim3(256), 0, st, p, tpc)); break;
      case 511: MTG_CUDA(launch_pdl(dw_wgrad_rows_kernel<5, 1, 1>, dim3(rgrid), dim3(256), 0, st, p, tpc)); break;
      case 521: MTG_CUDA(launch_pdl(dw_wgrad_rows_kernel<5, 2, 1>, dim3(rgrid), dim3(256), 0, st, p, tpc)); break;
      default: MTG_CUDA(launch_pdl(dw_wgrad_rows_kernel<5, 1, 2>, dim3(rgrid), dim3(256), 0, st, p, tpc)); break;
    }
    MTG_LAUNCH_CHECK();
    return MTG_OK;
  }
  if (a.k == 3) dw_wgrad_kernel<3><<<grid, 256, 0, st>>>(p);
  else dw_wgrad_kernel<5><<<grid, 256, 0, st>>>(p);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_stem_wgrad(const float* x, const bf16* dz, float* dw, int B, int H, int W, cudaStream_t st) {
  MTG_REQUIRE(x && dz && dw, MTG_ERR_ARG, "stem_wgrad: null pointer");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  MTG_REQUIRE(total < (1LL << 31) - 64 * 1024, MTG_ERR_UNSUPPORTED, "stem_wgrad: %lld output pixels exceed the 32-bit index range", total);
  int ctas = 148 * 2;  // two CTAs of 9 warps per SM
  if (ctas > B * Ho) ctas = B * Ho;
  MTG_CUDA(launch_pdl(stem_wgrad_kernel, dim3(static_cast<unsigned>(ctas)), dim3(288), 0, st, x, dz, dw, B, H, W, Ho, Wo));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
