// Small backward kernels of the training step: squeeze-excite, head tail (bilinear transposes, classifiers,
// scale branch), elementwise helpers and the fused multi-tensor AdamW.
// Everything here moves little data (pooled vectors, 2-channel maps) or is a single streaming pass.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "ops.h"

namespace mtgseg {

int group_vectors(int CV);  // dwconv.cu

namespace {

// ---------------------------------------------------------------------------------------------------------
// ds[n][c] partials = sum_p a[n,p,c] * b[n,p,c]   (grid: chunks, groups, B) -> out [B][chunks][C]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dot_pool_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, float* __restrict__ out,
                                                       int HW, int C, int CV, int CVc, int PL, int rows_per_chunk, int chunks) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256 * 8];
  const int vl = threadIdx.x % CVc, pl = threadIdx.x / CVc;
  const int v = blockIdx.y * CVc + vl;
  const bool active = pl < PL && v < CV;
  const int c0 = (active ? v : 0) * 8, n = blockIdx.z, chunk = blockIdx.x;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (active) {
    const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
    const size_t img = static_cast<size_t>(n) * HW * C + c0;
    for (int r = r0 + pl; r < r1; r += PL) {
      float fa[8], fb[8];
      unpack8(ldg16(a + img + static_cast<size_t>(r) * C), fa);
      unpack8(ldg16(b + img + static_cast<size_t>(r) * C), fb);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(fa[j], fb[j], acc[j]);
    }
  }
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(pl * CVc + vl) * 8 + j] = active ? acc[j] : 0.f;
  }
  __syncthreads();
  const int cw = CVc * 8;
  for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
    const int c = blockIdx.y * cw + cl;
    if (c < C) {
      float s = 0.f;
      for (int r = 0; r < PL; ++r) s += red[r * cw + cl];
      out[(static_cast<size_t>(n) * chunks + chunk) * C + c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// SE MLP backward.  Forward: mean -> hid = relu(W1 mean + b1) -> s = hsig(W2 hid + b2) (or, single layer: s = sigmoid(W1 mean)).
// In: ds partial sums.  Out: dpre2[n][C], dpre1[n][SQ], dmean[n][C].  Two launches, each a (64 outputs x image) grid so that
// B = 32 fills the machine (one CTA per image left 116 SMs idle and walked 920 KB of weights alone: 32 us per block):
//   hidden: dpre2 = hsig'(s) * ds ;  dpre1[j] = relu'(hid[j]) * sum_c W2[c][j] dpre2[c]
//   mean  : dmean[c] = sum_j W1[j][c] dpre1[j]           (single-layer form: dpre1[j] = s (1 - s) ds[j] first)
// Threads are (output, reduction group) pairs: a warp reads 32 consecutive outputs of one weight row (coalesced), the four
// groups split the reduction range and meet in shared memory in fixed order.
// ---------------------------------------------------------------------------------------------------------
struct SeBwdP {
  const float* ds_partial; int chunks;
  const float* s; const float* hid;
  const float* w1; const float* w2;  // fp32 master weights: w1 [SQ][C], w2 [C][SQ]
  float* dpre2; float* dpre1; float* dmean;
  int C, SQ, single;                 // single: s = sigmoid(W1 mean), C = pooled channels, SQ = outputs
};
constexpr int SE_TILE = 64;  // outputs per CTA

// out[o0 + o] = sum_r w[r * ld + o0 + o] * x[r] for o < 64, all 256 threads; result valid in threads 0..63
__device__ __forceinline__ float tile_matvec(const float* __restrict__ w, int ld, int o0, int O, const float* x, int R, float* scratch) {
  const int o = threadIdx.x & 63, grp = threadIdx.x >> 6;
  float a = 0.f;
  if (o0 + o < O) {
    const float* wp = w + o0 + o;
    int r = grp;
    for (; r + 12 < R; r += 16) {  // four independent loads in flight per thread
      const float w0 = __ldg(wp + static_cast<size_t>(r) * ld), w1 = __ldg(wp + static_cast<size_t>(r + 4) * ld);
      const float w2 = __ldg(wp + static_cast<size_t>(r + 8) * ld), w3 = __ldg(wp + static_cast<size_t>(r + 12) * ld);
      a = fmaf(w0, x[r], a); a = fmaf(w1, x[r + 4], a); a = fmaf(w2, x[r + 8], a); a = fmaf(w3, x[r + 12], a);
    }
    for (; r < R; r += 4) a = fmaf(__ldg(wp + static_cast<size_t>(r) * ld), x[r], a);
  }
  scratch[grp * 64 + o] = a;
  __syncthreads();
  return scratch[o] + scratch[64 + o] + scratch[128 + o] + scratch[192 + o];
}

// grid (ceil(SQ / 64), B): two-layer form only
__global__ void __launch_bounds__(256) se_bwd_hidden_kernel(const SeBwdP p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* dp2 = sm;            // [C]
  float* scratch = sm + p.C;  // [256]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float d = 0.f;
    for (int k = 0; k < p.chunks; ++k) d += p.ds_partial[(static_cast<size_t>(n) * p.chunks + k) * p.C + c];
    const float sv = p.s[static_cast<size_t>(n) * p.C + c];
    d = (sv > 0.f && sv < 1.f) ? d * (1.f / 6.f) : 0.f;  // hardsigmoid'
    dp2[c] = d;
    if (blockIdx.x == 0) p.dpre2[static_cast<size_t>(n) * p.C + c] = d;
  }
  __syncthreads();
  const int j0 = blockIdx.x * SE_TILE;
  const float v = tile_matvec(p.w2, p.SQ, j0, p.SQ, dp2, p.C, scratch);  // dhid[j] = sum_c w2[c][j] dpre2[c]
  const int j = j0 + threadIdx.x;
  if (threadIdx.x < SE_TILE && j < p.SQ)
    p.dpre1[static_cast<size_t>(n) * p.SQ + j] = p.hid[static_cast<size_t>(n) * p.SQ + j] > 0.f ? v : 0.f;  // relu'
}

// grid (ceil(C / 64), B)
__global__ void __launch_bounds__(256) se_bwd_mean_kernel(const SeBwdP p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  float* dp1 = sm;             // [SQ]
  float* scratch = sm + p.SQ;  // [256]
  const int n = blockIdx.y;
  for (int j = threadIdx.x; j < p.SQ; j += blockDim.x) {
    float d;
    if (p.single) {  // s[n][j] = sigmoid(sum_c w1[j][c] mean[c]); ds given per output j
      d = 0.f;
      for (int k = 0; k < p.chunks; ++k) d += p.ds_partial[(static_cast<size_t>(n) * p.chunks + k) * p.SQ + j];
      const float sv = p.s[static_cast<size_t>(n) * p.SQ + j];
      d *= sv * (1.f - sv);
      if (blockIdx.x == 0) p.dpre1[static_cast<size_t>(n) * p.SQ + j] = d;
    } else {
      d = p.dpre1[static_cast<size_t>(n) * p.SQ + j];
    }
    dp1[j] = d;
  }
  __syncthreads();
  const int c0 = blockIdx.x * SE_TILE;
  const float v = tile_matvec(p.w1, p.C, c0, p.C, dp1, p.SQ, scratch);  // dmean[c] = sum_j w1[j][c] dpre1[j]
  const int c = c0 + threadIdx.x;
  if (threadIdx.x < SE_TILE && c < p.C) p.dmean[static_cast<size_t>(n) * p.C + c] = v;
}

// dW[i][j] = sum_n u[n][i] * v[n][j] * vscale ; optional dbias[i] = sum_n u[n][i].   v may be chunked partial sums.
// CTA tile: 16 rows (i) x 64 columns (j); u and the chunk-reduced v of up to 64 images are staged in shared memory.
__global__ void __launch_bounds__(256) outer_sum_kernel(const float* __restrict__ u, const float* __restrict__ v, int v_chunks, float vscale,
                                                        float* __restrict__ dw, float* __restrict__ dbias, int B, int I, int J) {
  pdl_trigger();
  pdl_wait();
  __shared__ float us[64][16];
  __shared__ float vs[64][64];
  const int j0 = blockIdx.x * 64, i0 = blockIdx.y * 16;
  const int tj = threadIdx.x & 63, ti = threadIdx.x >> 6;  // 4 row lanes
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int nb = 0; nb < B; nb += 64) {
    const int nn = min(64, B - nb);
    for (int idx = threadIdx.x; idx < nn * 64; idx += 256) {
      const int n = idx >> 6, j = j0 + (idx & 63);
      float t = 0.f;
      if (j < J)
        for (int k = 0; k < v_chunks; ++k) t += v[(static_cast<size_t>(nb + n) * v_chunks + k) * J + j];
      vs[n][idx & 63] = t * vscale;
    }
    for (int idx = threadIdx.x; idx < nn * 16; idx += 256) {
      const int n = idx >> 4, i = i0 + (idx & 15);
      us[n][idx & 15] = i < I ? u[static_cast<size_t>(nb + n) * I + i] : 0.f;
    }
    __syncthreads();
    for (int n = 0; n < nn; ++n) {
      const float vv = vs[n][tj];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float uu = us[n][ti + 4 * m];
        acc[m] = fmaf(uu, vv, acc[m]);
        bsum[m] += uu;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int i = i0 + ti + 4 * m, j = j0 + tj;
    if (i < I && j < J) dw[static_cast<size_t>(i) * J + j] = acc[m];
    if (dbias && blockIdx.x == 0 && tj == 0 && i < I) dbias[i] = bsum[m];
  }
}

// ---------------------------------------------------------------------------------------------------------
// transpose of the bilinear upsample (align_corners=False): dlow[n,q,c] = sum_pix w(pix -> q) dhigh_res[n,c,pix]
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

// g: gradient at the fine resolution, element (n, c, y, x) at g[n*sn + c*sc + (y*Wf + x)*sp]; out fp32 [n][q][c]
template <typename T>
__global__ void __launch_bounds__(128) upsample_bwd_kernel(const T* __restrict__ g, float* __restrict__ out, int B, int NC, int Hc, int Wc,
                                                           int Hf, int Wf, long long sn, long long sc, long long sp) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Hc * Wc * NC) return;
  const int c = idx % NC;
  int t = idx / NC;
  const int qx = t % Wc; t /= Wc;
  const int qy = t % Hc;
  const int n = t / Hc;
  const float sy = static_cast<float>(Hc) / Hf, sx = static_cast<float>(Wc) / Wf;
  // fine rows whose source interval can touch qy: src in (qy-1, qy+1)
  const int y_lo = max(0, static_cast<int>(floorf((qy - 1 + 0.5f) / sy - 0.5f)) - 1);
  const int y_hi = min(Hf - 1, static_cast<int>(ceilf((qy + 1 + 0.5f) / sy - 0.5f)) + 1);
  const int x_lo = max(0, static_cast<int>(floorf((qx - 1 + 0.5f) / sx - 0.5f)) - 1);
  const int x_hi = min(Wf - 1, static_cast<int>(ceilf((qx + 1 + 0.5f) / sx - 0.5f)) + 1);
  const T* gp = g + n * sn + c * sc;
  float acc = 0.f;
  for (int y = y_lo; y <= y_hi; ++y) {
    int y0, y1; float ly;
    src_index(y, sy, Hc, y0, y1, ly);
    const float wy = (y0 == qy ? 1.f - ly : 0.f) + (y1 == qy ? ly : 0.f);
    if (wy == 0.f) continue;
    float row = 0.f;
    for (int x = x_lo; x <= x_hi; ++x) {
      int x0, x1; float lx;
      src_index(x, sx, Wc, x0, x1, lx);
      const float wx = (x0 == qx ? 1.f - lx : 0.f) + (x1 == qx ? lx : 0.f);
      if (wx != 0.f) row = fmaf(wx, ldf(gp + (static_cast<long long>(y) * Wf + x) * sp), row);
    }
    acc = fmaf(wy, row, acc);
  }
  out[idx] = acc;
}

// ---------------------------------------------------------------------------------------------------------
// head tail backward, one CTA per image (see tail.cu for the forward):
//   in : dlow_logits do[n,q,c] (40x30), dh2[n,p,c] (20x15, = up2^T(do)), cbr, s, low
//   out: dcbr[n,p,i] = s[n,i] * sum_c dh2 Wh[c,i] ; ds partial[n][i] = sum_p (sum_c dh2 Wh[c,i]) cbr ;
//        dlow[n,q,k] = sum_c do Wl[c,k] ; atomics into dWh[c,i], dWl[c,k], dbias[c] (both biases get the same gradient)
// ---------------------------------------------------------------------------------------------------------
struct HeadBwdP {
  const float* d_o; const float* dh2; const bf16* cbr; const float* s; const bf16* low;
  const float* w_high; const float* w_low;
  bf16* dcbr; float* ds; bf16* dlow; float* dw_high; float* dw_low; float* db_high; float* db_low;
  int Hh, Wh, Hl, Wl, IC, LC, NC;
};
constexpr int MAX_NC = 8;
// grid (B, segments): a CTA owns 1/segments of the image's pixels (one CTA per image: 32 CTAs walking 300 + 1200 pixels serially, 171 us)
__global__ void __launch_bounds__(256) head_bwd_kernel(const HeadBwdP p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_ds[256];
  const int n = blockIdx.x;
  const int nh = p.Hh * p.Wh, nl = p.Hl * p.Wl;
  const int seg = blockIdx.y, nseg = gridDim.y;
  const int h_lo = static_cast<int>(static_cast<long long>(nh) * seg / nseg), h_hi = static_cast<int>(static_cast<long long>(nh) * (seg + 1) / nseg);
  const int l_lo = static_cast<int>(static_cast<long long>(nl) * seg / nseg), l_hi = static_cast<int>(static_cast<long long>(nl) * (seg + 1) / nseg);
  // part 1: thread (i, grp) owns inter channel i over every (blockDim/IC)-th high-res pixel
  const int g1 = max(1, static_cast<int>(blockDim.x) / p.IC);
  for (int i = threadIdx.x % p.IC, grp = threadIdx.x / p.IC; grp < g1; grp = g1) {
    const float sv = p.s[static_cast<size_t>(n) * p.IC + i];
    float wh[MAX_NC], dwh[MAX_NC];
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c) { wh[c] = c < p.NC ? p.w_high[c * p.IC + i] : 0.f; dwh[c] = 0.f; }
    float dsv = 0.f;
    for (int px = h_lo + grp; px < h_hi; px += g1) {
      const size_t row = static_cast<size_t>(n) * nh + px;
      const float cv = __bfloat162float(p.cbr[row * p.IC + i]);
      float dt = 0.f;
#pragma unroll
      for (int c = 0; c < MAX_NC; ++c)
        if (c < p.NC) {
          const float g = p.dh2[row * p.NC + c];
          dt = fmaf(g, wh[c], dt);
          dwh[c] = fmaf(g, sv * cv, dwh[c]);
        }
      p.dcbr[row * p.IC + i] = __float2bfloat16(dt * sv);
      dsv = fmaf(dt, cv, dsv);
    }
    s_ds[threadIdx.x] = dsv;  // threadIdx.x == grp * IC + i
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c)
      if (c < p.NC) atomicAdd(p.dw_high + c * p.IC + i, dwh[c]);
  }
  // ds partial of this (image, segment): fixed-order sum over the pixel groups, plain store.  (Atomics here made the whole
  // backward chain irreproducible: ds feeds the activation gradients, and in bf16 one last-bit difference early in the chain
  // decorrelates every later rounding -- two runs then differ by the rounding-noise floor, ~1 % of the gradient norm.)
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < p.IC) {
    float t = 0.f;
    for (int g = 0; g < g1; ++g) t += s_ds[g * p.IC + threadIdx.x];
    p.ds[(static_cast<size_t>(n) * nseg + seg) * p.IC + threadIdx.x] = t;
  }
  // part 2: thread (k, grp) owns low channel k over every (blockDim/LC)-th low-res pixel
  const int g2 = max(1, static_cast<int>(blockDim.x) / p.LC);
  for (int k = threadIdx.x % p.LC, grp = threadIdx.x / p.LC; grp < g2; grp = g2) {
    float wl[MAX_NC], dwl[MAX_NC];
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c) { wl[c] = c < p.NC ? p.w_low[c * p.LC + k] : 0.f; dwl[c] = 0.f; }
    for (int q = l_lo + grp; q < l_hi; q += g2) {
      const size_t row = static_cast<size_t>(n) * nl + q;
      const float lv = __bfloat162float(p.low[row * p.LC + k]);
      float dl = 0.f;
#pragma unroll
      for (int c = 0; c < MAX_NC; ++c)
        if (c < p.NC) {
          const float g = p.d_o[row * p.NC + c];
          dl = fmaf(g, wl[c], dl);
          dwl[c] = fmaf(g, lv, dwl[c]);
        }
      p.dlow[row * p.LC + k] = __float2bfloat16(dl);
    }
#pragma unroll
    for (int c = 0; c < MAX_NC; ++c)
      if (c < p.NC) atomicAdd(p.dw_low + c * p.LC + k, dwl[c]);
  }
  // part 3: bias gradients
  if (threadIdx.x < p.NC) {
    float b = 0.f;
    for (int q = l_lo; q < l_hi; ++q) b += p.d_o[(static_cast<size_t>(n) * nl + q) * p.NC + threadIdx.x];
    atomicAdd(p.db_high + threadIdx.x, b);
    atomicAdd(p.db_low + threadIdx.x, b);
  }
}

__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, size_t nvec) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float fa[8], fb[8];
    unpack8(ldg16(a + i * 8), fa);
    unpack8(ldg16(b + i * 8), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(fa);
  }
}

__global__ void fill_f32_kernel(float* __restrict__ p, float v, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------
// AdamW, all parameter tensors in one launch (torch.optim.AdamW maths, decoupled weight decay):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// grads are multiplied by *inv_scale (GradScaler) when given; the whole step is skipped when *found_inf != 0.
// ---------------------------------------------------------------------------------------------------------
struct AdamChunk { float* p; const float* g; float* m; float* v; int n; int pad; };
__global__ void __launch_bounds__(256) adamw_kernel(const AdamChunk* __restrict__ chunks, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, const float* __restrict__ inv_scale,
                                                    const float* __restrict__ found_inf, const float* __restrict__ hyper) {
  pdl_trigger();
  pdl_wait();
  if (found_inf && *found_inf != 0.f) return;
  if (hyper) {  // captured step: the scalars of THIS replay live in device memory (adamw_hyper_kernel)
    lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3]; wd = hyper[4]; bc1 = hyper[5]; bc2_sqrt = hyper[6];
  }
  const AdamChunk c = chunks[blockIdx.x];
  const float gs = inv_scale ? *inv_scale : 1.f;
  const float step_size = lr / bc1, decay = 1.f - lr * wd;
  auto upd = [&](float& pv, float g, float& m, float& v) {
    g *= gs;
    pv *= decay;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    pv -= step_size * m / (sqrtf(v) / bc2_sqrt + eps);
  };
  const bool aligned = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) | reinterpret_cast<uintptr_t>(c.m) |
                         reinterpret_cast<uintptr_t>(c.v)) & 15) == 0;
  const int n4 = aligned ? c.n / 4 : 0;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(c.p)[i], m4 = reinterpret_cast<float4*>(c.m)[i], v4 = reinterpret_cast<float4*>(c.v)[i];
    const float4 g4 = reinterpret_cast<const float4*>(c.g)[i];
    upd(p4.x, g4.x, m4.x, v4.x); upd(p4.y, g4.y, m4.y, v4.y); upd(p4.z, g4.z, m4.z, v4.z); upd(p4.w, g4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(c.p)[i] = p4; reinterpret_cast<float4*>(c.m)[i] = m4; reinterpret_cast<float4*>(c.v)[i] = v4;
  }
  for (int i = n4 * 4 + threadIdx.x; i < c.n; i += blockDim.x) upd(c.p[i], c.g[i], c.m[i], c.v[i]);
}

__global__ void adamw_hyper_kernel(float* hyper, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  hyper[0] = lr; hyper[1] = b1; hyper[2] = b2; hyper[3] = eps; hyper[4] = wd; hyper[5] = bc1; hyper[6] = bc2_sqrt; hyper[7] = 0.f;
}

}  // namespace

int launch_dot_pool(const bf16* a, const bf16* b, float* out, int B, int HW, int C, int chunks, cudaStream_t st) {
  MTG_REQUIRE(a && b && out && C % 8 == 0, MTG_ERR_ARG, "dot_pool: bad arguments");
  const int CV = C / 8, CVc = group_vectors(CV), PL = 256 / CVc;
  dim3 grid(chunks, ceil_div(CV, CVc), B);
  MTG_CUDA(launch_pdl(dot_pool_kernel, dim3(grid), dim3(256), 0, st, a, b, out, HW, C, CV, CVc, PL, ceil_div(HW, chunks), chunks));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_se_bwd(const SeBwdArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.ds_partial && a.s && a.w1 && a.dpre1 && a.dmean, MTG_ERR_ARG, "se_bwd: null pointer");
  SeBwdP p{a.ds_partial, a.chunks, a.s, a.hid, a.w1, a.w2, a.dpre2, a.dpre1, a.dmean, a.C, a.SQ, a.w2 ? 0 : 1};
  MTG_REQUIRE(a.C <= 4096 && a.SQ <= 4096, MTG_ERR_UNSUPPORTED, "se_bwd: C / SQ above 4096");
  if (a.w2) {
    MTG_REQUIRE(a.hid && a.dpre2, MTG_ERR_ARG, "se_bwd: the two-layer form needs hid and dpre2");
    MTG_CUDA(launch_pdl(se_bwd_hidden_kernel, dim3(dim3(ceil_div(a.SQ, SE_TILE), a.B)), dim3(256), sizeof(float) * (a.C + 256), st, p));
    MTG_LAUNCH_CHECK();
  }
  MTG_CUDA(launch_pdl(se_bwd_mean_kernel, dim3(dim3(ceil_div(a.C, SE_TILE), a.B)), dim3(256), sizeof(float) * (a.SQ + 256), st, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_outer_sum(const float* u, const float* v, int v_chunks, float vscale, float* dw, float* dbias, int B, int I, int J,
                     cudaStream_t st) {
  MTG_REQUIRE(u && v && dw, MTG_ERR_ARG, "outer_sum: null pointer");
  MTG_CUDA(launch_pdl(outer_sum_kernel, dim3(dim3(ceil_div(J, 64), ceil_div(I, 16))), dim3(256), 0, st, u, v, v_chunks, vscale, dw, dbias, B, I, J));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_upsample_bwd(const void* g, int dtype, float* out, int B, int NC, int Hc, int Wc, int Hf, int Wf, long long sn, long long sc,
                        long long sp, cudaStream_t st) {
  MTG_REQUIRE(g && out, MTG_ERR_ARG, "upsample_bwd: null pointer");
  const int total = B * Hc * Wc * NC;
  const int grid = ceil_div(total, 128);
  if (dtype == LOGITS_F32) MTG_CUDA(launch_pdl(upsample_bwd_kernel<float>, dim3(grid), dim3(128), 0, st, static_cast<const float*>(g), out, B, NC, Hc, Wc, Hf, Wf, sn, sc, sp));
  else if (dtype == LOGITS_BF16) MTG_CUDA(launch_pdl(upsample_bwd_kernel<bf16>, dim3(grid), dim3(128), 0, st, static_cast<const bf16*>(g), out, B, NC, Hc, Wc, Hf, Wf, sn, sc, sp));
  else if (dtype == LOGITS_F16) MTG_CUDA(launch_pdl(upsample_bwd_kernel<__half>, dim3(grid), dim3(128), 0, st, static_cast<const __half*>(g), out, B, NC, Hc, Wc, Hf, Wf, sn, sc, sp));
  else MTG_REQUIRE(false, MTG_ERR_ARG, "upsample_bwd: unknown dtype %d", dtype);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int head_bwd_segments(int B) { return B >= 128 ? 2 : 8; }

int launch_head_bwd(const HeadBwdArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.NC >= 1 && a.NC <= MAX_NC, MTG_ERR_UNSUPPORTED, "head_bwd: num_classes out of range");
  HeadBwdP p{a.d_o, a.dh2, a.cbr, a.s, a.low, a.w_high, a.w_low, a.dcbr, a.ds, a.dlow, a.dw_high, a.dw_low, a.db_high, a.db_low,
             a.Hh, a.Wh, a.Hl, a.Wl, a.IC, a.LC, a.NC};
  MTG_REQUIRE(a.IC <= 256, MTG_ERR_UNSUPPORTED, "head_bwd: inter_channels above 256");
  MTG_CUDA(launch_pdl(head_bwd_kernel, dim3(dim3(a.B, head_bwd_segments(a.B))), dim3(256), 0, st, p));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_add_bf16(const bf16* a, const bf16* b, bf16* out, size_t n, cudaStream_t st) {
  MTG_REQUIRE(n % 8 == 0, MTG_ERR_ARG, "add_bf16: n %% 8 != 0");
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  MTG_CUDA(launch_pdl(add_bf16_kernel, dim3(static_cast<unsigned>(blocks ? blocks : 1)), dim3(256), 0, st, a, b, out, n / 8));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_fill_f32(float* p, float v, size_t n, cudaStream_t st) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fill_f32_kernel<<<static_cast<unsigned>(blocks ? blocks : 1), 256, 0, st>>>(p, v, n);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_adamw(const void* chunk_table, int n_chunks, float lr, float b1, float b2, float eps, float wd, int step,
                 const float* inv_scale, const float* found_inf, cudaStream_t st) {
  MTG_REQUIRE(chunk_table && n_chunks > 0 && step > 0, MTG_ERR_ARG, "adamw: bad arguments");
  const float bc1 = 1.f - powf(b1, static_cast<float>(step));
  const float bc2s = sqrtf(1.f - powf(b2, static_cast<float>(step)));
  MTG_CUDA(launch_pdl(adamw_kernel, dim3(n_chunks), dim3(256), 0, st, static_cast<const AdamChunk*>(chunk_table), lr, b1, b2, eps, wd, bc1, bc2s, inv_scale, found_inf,
                                         nullptr));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_adamw_hyper(float* hyper, float lr, float b1, float b2, float eps, float wd, int step, cudaStream_t st) {
  MTG_REQUIRE(hyper && step > 0, MTG_ERR_ARG, "adamw_hyper: bad arguments");
  adamw_hyper_kernel<<<1, 1, 0, st>>>(hyper, lr, b1, b2, eps, wd, 1.f - powf(b1, static_cast<float>(step)),
                                      sqrtf(1.f - powf(b2, static_cast<float>(step))));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_adamw_dev(const void* chunk_table, int n_chunks, const float* hyper, cudaStream_t st) {
  MTG_REQUIRE(chunk_table && n_chunks > 0 && hyper, MTG_ERR_ARG, "adamw_dev: bad arguments");
  MTG_CUDA(launch_pdl(adamw_kernel, dim3(n_chunks), dim3(256), 0, st, static_cast<const AdamChunk*>(chunk_table), 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1.f, nullptr, nullptr, hyper));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
