// Training-mode BatchNorm (+ activation, residual, squeeze-excite hooks) on NHWC bf16, forward and backward.
//
// forward :  z (raw conv output, bf16) --stats--> mean / biased var --finalize--> scale = gamma*rstd,
//            shift = beta - mean*scale, running-stat EMA (unbiased var), num_batches_tracked++
//            y = act(z*scale + shift) (+ residual)            [+ per-chunk channel sums of y for the SE pool]
// backward:  dyh = dy' * act'(z*scale+shift) with dy' = dy (* s[n,c] + dmean[n,c]/HW for SE inputs)
//            reduce: sum dyh, sum dyh*xhat -> dbeta, dgamma ;  dz = scale*(dyh - mean(dyh) - xhat*mean(dyh*xhat))
// Semantics: nn.BatchNorm2d in train mode (tv:models/mobilenetv3.py:155 eps 1e-3 / momentum 1e-2 for the backbone,
// train/model.py:111 defaults for the head), nn.Hardswish / nn.ReLU derivatives as in SURVEY.md App. B.
// Statistics are per-channel fp64 accumulators ([2][C]: sum, sum of squares; or sum dyh, sum dyh*xhat): CTAs reduce their rows
// in fixed order in fp32 and add ONE fp64 atomic per channel (the producing conv kernels do the same from their epilogues, so a
// layer needs no statistics pass at all).  There is no finalize launch: every CTA of the elementwise pass derives scale / shift
// (or the two backward means) for its own channels from the accumulators; the CTA (chunk 0, image 0) of every channel group
// also writes the saved statistics, the running-stat EMA and dgamma / dbeta.  (fp64 sums of fp32 partials of similar magnitude
// are exact, so the result does not depend on the order in which the atomics land.)
#include "ops.h"

namespace mtgseg {

int group_vectors(int CV);  // dwconv.cu

namespace {

__device__ __forceinline__ float act_grad(float zh, int act) {
  switch (act) {
    case ACT_RELU: return zh > 0.f ? 1.f : 0.f;
    case ACT_HSWISH: return zh <= -3.f ? 0.f : (zh >= 3.f ? 1.f : zh * (1.f / 3.f) + 0.5f);
    default: return 1.f;
  }
}

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

constexpr int kRowsInFlight = 4;  // rows (16-byte vectors per operand) a thread loads before consuming them

struct Geo {  // thread -> (channel vector, row lane) mapping shared by all kernels here
  int C, CV, CVc, PL, HW, rows_per_chunk, chunks;
  int B, ipc;  // batch; images per CTA of the REDUCING kernels (1 for the elementwise ones)
};

struct Lane {
  bool active; int c0, pl, vl;
};
__device__ __forceinline__ Lane lane_of(const Geo& g) {
  Lane l;
  l.vl = threadIdx.x % g.CVc;
  l.pl = threadIdx.x / g.CVc;
  const int v = blockIdx.y * g.CVc + l.vl;
  l.active = l.pl < g.PL && v < g.CV;
  l.c0 = (l.active ? v : 0) * 8;
  return l;
}

// fixed-order reduction of K per-thread 8-vectors over the row lanes; result for channel c (of this CTA) in out[k]
template <int K>
__device__ __forceinline__ void block_reduce_store(const Geo& g, const Lane& l, float (&vals)[K][8], float* red,
                                                   float* const (&dst)[K], size_t dst_index_base) {
  const int cw = g.CVc * 8;
  for (int k = 0; k < K; ++k) {
    __syncthreads();
    if (l.pl < g.PL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(l.pl * g.CVc + l.vl) * 8 + j] = l.active ? vals[k][j] : 0.f;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < g.C) {
        float s = 0.f;
        for (int r = 0; r < g.PL; ++r) s += red[r * cw + cl];
        dst[k][dst_index_base + c] = s;
      }
    }
  }
}

// same reduction, result ADDED to the fp64 accumulators dst[k * C + c]
template <int K>
__device__ __forceinline__ void block_reduce_atomic(const Geo& g, const Lane& l, float (&vals)[K][8], float* red, double* dst) {
  const int cw = g.CVc * 8;
  for (int k = 0; k < K; ++k) {
    __syncthreads();
    if (l.pl < g.PL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(l.pl * g.CVc + l.vl) * 8 + j] = l.active ? vals[k][j] : 0.f;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < g.C) {
        float s = 0.f;
        for (int r = 0; r < g.PL; ++r) s += red[r * cw + cl];
        atomicAdd(dst + static_cast<size_t>(k) * g.C + c, static_cast<double>(s));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
// grid (chunks, groups, ceil(B / ipc)): stat[c] += sum z, stat[C + c] += sum z^2 over the chunk's rows of the CTA's ipc images.
// Only for layers whose producer does not deliver the statistics from its epilogue (the stem, the per-op ABI entry).
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ z, double* __restrict__ stat, const Geo g) {
  __shared__ float red[256 * 8];
  pdl_trigger();
  pdl_wait();
  const Lane l = lane_of(g);
  const int chunk = blockIdx.x;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (l.active) {
    const int r0 = chunk * g.rows_per_chunk, r1 = min(g.HW, r0 + g.rows_per_chunk);
    const int n0 = blockIdx.z * g.ipc, n1 = min(g.B, n0 + g.ipc);
    for (int n = n0; n < n1; ++n) {
      const bf16* base = z + static_cast<size_t>(n) * g.HW * g.C + l.c0;
      for (int r = r0 + l.pl; r < r1; r += g.PL) {
        float f[8];
        unpack8(ldg16(base + static_cast<size_t>(r) * g.C), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] += f[j]; acc[1][j] = fmaf(f[j], f[j], acc[1][j]); }
      }
    }
  }
  block_reduce_atomic<2>(g, l, acc, red, stat);
}

struct BnApplyP {
  const bf16* z; int act; const bf16* residual; bf16* y; float* gap;
  // finalize (folded in): batch statistics -> scale / shift for this CTA's channels; the CTA (chunk 0, image 0) of every channel
  // group also saves them and updates the running statistics
  const double* stat; double count, inv_count;
  const float* gamma; const float* beta; float eps, momentum;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float* scale; float* shift; float* save_mean; float* save_rstd;
};
// grid (chunks, groups, B).  kRows: rows (16-byte loads per operand) a thread has in flight; layers without a residual take 8
// (one operand stream: 4 rows leave too few bytes in flight per SM to cover the HBM latency at 3 CTAs per SM)
template <int kRows, bool kRes>
__global__ void __launch_bounds__(256, kRes ? 2 : 3) bn_apply_kernel(const BnApplyP p, const Geo g) {
  __shared__ float red[256 * 8];
  __shared__ float s_sc[128], s_sh[128];
  pdl_trigger();
  const Lane l = lane_of(g);
  const int chunk = blockIdx.x;
  const bool writer = chunk == 0 && blockIdx.z == 0;
  pdl_wait();
  if (writer && blockIdx.y == 0 && threadIdx.x == 0 && p.num_batches_tracked) *p.num_batches_tracked += 1;
  // scale / shift of this CTA's channels: ONE thread per channel does the fp64 arithmetic (the fp64 issue rate is 1/64 of
  // fp32 here: done by every thread for its 8 channels it cost more than the whole streaming loop of a small layer)
  {
    const int cw = g.CVc * 8;
    for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < g.C) {
        // fp64 only where cancellation needs it (E[z^2] - mean^2): three multiplies and one FMA; the reciprocal square root is the
        // correctly rounded fp32 one (fp64 division / sqrt issue at 1/64 rate: with one prologue per CTA they cost ~15 % of the
        // pass at B = 256)
        const double mean = p.stat[c] * p.inv_count;
        double var = fma(-mean, mean, p.stat[g.C + c] * p.inv_count);
        if (var < 0.0) var = 0.0;
        const float rstd = 1.0f / sqrtf(static_cast<float>(var) + p.eps);
        const float sc = __ldg(p.gamma + c) * rstd;
        const float sh = __ldg(p.beta + c) - static_cast<float>(mean) * sc;
        s_sc[cl] = sc;
        s_sh[cl] = sh;
        if (writer) {
          p.scale[c] = sc;
          p.shift[c] = sh;
          p.save_mean[c] = static_cast<float>(mean);
          p.save_rstd[c] = rstd;
          if (p.running_mean) {
            const double unbiased = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
            p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * static_cast<float>(mean);
            p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * static_cast<float>(unbiased);
          }
        }
      }
    }
    __syncthreads();
  }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = l.active ? s_sc[l.vl * 8 + j] : 0.f; sh[j] = l.active ? s_sh[l.vl * 8 + j] : 0.f; }
  const int r0 = chunk * g.rows_per_chunk, r1 = min(g.HW, r0 + g.rows_per_chunk);
  // a CTA walks g.ipc images (large batches: fewer, longer CTAs amortise the prologue above)
  for (int n = blockIdx.z * g.ipc; n < min(g.B, (static_cast<int>(blockIdx.z) + 1) * g.ipc); ++n) {
    float gsum[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gsum[0][j] = 0.f;
    if (l.active) {
      const size_t img = static_cast<size_t>(n) * g.HW * g.C + l.c0;
      // kRowsInFlight independent 16-byte loads per thread before the first use: the loop is latency bound otherwise (the store
      // in the body keeps the compiler from hoisting the next row's load)
      for (int r = r0 + l.pl; r < r1; r += kRows * g.PL) {
        uint4 zq[kRows], rq[kRes ? kRows : 1];
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int rr = r + u * g.PL;
          const size_t off = img + static_cast<size_t>(rr < r1 ? rr : r) * g.C;
          zq[u] = ldg16(p.z + off);
          if (kRes) rq[u] = ldg16(p.residual + off);
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const int rr = r + u * g.PL;
          if (rr >= r1) break;
          const size_t off = img + static_cast<size_t>(rr) * g.C;
          float f[8];
          unpack8(zq[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = apply_act(fmaf(f[j], sc[j], sh[j]), p.act);
          if (kRes) {
            float rf[8];
            unpack8(rq[kRes ? u : 0], rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += rf[j];
          }
          const uint4 q = pack8(f);
          *reinterpret_cast<uint4*>(p.y + off) = q;
          if (p.gap) {
            float rf[8];
            unpack8(q, rf);
#pragma unroll
            for (int j = 0; j < 8; ++j) gsum[0][j] += rf[j];
          }
        }
      }
    }
    if (p.gap) {
      float* const dst[1] = {p.gap};
      block_reduce_store<1>(g, l, gsum, red, dst, (static_cast<size_t>(n) * g.chunks + chunk) * g.C);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
struct BnBwdP {
  const bf16* z; const bf16* dy;            // raw conv output, incoming gradient w.r.t. y (or w.r.t. s*y for SE inputs)
  const float* scale; const float* shift;   // forward scale/shift (gamma*rstd, beta - mean*scale)
  const float* mean; const float* rstd;     // saved batch statistics
  int act;
  const float* se_s; const float* se_dmean; float inv_hw;  // optional: dy' = dy*se_s[n,c] + se_dmean[n,c]*inv_hw
  double* bstat; double inv_count;          // [2][C]: sum dyh, sum dyh*xhat (reduce adds, apply reads); 1 / (B*HW)
  float* dgamma; float* dbeta;              // (apply, designated CTA) = the two sums
  bf16* dz;                                 // (apply)
};

__device__ __forceinline__ void bwd_terms(const BnBwdP& p, const uint4& zq, const uint4& dq, const float (&sc)[8],
                                          const float (&sh)[8], const float (&mu)[8], const float (&rs)[8], const float (&ses)[8],
                                          const float (&sed)[8], float (&dyh)[8], float (&xh)[8]) {
  float zf[8], df[8];
  unpack8(zq, zf);
  unpack8(dq, df);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d = p.se_s ? fmaf(df[j], ses[j], sed[j]) : df[j];
    dyh[j] = d * act_grad(fmaf(zf[j], sc[j], sh[j]), p.act);
    xh[j] = (zf[j] - mu[j]) * rs[j];
  }
}

template <bool kApply>
__global__ void __launch_bounds__(256) bn_bwd_kernel(const BnBwdP p, const Geo g) {
  __shared__ float red[256 * 8];
  __shared__ float s_c1[kApply ? 128 : 1], s_c2[kApply ? 128 : 1];
  pdl_trigger();
  const Lane l = lane_of(g);
  const int chunk = blockIdx.x;
  pdl_wait();
  if (kApply) {  // mean(dyh), mean(dyh*xhat) of this CTA's channels: one thread per channel does the fp64 part (see bn_apply_kernel)
    const int cw = g.CVc * 8;
    const bool writer = chunk == 0 && blockIdx.z == 0;
    for (int cl = threadIdx.x; cl < cw; cl += blockDim.x) {
      const int c = blockIdx.y * cw + cl;
      if (c < g.C) {
        const double s = p.bstat[c], sx = p.bstat[g.C + c];
        s_c1[cl] = static_cast<float>(s * p.inv_count);
        s_c2[cl] = static_cast<float>(sx * p.inv_count);
        if (writer) { p.dbeta[c] = static_cast<float>(s); p.dgamma[c] = static_cast<float>(sx); }
      }
    }
    __syncthreads();
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  if (l.active) {
    float sc[8], sh[8], mu[8], rs[8], ses[8], sed[8], c1[8], c2[8];
    load8f(p.scale + l.c0, sc); load8f(p.shift + l.c0, sh); load8f(p.mean + l.c0, mu); load8f(p.rstd + l.c0, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ses[j] = 1.f; sed[j] = 0.f; c1[j] = c2[j] = 0.f; }
    if (kApply) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { c1[j] = s_c1[l.vl * 8 + j]; c2[j] = s_c2[l.vl * 8 + j]; }
    }
    const int r0 = chunk * g.rows_per_chunk, r1 = min(g.HW, r0 + g.rows_per_chunk);
    const int n0 = blockIdx.z * g.ipc, n1 = min(g.B, n0 + g.ipc);  // the reduce pass walks ipc images per CTA (fewer atomics)
    for (int n = n0; n < n1; ++n) {
      if (p.se_s) {
        load8f(p.se_s + static_cast<size_t>(n) * g.C + l.c0, ses);
        load8f(p.se_dmean + static_cast<size_t>(n) * g.C + l.c0, sed);
#pragma unroll
        for (int j = 0; j < 8; ++j) sed[j] *= p.inv_hw;
      }
      const size_t img = static_cast<size_t>(n) * g.HW * g.C + l.c0;
      for (int r = r0 + l.pl; r < r1; r += kRowsInFlight * g.PL) {  // loads of kRowsInFlight rows first, see bn_apply_kernel
        uint4 zq[kRowsInFlight], dq[kRowsInFlight];
#pragma unroll
        for (int u = 0; u < kRowsInFlight; ++u) {
          const int rr = r + u * g.PL;
          const size_t off = img + static_cast<size_t>(rr < r1 ? rr : r) * g.C;
          zq[u] = ldg16(p.z + off);
          dq[u] = ldg16(p.dy + off);
        }
#pragma unroll
        for (int u = 0; u < kRowsInFlight; ++u) {
          const int rr = r + u * g.PL;
          if (rr >= r1) break;
          float dyh[8], xh[8];
          bwd_terms(p, zq[u], dq[u], sc, sh, mu, rs, ses, sed, dyh, xh);
          if (kApply) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = sc[j] * (dyh[j] - c1[j] - xh[j] * c2[j]);
            *reinterpret_cast<uint4*>(p.dz + img + static_cast<size_t>(rr) * g.C) = pack8(o);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[0][j] += dyh[j]; acc[1][j] = fmaf(dyh[j], xh[j], acc[1][j]); }
          }
        }
      }
    }
  }
  if (!kApply) block_reduce_atomic<2>(g, l, acc, red, p.bstat);
}

Geo make_geo(int C, int HW, int want_chunks, int B) {
  Geo g{};
  g.C = C; g.CV = C / 8; g.CVc = group_vectors(g.CV); g.PL = 256 / g.CVc; g.HW = HW;
  g.chunks = want_chunks;
  g.rows_per_chunk = ceil_div(HW, g.chunks);
  g.B = B; g.ipc = 1;
  return g;
}
// the elementwise passes: at most ~2048 CTAs per channel group (each CTA pays a statistics prologue)
Geo elementwise_geo(Geo g) {
  g.ipc = ceil_div(g.B * g.chunks, 2048);
  return g;
}
// the reducing passes: at most ~512 CTAs (= fp64 atomics) per channel
Geo reducing_geo(Geo g) {
  g.ipc = ceil_div(g.B * g.chunks, 512);
  return g;
}

}  // namespace

// chunks per image for the BN kernels: ~16 rows per thread, at most 32 per image
int bn_chunks(int HW, int C) {
  const int PL = 256 / group_vectors(C / 8);
  int ch = ceil_div(HW, PL * 16);
  if (ch > 32) ch = 32;
  if (ch < 1) ch = 1;
  return ch;
}
size_t bn_partial_floats(int B, int HW, int C) { (void)B; (void)HW; return static_cast<size_t>(8) * C; }  // 2 x [2][C] fp64 accumulators

int launch_bn_train_fwd(const BnTrainFwdArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.z && a.y && a.gamma && a.beta && a.scale && a.shift && a.save_mean && a.save_rstd && a.stat, MTG_ERR_ARG,
              "bn_train_fwd: null pointer");
  MTG_REQUIRE(a.C % 8 == 0, MTG_ERR_UNSUPPORTED, "bn_train_fwd: C %% 8 != 0");
  const Geo g = make_geo(a.C, a.HW, bn_chunks(a.HW, a.C), a.B), gr = reducing_geo(g);
  dim3 grid(g.chunks, ceil_div(g.CV, g.CVc), a.B);
  if (!a.stats_done) {  // the producer did not accumulate the statistics from its epilogue
    MTG_CUDA(cudaMemsetAsync(a.stat, 0, sizeof(double) * 2 * a.C, st));
    const dim3 rgrid(g.chunks, grid.y, ceil_div(a.B, gr.ipc));
    MTG_CUDA(launch_pdl(bn_stats_kernel, rgrid, dim3(256), 0, st, a.z, a.stat, gr));
    MTG_LAUNCH_CHECK();
  }
  Geo ga = g;
  if (a.gap) {  // the SE pool wants few partials per image
    ga = make_geo(a.C, a.HW, a.gap_chunks, a.B);
    grid = dim3(ga.chunks, ceil_div(ga.CV, ga.CVc), a.B);
  }
  BnApplyP ap{a.z, a.act, a.residual, a.y, a.gap, a.stat, static_cast<double>(a.B) * a.HW, 1.0 / (static_cast<double>(a.B) * a.HW), a.gamma, a.beta, a.eps, a.momentum,
              a.running_mean, a.running_var, a.num_batches_tracked, a.scale, a.shift, a.save_mean, a.save_rstd};
  ga = elementwise_geo(ga);
  grid.z = ceil_div(a.B, ga.ipc);
  if (a.residual) MTG_CUDA(launch_pdl(bn_apply_kernel<kRowsInFlight, true>, grid, dim3(256), 0, st, ap, ga));
  else MTG_CUDA(launch_pdl(bn_apply_kernel<8, false>, grid, dim3(256), 0, st, ap, ga));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int launch_bn_train_bwd(const BnTrainBwdArgs& a, cudaStream_t st) {
  MTG_REQUIRE(a.z && a.dy && a.dz && a.scale && a.shift && a.save_mean && a.save_rstd && a.bstat && a.dgamma && a.dbeta, MTG_ERR_ARG,
              "bn_train_bwd: null pointer");
  const Geo g = make_geo(a.C, a.HW, bn_chunks(a.HW, a.C), a.B), gr = reducing_geo(g);
  dim3 grid(g.chunks, ceil_div(g.CV, g.CVc), a.B);
  const dim3 rgrid(g.chunks, grid.y, ceil_div(a.B, gr.ipc));
  BnBwdP p{a.z, a.dy, a.scale, a.shift, a.save_mean, a.save_rstd, a.act, a.se_s, a.se_dmean, 1.f / static_cast<float>(a.HW),
           a.bstat, 1.0 / (static_cast<double>(a.B) * a.HW), a.dgamma, a.dbeta, a.dz};
  if (!a.bstat_zeroed) MTG_CUDA(cudaMemsetAsync(a.bstat, 0, sizeof(double) * 2 * a.C, st));
  MTG_CUDA(launch_pdl(bn_bwd_kernel<false>, rgrid, dim3(256), 0, st, p, gr));
  MTG_LAUNCH_CHECK();
  const Geo ge = elementwise_geo(g);
  grid.z = ceil_div(a.B, ge.ipc);
  MTG_CUDA(launch_pdl(bn_bwd_kernel<true>, grid, dim3(256), 0, st, p, ge));
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

}  // namespace mtgseg
