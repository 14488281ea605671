// Weight packing: reference state_dict tensors (fp32, OIHW) -> the layouts the kernels read.
// Runs once per weight update; bandwidth is irrelevant (16 MB), so these are plain grid-stride kernels.
// A training step repacks after every optimizer update: ~160 tiny launches (0.49 ms of a 9.7 ms step in the ncu launch
// list).  Between pack_batch_begin() and pack_batch_flush() the launchers below only RECORD their job; flush runs all of
// them in ONE kernel: the job table travels as a __grid_constant__ kernel parameter (<= 256 jobs x 88 B, CUDA 12 allows 32 KB
// of parameters), so there is no upload and nothing that could synchronise the stream with the host.
#include <vector>

#include "ops.h"

namespace mtgseg {
namespace {

enum PackType : int { PK_CAST = 0, PK_COPY, PK_OTAPI, PK_DW, PK_STEM, PK_FOLD, PK_TRANSPOSE, PK_DGRAD3, PK_BLOCKDIAG };
struct PackJob {
  int type, a, b, c;  // dims (meaning per type)
  float eps;
  int first_block, nblocks;
  unsigned long long n;  // elements
  const void* in0; const void* in1; const void* in2; const void* in3;
  void* out0; void* out1;
};
constexpr int PACK_MAX_JOBS = 256;
struct PackTable {
  int njobs;
  PackJob jobs[PACK_MAX_JOBS];
};
static_assert(sizeof(PackTable) <= 32 * 1024 - 64, "kernel parameter space");
thread_local std::vector<PackJob>* g_batch = nullptr;

__global__ void __launch_bounds__(256) pack_batch_kernel(const __grid_constant__ PackTable t) {
  // the job of this CTA: last job whose first_block <= blockIdx.x (first_block is increasing)
  int lo = 0, hi = t.njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t.jobs[mid].first_block <= static_cast<int>(blockIdx.x)) lo = mid;
    else hi = mid - 1;
  }
  const PackJob& j = t.jobs[lo];
  const size_t start = static_cast<size_t>(blockIdx.x - j.first_block) * 256 + threadIdx.x, stride = static_cast<size_t>(j.nblocks) * 256;
  const float* in = static_cast<const float*>(j.in0);
  switch (j.type) {
    case PK_CAST: {
      bf16* out = static_cast<bf16*>(j.out0);
      for (size_t i = start; i < j.n; i += stride) out[i] = __float2bfloat16(in[i]);
    } break;
    case PK_COPY: {
      float* out = static_cast<float*>(j.out0);
      for (size_t i = start; i < j.n; i += stride) out[i] = in[i];
    } break;
    case PK_OTAPI: {  // in [O][I][T] -> out [O][T][I]
      bf16* out = static_cast<bf16*>(j.out0);
      const int I = j.b, T = j.c;
      for (size_t i = start; i < j.n; i += stride) {
        const int ci = static_cast<int>(i % I);
        const size_t r = i / I;
        const int t = static_cast<int>(r % T);
        const size_t o = r / T;
        out[i] = __float2bfloat16(in[(o * I + ci) * T + t]);
      }
    } break;
    case PK_DW: {  // in [C][T] -> out [T][C] (+ tap-reversed copy)
      bf16* out = static_cast<bf16*>(j.out0);
      bf16* out_flip = static_cast<bf16*>(j.out1);
      const int C = j.a, T = j.b;
      for (size_t i = start; i < j.n; i += stride) {
        const int c = static_cast<int>(i % C), t = static_cast<int>(i / C);
        const bf16 v = __float2bfloat16(in[c * T + t]);
        out[i] = v;
        if (out_flip) out_flip[(T - 1 - t) * C + c] = v;
      }
    } break;
    case PK_STEM: {  // in [16][27] -> out [27][16]
      float* out = static_cast<float*>(j.out0);
      for (size_t i = start; i < j.n; i += stride) out[i] = in[(i % 16) * 27 + i / 16];
    } break;
    case PK_FOLD: {
      const float* b = static_cast<const float*>(j.in1);
      const float* m = static_cast<const float*>(j.in2);
      const float* v = static_cast<const float*>(j.in3);
      float* scale = static_cast<float*>(j.out0);
      float* shift = static_cast<float*>(j.out1);
      for (size_t i = start; i < j.n; i += stride) {
        const float s = in[i] / sqrtf(v[i] + j.eps);
        scale[i] = s;
        shift[i] = b[i] - m[i] * s;
      }
    } break;
    case PK_TRANSPOSE: {  // in [N][K] -> out [K][N]
      bf16* out = static_cast<bf16*>(j.out0);
      const int N = j.a, K = j.b;
      for (size_t i = start; i < j.n; i += stride) {
        const int nn = static_cast<int>(i % N), kk = static_cast<int>(i / N);
        out[i] = __float2bfloat16(in[static_cast<size_t>(nn) * K + kk]);
      }
    } break;
    case PK_DGRAD3: {  // in [O][I][9] -> out [I][9][O], taps flipped
      bf16* out = static_cast<bf16*>(j.out0);
      const int O = j.a, I = j.b;
      for (size_t i = start; i < j.n; i += stride) {
        const int o = static_cast<int>(i % O);
        const size_t r = i / O;
        const int t = static_cast<int>(r % 9);
        const size_t ci = r / 9;
        out[i] = __float2bfloat16(in[(static_cast<size_t>(o) * I + ci) * 9 + (8 - t)]);
      }
    } break;
    case PK_BLOCKDIAG: {  // in [N][K] -> out [pp * N][pp * K], pp copies on the diagonal
      bf16* out = static_cast<bf16*>(j.out0);
      const int N = j.a, K = j.b, pp = j.c;
      const int KK = K * pp;
      for (size_t i = start; i < j.n; i += stride) {
        const int col = static_cast<int>(i % KK), row = static_cast<int>(i / KK);
        const bool on = row / N == col / K;
        out[i] = __float2bfloat16(on ? in[static_cast<size_t>(row % N) * K + col % K] : 0.f);
      }
    } break;
    default: break;
  }
}

// records the job when a batch is open; returns false when the caller has to launch on its own
bool record(int type, int a, int b, int c, float eps, unsigned long long n, const void* in0, const void* in1, const void* in2,
            const void* in3, void* out0, void* out1) {
  if (!g_batch) return false;
  g_batch->push_back(PackJob{type, a, b, c, eps, 0, 0, n, in0, in1, in2, in3, out0, out1});
  return true;
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}
__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = in[i];
}
// in [O][I][T] -> out [O][T][I]
__global__ void oihw_to_otapi_kernel(const float* __restrict__ in, bf16* __restrict__ out, int O, int I, int T) {
  const size_t n = static_cast<size_t>(O) * I * T;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % I);
    const size_t r = i / I;
    const int t = static_cast<int>(r % T);
    const size_t o = r / T;
    out[i] = __float2bfloat16(in[(o * I + ci) * T + t]);
  }
}
// in [C][T] -> out [T][C] and (optional) the tap-reversed copy used by the stride-1 depthwise dgrad
__global__ void pack_dw_kernel(const float* __restrict__ in, bf16* __restrict__ out, bf16* __restrict__ out_flip, int C, int T) {
  const int n = C * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i % C, t = i / C;
    const bf16 v = __float2bfloat16(in[c * T + t]);
    out[i] = v;
    if (out_flip) out_flip[(T - 1 - t) * C + c] = v;
  }
}
// in [16][27] -> out [27][16]
__global__ void pack_stem_kernel(const float* __restrict__ in, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 27 * 16) out[i] = in[(i % 16) * 27 + i / 16];
}
__global__ void fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ m,
                               const float* __restrict__ v, float eps, float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float s = g[i] / sqrtf(v[i] + eps);
    scale[i] = s;
    shift[i] = b[i] - m[i] * s;
  }
}

// in [N][K] -> out [K][N]
__global__ void transpose_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int N, int K) {
  const size_t n = static_cast<size_t>(N) * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int nn = static_cast<int>(i % N), kk = static_cast<int>(i / N);
    out[i] = __float2bfloat16(in[static_cast<size_t>(nn) * K + kk]);
  }
}
// in [O][I][9] -> out [I][9][O] with the taps flipped (dgrad of a stride-1 pad-1 3x3 conv is a 3x3 conv with w^T, rot180)
__global__ void dgrad3x3_kernel(const float* __restrict__ in, bf16* __restrict__ out, int O, int I) {
  const size_t n = static_cast<size_t>(O) * I * 9;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i % O);
    const size_t r = i / O;
    const int t = static_cast<int>(r % 9);
    const size_t ci = r / 9;
    out[i] = __float2bfloat16(in[(static_cast<size_t>(o) * I + ci) * 9 + (8 - t)]);
  }
}

__global__ void blockdiag_kernel(const float* __restrict__ in, bf16* __restrict__ out, int N, int K, int pp) {
  const int KK = K * pp;
  const size_t n = static_cast<size_t>(N) * pp * KK;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % KK), row = static_cast<int>(i / KK);
    const bool on = row / N == col / K;
    out[i] = __float2bfloat16(on ? in[static_cast<size_t>(row % N) * K + col % K] : 0.f);
  }
}

inline int blocks_for(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 1024) b = 1024;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

int launch_cast_bf16(const float* in, bf16* out, size_t n, cudaStream_t st) {
  if (record(PK_CAST, 0, 0, 0, 0.f, n, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  cast_bf16_kernel<<<blocks_for(n), 256, 0, st>>>(in, out, n);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_copy_f32(const float* in, float* out, size_t n, cudaStream_t st) {
  if (record(PK_COPY, 0, 0, 0, 0.f, n, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  copy_f32_kernel<<<blocks_for(n), 256, 0, st>>>(in, out, n);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_oihw_to_otapi(const float* in, bf16* out, int O, int I, int taps, cudaStream_t st) {
  if (record(PK_OTAPI, O, I, taps, 0.f, static_cast<unsigned long long>(O) * I * taps, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  oihw_to_otapi_kernel<<<blocks_for(static_cast<size_t>(O) * I * taps), 256, 0, st>>>(in, out, O, I, taps);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_dw(const float* in, bf16* out, bf16* out_flip, int C, int taps, cudaStream_t st) {
  if (record(PK_DW, C, taps, 0, 0.f, static_cast<unsigned long long>(C) * taps, in, nullptr, nullptr, nullptr, out, out_flip)) return MTG_OK;
  pack_dw_kernel<<<blocks_for(static_cast<size_t>(C) * taps), 256, 0, st>>>(in, out, out_flip, C, taps);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_transpose_bf16(const float* in, bf16* out, int N, int K, cudaStream_t st) {
  if (record(PK_TRANSPOSE, N, K, 0, 0.f, static_cast<unsigned long long>(N) * K, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  transpose_bf16_kernel<<<blocks_for(static_cast<size_t>(N) * K), 256, 0, st>>>(in, out, N, K);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_dgrad3x3(const float* in, bf16* out, int O, int I, cudaStream_t st) {
  if (record(PK_DGRAD3, O, I, 0, 0.f, static_cast<unsigned long long>(O) * I * 9, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  dgrad3x3_kernel<<<blocks_for(static_cast<size_t>(O) * I * 9), 256, 0, st>>>(in, out, O, I);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_blockdiag(const float* in, bf16* out, int N, int K, int pp, cudaStream_t st) {
  const unsigned long long n = static_cast<unsigned long long>(N) * pp * K * pp;
  if (record(PK_BLOCKDIAG, N, K, pp, 0.f, n, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  blockdiag_kernel<<<blocks_for(n), 256, 0, st>>>(in, out, N, K, pp);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_pack_stem(const float* in, float* out, cudaStream_t st) {
  if (record(PK_STEM, 0, 0, 0, 0.f, 27 * 16, in, nullptr, nullptr, nullptr, out, nullptr)) return MTG_OK;
  pack_stem_kernel<<<2, 256, 0, st>>>(in, out);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}
int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                   float* shift, int C, cudaStream_t st) {
  if (record(PK_FOLD, C, 0, 0, eps, static_cast<unsigned long long>(C), gamma, beta, mean, var, scale, shift)) return MTG_OK;
  fold_bn_kernel<<<ceil_div(C, 256), 256, 0, st>>>(gamma, beta, mean, var, eps, scale, shift, C);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

void pack_batch_begin() {
  if (!g_batch) g_batch = new std::vector<PackJob>();
  g_batch->clear();
}

void pack_batch_abort() {
  delete g_batch;
  g_batch = nullptr;
}

int pack_batch_flush(cudaStream_t st) {
  if (!g_batch) return MTG_OK;
  std::vector<PackJob> jobs;
  jobs.swap(*g_batch);
  pack_batch_abort();
  for (size_t first = 0; first < jobs.size(); first += PACK_MAX_JOBS) {
    PackTable t;
    t.njobs = static_cast<int>(jobs.size() - first < static_cast<size_t>(PACK_MAX_JOBS) ? jobs.size() - first : PACK_MAX_JOBS);
    int blocks = 0;
    for (int k = 0; k < t.njobs; ++k) {
      t.jobs[k] = jobs[first + k];
      unsigned long long nb = (t.jobs[k].n + 1023) / 1024;  // >= 4 elements per thread
      if (nb > 128) nb = 128;
      if (nb < 1) nb = 1;
      t.jobs[k].first_block = blocks;
      t.jobs[k].nblocks = static_cast<int>(nb);
      blocks += static_cast<int>(nb);
    }
    pack_batch_kernel<<<blocks, 256, 0, st>>>(t);
    MTG_LAUNCH_CHECK();
  }
  return MTG_OK;
}

}  // namespace mtgseg
