// Corner-keypoint head of the pose pipeline (BASELINE.json configs[4]):
// HRNetPoseHead.forward (train-pose-estimation_custom/model.py:10-77) in eval mode and LiteHRNet.decode_heatmaps
// (model.py:133-164).  The timm backbone is out of scope (not importable offline); the head consumes its feature map.
//
//   ConvTranspose2d(Cin,256,4,s2,p1)+BN+ReLU -> ConvTranspose2d(256,256,4,s2,p1)+BN+ReLU
//   -> 2 x [Conv3x3(256,256,bias)+BN+ReLU] -> Conv1x1(256,K,bias) -> AdaptiveAvgPool2d((out_h,out_w)) ; argmax decode
//
// A stride-2 4x4 transposed convolution is four independent 2x2 convolutions, one per output parity class
// (oy%2, ox%2); each is an implicit GEMM with four shifted TMA boxes (offsets in {-1,0,+1}) whose epilogue stores
// through a strided tensor map straight into every other pixel of the full-resolution NHWC tensor.  All dense layers
// run on the tcgen05 kernel of gemm_tc.cu (multi-tap mode); this file only adds the glue kernels.
#include "net.h"

namespace mtgseg {
namespace {

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; }
};

// [B][C][HW] fp32 -> [B][HW][C] bf16
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, bf16* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, px = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && px < HW) ? in[(static_cast<size_t>(b) * C + c) * HW + px] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int px = p0 + i, c = c0 + threadIdx.x;
    if (c < C && px < HW) out[(static_cast<size_t>(b) * HW + px) * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// ConvTranspose2d weight [Cin][Cout][4][4] fp32 -> four parity packs [parity][Cout][4 taps][Cin] bf16
__global__ void pack_deconv_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout) {
  const size_t per = static_cast<size_t>(Cout) * 4 * Cin, n = 4 * per;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    size_t r = i / Cin;
    const int t = static_cast<int>(r % 4); r /= 4;
    const int co = static_cast<int>(r % Cout);
    const int par = static_cast<int>(r / Cout);
    const int a = par >> 1, b = par & 1, ty = t >> 1, tx = t & 1;
    const int ky = a == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 2 : 0);
    const int kx = b == 0 ? (tx == 0 ? 1 : 3) : (tx == 0 ? 2 : 0);
    out[i] = __float2bfloat16(w[((static_cast<size_t>(ci) * Cout + co) * 4 + ky) * 4 + kx]);
  }
}

// scale = gamma/sqrt(var+eps), shift = beta + (bias - mean)*scale   (conv bias folded; bias may be null)
__global__ void fold_bn_bias_kernel(const float* g, const float* b, const float* m, const float* v, const float* bias, float eps,
                                    float* scale, float* shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float s = g[i] / sqrtf(v[i] + eps);
    scale[i] = s;
    shift[i] = b[i] + ((bias ? bias[i] : 0.f) - m[i]) * s;
  }
}

// final 1x1 weights [NK][256] fp32 -> [8][256] bf16 (zero rows above NK) ; shift8 = bias padded
__global__ void pack_final_kernel(const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ wp, float* __restrict__ shift8,
                                  int NK, int C) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 8 * C; i += gridDim.x * blockDim.x) {
    const int r = i / C;
    wp[i] = __float2bfloat16(r < NK ? w[i] : 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < 8) shift8[threadIdx.x] = threadIdx.x < NK ? bias[threadIdx.x] : 0.f;
}

// nn.AdaptiveAvgPool2d: window [floor(i*in/out), ceil((i+1)*in/out)) ; in [B][Hi][Wi][8] bf16 -> out fp32 [B][NK][Ho][Wo]
__global__ void adaptive_pool_kernel(const bf16* __restrict__ in, float* __restrict__ out, int B, int NK, int Hi, int Wi, int Ho, int Wo) {
  const size_t total = static_cast<size_t>(B) * Ho * Wo;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % Wo), oy = static_cast<int>((i / Wo) % Ho), b = static_cast<int>(i / (static_cast<size_t>(Wo) * Ho));
    const int y0 = (oy * Hi) / Ho, y1 = ((oy + 1) * Hi + Ho - 1) / Ho;
    const int x0 = (ox * Wi) / Wo, x1 = ((ox + 1) * Wi + Wo - 1) / Wo;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        float f[8];
        unpack8(ldg16(in + ((static_cast<size_t>(b) * Hi + y) * Wi + x) * 8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    const float inv = 1.f / static_cast<float>((y1 - y0) * (x1 - x0));
    for (int k = 0; k < NK; ++k) out[((static_cast<size_t>(b) * NK + k) * Ho + oy) * Wo + ox] = acc[k] * inv;
  }
}

// per (image, keypoint): argmax over H*W (lowest index on ties) -> (x/(W-1), y/(H-1)) interleaved
// first maximum of p[0..n) over the block (ties -> lowest index, as torch.max on the CPU); result valid in thread 0
__device__ __forceinline__ int block_argmax(const float* __restrict__ p, int n, float* sv, int* si) {
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = p[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v = sv[threadIdx.x + o]; const int i = si[threadIdx.x + o];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = i; }
    }
    __syncthreads();
  }
  const int r = si[0];
  __syncthreads();  // sv / si may be reused by the caller
  return r;
}

__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ hm, float* __restrict__ coords, int NK, int H, int W) {
  __shared__ float sv[256];
  __shared__ int si[256];
  const int bk = blockIdx.x, n = H * W;
  const int idx = block_argmax(hm + static_cast<size_t>(bk) * n, n, sv, si);
  if (threadIdx.x == 0) {
    const int b = bk / NK, k = bk % NK;
    coords[static_cast<size_t>(b) * NK * 2 + 2 * k] = static_cast<float>(idx % W) / static_cast<float>(W - 1);
    coords[static_cast<size_t>(b) * NK * 2 + 2 * k + 1] = static_cast<float>(idx / W) / static_cast<float>(H - 1);
  }
}

// CornerMetrics.update (train-pose-estimation_custom/metrics.py:29-73): argmax of the predicted and of the target heatmap,
// both scaled to image pixels, Euclidean distance; accumulated as {sum of distances (fp64), n, n(<=3px), n(<=6px)}.
// The fp32 arithmetic follows the reference's operation order (x * image_w / (W-1); sqrt(dx*dx + dy*dy) without FMA
// contraction), so the threshold counts are exactly the reference's.
struct CornerAcc { double sum; unsigned long long n, n3, n6; };
__global__ void __launch_bounds__(256) corner_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, CornerAcc* acc,
                                                             int H, int W, float image_w, float image_h) {
  __shared__ float sv[256];
  __shared__ int si[256];
  const int n = H * W;
  const int ip = block_argmax(pred + static_cast<size_t>(blockIdx.x) * n, n, sv, si);
  const int it = block_argmax(target + static_cast<size_t>(blockIdx.x) * n, n, sv, si);
  if (threadIdx.x == 0) {
    const float wd = static_cast<float>(W - 1), hd = static_cast<float>(H - 1);
    const float px = __fdiv_rn(__fmul_rn(static_cast<float>(ip % W), image_w), wd), py = __fdiv_rn(__fmul_rn(static_cast<float>(ip / W), image_h), hd);
    const float tx = __fdiv_rn(__fmul_rn(static_cast<float>(it % W), image_w), wd), ty = __fdiv_rn(__fmul_rn(static_cast<float>(it / W), image_h), hd);
    const float dx = __fsub_rn(px, tx), dy = __fsub_rn(py, ty);
    const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    atomicAdd(&acc->sum, static_cast<double>(dist));
    atomicAdd(&acc->n, 1ull);
    if (dist <= 3.0f) atomicAdd(&acc->n3, 1ull);
    if (dist <= 6.0f) atomicAdd(&acc->n6, 1ull);
  }
}

// CornerLoss = nn.MSELoss() on the heatmaps (metrics.py:105-136): mean((p - t)^2), optional dpred = 2 (p - t) / n.
// Two fixed-order stages (per-block partial sums, then one block): run-to-run deterministic.
constexpr int MSE_BLOCKS = 1024;
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ p, const float* __restrict__ t, float* __restrict__ dpred,
                                                          float* __restrict__ partial, long long n, float gscale) {
  __shared__ float red[256];
  float s = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float d = p[i] - t[i];
    s = fmaf(d, d, s);
    if (dpred) dpred[i] = d * gscale;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) mse_final_kernel(const float* __restrict__ partial, int blocks, double inv_n, float* __restrict__ loss) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 256) s += static_cast<double>(partial[i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = static_cast<float>(red[0] * inv_n);
}

struct PosePlan {
  mtgseg_pose_desc d;
  // packed arena
  size_t dc_w[2], dc_scale[2], dc_shift[2];   // deconv parity packs + folded BN
  size_t cv_w[2], cv_scale[2], cv_shift[2];   // 3x3 convs
  size_t fin_w, fin_shift;
  size_t packed_bytes;
};

int plan_pose(const mtgseg_pose_desc& d, PosePlan& P) {
  MTG_REQUIRE(d.in_channels % 8 == 0 && d.in_channels >= 8, MTG_ERR_UNSUPPORTED, "pose: in_channels must be a multiple of 8");
  MTG_REQUIRE(d.num_keypoints >= 1 && d.num_keypoints <= 8, MTG_ERR_UNSUPPORTED, "pose: num_keypoints not in [1,8]");
  MTG_REQUIRE(d.feat_w * 4 <= 128 && d.feat_h > 0 && d.feat_w > 0, MTG_ERR_UNSUPPORTED, "pose: feature map wider than 32 is not supported");
  P.d = d;
  Bump a;
  int cin = d.in_channels;
  for (int i = 0; i < 2; ++i) {
    P.dc_w[i] = a.take(static_cast<size_t>(4) * 256 * 4 * cin * 2);
    P.dc_scale[i] = a.take(256 * 4); P.dc_shift[i] = a.take(256 * 4);
    cin = 256;
  }
  for (int i = 0; i < 2; ++i) {
    P.cv_w[i] = a.take(static_cast<size_t>(256) * 9 * 256 * 2);
    P.cv_scale[i] = a.take(256 * 4); P.cv_shift[i] = a.take(256 * 4);
  }
  P.fin_w = a.take(8 * 256 * 2);
  P.fin_shift = a.take(8 * 4);
  P.packed_bytes = a.off;
  return MTG_OK;
}

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

}  // namespace

}  // namespace mtgseg

using namespace mtgseg;

extern "C" {

int mtgseg_pose_param_count(void) { return 28; }

size_t mtgseg_pose_packed_bytes(const mtgseg_pose_desc* d) {
  PosePlan P;
  return (d && plan_pose(*d, P) == MTG_OK) ? P.packed_bytes : 0;
}

size_t mtgseg_pose_workspace_bytes(const mtgseg_pose_desc* d, int batch) {
  PosePlan P;
  if (!d || plan_pose(*d, P) != MTG_OK || batch <= 0) return 0;
  const size_t B = batch, hw = static_cast<size_t>(d->feat_h) * d->feat_w;
  Bump b;
  b.take(B * hw * d->in_channels * 2);
  b.take(B * hw * 4 * 256 * 2);
  b.take(B * hw * 16 * 256 * 2);
  b.take(B * hw * 16 * 256 * 2);
  b.take(B * hw * 16 * 8 * 2);
  return b.off;
}

// params: the 28 state_dict entries of HRNetPoseHead in order (device pointers)
int mtgseg_pose_pack_weights(const mtgseg_pose_desc* d, const void* const* params, int n_params, void* packed, void* stream) {
  PosePlan P;
  MTG_REQUIRE(d && params && packed, MTG_ERR_ARG, "pose_pack: null pointer");
  RC(plan_pose(*d, P));
  MTG_REQUIRE(n_params == 28, MTG_ERR_ARG, "pose_pack: expected 28 state_dict entries, got %d", n_params);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(packed);
  auto f = [&](int i) { return static_cast<const float*>(params[i]); };
  int cin = d->in_channels;
  for (int i = 0; i < 2; ++i) {  // entries: w, bn.weight, bn.bias, bn.mean, bn.var, bn.nbt
    const int o = i * 6;
    const size_t n = static_cast<size_t>(4) * 256 * 4 * cin;
    pack_deconv_kernel<<<static_cast<unsigned>((n + 255) / 256 > 2048 ? 2048 : (n + 255) / 256), 256, 0, st>>>(f(o), reinterpret_cast<bf16*>(base + P.dc_w[i]), cin, 256);
    MTG_LAUNCH_CHECK();
    fold_bn_bias_kernel<<<1, 256, 0, st>>>(f(o + 1), f(o + 2), f(o + 3), f(o + 4), nullptr, 1e-5f, reinterpret_cast<float*>(base + P.dc_scale[i]),
                                           reinterpret_cast<float*>(base + P.dc_shift[i]), 256);
    MTG_LAUNCH_CHECK();
    cin = 256;
  }
  for (int i = 0; i < 2; ++i) {  // entries: conv w, conv b, bn.weight, bn.bias, bn.mean, bn.var, bn.nbt
    const int o = 12 + i * 7;
    RC(launch_pack_oihw_to_otapi(f(o), reinterpret_cast<bf16*>(base + P.cv_w[i]), 256, 256, 9, st));
    fold_bn_bias_kernel<<<1, 256, 0, st>>>(f(o + 2), f(o + 3), f(o + 4), f(o + 5), f(o + 1), 1e-5f, reinterpret_cast<float*>(base + P.cv_scale[i]),
                                           reinterpret_cast<float*>(base + P.cv_shift[i]), 256);
    MTG_LAUNCH_CHECK();
  }
  pack_final_kernel<<<8, 256, 0, st>>>(f(26), f(27), reinterpret_cast<bf16*>(base + P.fin_w), reinterpret_cast<float*>(base + P.fin_shift),
                                       d->num_keypoints, 256);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int mtgseg_decode_heatmaps(const float* heatmaps, float* coords, int batch, int num_keypoints, int H, int W, void* stream) {
  MTG_REQUIRE(heatmaps && coords && batch > 0 && num_keypoints > 0 && H > 1 && W > 1, MTG_ERR_ARG, "decode_heatmaps: bad arguments");
  decode_kernel<<<batch * num_keypoints, 256, 0, static_cast<cudaStream_t>(stream)>>>(heatmaps, coords, num_keypoints, H, W);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int mtgseg_corner_metrics(const float* pred, const float* target, void* acc, int batch, int num_keypoints, int H, int W, float image_w,
                          float image_h, void* stream) {
  MTG_REQUIRE(pred && target && acc && batch > 0 && num_keypoints > 0 && H > 1 && W > 1, MTG_ERR_ARG, "corner_metrics: bad arguments");
  static_assert(sizeof(CornerAcc) == 32, "accumulator layout is part of the ABI: {double sum; uint64 n, n3, n6}");
  corner_metrics_kernel<<<batch * num_keypoints, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, static_cast<CornerAcc*>(acc), H, W,
                                                                                             image_w, image_h);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

size_t mtgseg_mse_scratch_floats(void) { return MSE_BLOCKS; }

int mtgseg_mse_loss(const float* pred, const float* target, float* dpred, float* loss, float* scratch, long long n, void* stream) {
  MTG_REQUIRE(pred && target && loss && scratch && n > 0, MTG_ERR_ARG, "mse_loss: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > MSE_BLOCKS) blocks = MSE_BLOCKS;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mse_partial_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(pred, target, dpred, scratch, n, 2.0f / static_cast<float>(n));
  MTG_LAUNCH_CHECK();
  mse_final_kernel<<<1, 256, 0, st>>>(scratch, static_cast<int>(blocks), 1.0 / static_cast<double>(n), loss);
  MTG_LAUNCH_CHECK();
  return MTG_OK;
}

int mtgseg_pose_forward(const mtgseg_pose_desc* d, const float* features, const void* packed, float* heatmaps, float* coords,
                        void* workspace, size_t workspace_bytes, int batch, void* stream) {
  PosePlan P;
  MTG_REQUIRE(d && features && packed && heatmaps && workspace, MTG_ERR_ARG, "pose_forward: null pointer");
  RC(plan_pose(*d, P));
  MTG_REQUIRE(batch > 0, MTG_ERR_ARG, "pose_forward: batch must be positive");
  MTG_REQUIRE(workspace_bytes >= mtgseg_pose_workspace_bytes(d, batch), MTG_ERR_WORKSPACE, "pose_forward: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int B = batch, Hf = d->feat_h, Wf = d->feat_w, Cin = d->in_channels;
  const size_t hw = static_cast<size_t>(Hf) * Wf;
  Bump b;
  bf16* x0 = reinterpret_cast<bf16*>(ws + b.take(B * hw * Cin * 2));
  bf16* x1 = reinterpret_cast<bf16*>(ws + b.take(B * hw * 4 * 256 * 2));
  bf16* x2 = reinterpret_cast<bf16*>(ws + b.take(B * hw * 16 * 256 * 2));
  bf16* x3 = reinterpret_cast<bf16*>(ws + b.take(B * hw * 16 * 256 * 2));
  bf16* x4 = reinterpret_cast<bf16*>(ws + b.take(B * hw * 16 * 8 * 2));
  {
    dim3 grid(ceil_div(static_cast<int>(hw), 32), ceil_div(Cin, 32), B), blk(32, 8);
    nchw_to_nhwc_kernel<<<grid, blk, 0, st>>>(features, x0, Cin, static_cast<int>(hw));
    MTG_LAUNCH_CHECK();
  }
  // two transposed convolutions: 4 parity GEMMs each
  const bf16* in = x0;
  bf16* outs[2] = {x1, x2};
  int H = Hf, W = Wf, cin = Cin;
  for (int i = 0; i < 2; ++i) {
    for (int par = 0; par < 4; ++par) {
      const int a = par >> 1, bb = par & 1;
      ConvGemmArgs g;
      g.a = in; g.w = reinterpret_cast<const bf16*>(pk + P.dc_w[i]) + static_cast<size_t>(par) * 256 * 4 * cin;
      g.out = outs[i] + (static_cast<size_t>(a) * 2 * W + bb) * 256;
      g.M = B * H * W; g.N = 256; g.K = cin;
      g.scale = reinterpret_cast<const float*>(pk + P.dc_scale[i]); g.shift = reinterpret_cast<const float*>(pk + P.dc_shift[i]);
      g.act = ACT_RELU; g.conv3x3 = 1; g.B = B; g.H = H; g.W = W; g.ntaps = 4;
      for (int t = 0; t < 4; ++t) {
        const int ty = t >> 1, tx = t & 1;
        g.tap_dy[t] = a == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 0 : 1);
        g.tap_dx[t] = bb == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 0 : 1);
      }
      g.out_sx = 2 * 256; g.out_sy = static_cast<long long>(2) * (2 * W) * 256; g.out_sn = static_cast<long long>(2 * H) * (2 * W) * 256;
      RC(launch_conv_gemm(g, st));
    }
    in = outs[i]; H *= 2; W *= 2; cin = 256;
  }
  // two 3x3 convolutions (bias folded into the BN shift) + ReLU
  bf16* pp[3] = {x2, x3, x2};
  for (int i = 0; i < 2; ++i) {
    ConvGemmArgs g;
    g.a = pp[i]; g.w = reinterpret_cast<const bf16*>(pk + P.cv_w[i]); g.out = pp[i + 1];
    g.M = B * H * W; g.N = 256; g.K = 256;
    g.scale = reinterpret_cast<const float*>(pk + P.cv_scale[i]); g.shift = reinterpret_cast<const float*>(pk + P.cv_shift[i]);
    g.act = ACT_RELU; g.conv3x3 = 1; g.B = B; g.H = H; g.W = W;
    RC(launch_conv_gemm(g, st));
  }
  {  // final 1x1 (+bias), channels padded to 8
    ConvGemmArgs g;
    g.a = x2; g.w = reinterpret_cast<const bf16*>(pk + P.fin_w); g.out = x4; g.M = B * H * W; g.N = 8; g.K = 256;
    g.shift = reinterpret_cast<const float*>(pk + P.fin_shift); g.act = ACT_NONE;
    RC(launch_conv_gemm(g, st));
  }
  {
    const size_t total = static_cast<size_t>(B) * d->out_h * d->out_w;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adaptive_pool_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x4, heatmaps, B, d->num_keypoints, H, W, d->out_h, d->out_w);
    MTG_LAUNCH_CHECK();
  }
  if (coords) return mtgseg_decode_heatmaps(heatmaps, coords, B, d->num_keypoints, d->out_h, d->out_w, stream);
  return MTG_OK;
}

}  // extern "C"
