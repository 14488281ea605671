// Internal launcher interface between the network plan (net.cu), the C-ABI (api.cu) and the kernels.
// All activations are NHWC bf16 ("pixels x channels" matrices); all launchers are asynchronous on `st`.
#pragma once
#include "common.cuh"

namespace mtgseg {

// ---- tcgen05 implicit-GEMM convolution (gemm_tc.cu) ---------------------------------------------
struct ConvGemmArgs {
  const bf16* a = nullptr;     // 1x1: [M][K]; 3x3: NHWC [B][H][W][K]
  const bf16* w = nullptr;     // 1x1: [N][K]; 3x3: [N][9][K]   (K contiguous)
  bf16* out = nullptr;         // [M][N]
  int M = 0, N = 0, K = 0;
  const float* scale = nullptr;  // [N] folded BatchNorm scale (nullptr -> 1)
  const float* shift = nullptr;  // [N] folded BatchNorm shift (nullptr -> 0)
  int act = ACT_NONE;
  const bf16* residual = nullptr;  // [M][N] added after the activation-less projection
  const float* a_scale = nullptr;  // squeeze-excite: [B][K] multiplier applied to A rows of image b
  int hw = 0;                      // rows per image (needed with a_scale)
  int conv3x3 = 0;                 // multi-tap mode: 3x3, stride 1, pad 1 (default taps) or an explicit tap list
  int B = 0, H = 0, W = 0;         // geometry for the multi-tap mode (M == B*H*W)
  int ntaps = 0;                   // 0 -> the nine 3x3 taps; else 1..9 explicit (dy, dx) offsets in [-1, 1], weights [N][ntaps][K]
  int tap_dy[9] = {0}, tap_dx[9] = {0};
  long long out_sx = 0, out_sy = 0, out_sn = 0;  // output strides in elements (0 -> dense NHWC); lets a parity class of a
                                                 // stride-2 transposed conv write every other pixel of the full-size tensor
  double* stat = nullptr;  // training: [2][N] fp64, += per-channel sum / sum of squares of the stored outputs (zero it first)
};
int launch_conv_gemm(const ConvGemmArgs& g, cudaStream_t st);
// haloed-tile 3x3 (conv3_tc.cu): MTG_OK / error, or 1 when the shape is left to launch_conv_gemm's nine-shifted-boxes path
int launch_conv3x3_halo(const ConvGemmArgs& g, cudaStream_t st);

// ---- bandwidth-bound kernels (dwconv.cu, stem.cu, se.cu, tail.cu, metrics.cu) -------------------
struct DwConvArgs {
  const bf16* in = nullptr;   // [B][H][W][C]
  const bf16* w = nullptr;    // [k*k][C] bf16, tap-major
  bf16* out = nullptr;        // [B][Ho][Wo][C]
  const float* scale = nullptr;
  const float* shift = nullptr;
  int act = ACT_NONE;
  int B = 0, H = 0, W = 0, C = 0, k = 3, stride = 1, dil = 1;
  float* gap_partial = nullptr;  // optional [B][chunks][C] per-chunk channel sums of the output
  int chunks = 1;                // pixel chunks per image (grid.x); see dwconv_chunks()
  double* stat = nullptr;        // training: [2][C] fp64, += per-channel sum / sum of squares of the stored outputs (zero it first)
};
int dwconv_chunks(int H, int W, int C, int k, int stride, int dil, bool need_gap);
int launch_dwconv(const DwConvArgs& a, cudaStream_t st);

struct StemArgs {
  const float* x = nullptr;  // [B][3][H][W] fp32 NCHW (the reference's data contract), or
  const uint8_t* x_u8 = nullptr;  // [B][H][W][3] raw uint8 pixels, normalised on load with mean/std below
  float mean[3] = {0.485f, 0.456f, 0.406f}, std[3] = {0.229f, 0.224f, 0.225f};  // train/dataset.py:182-185
  const float* w = nullptr;  // [27][16] fp32, index (ci*9 + ky*3 + kx)
  const float* scale = nullptr;
  const float* shift = nullptr;
  bf16* out = nullptr;  // [B][Ho][Wo][16]
  int B = 0, H = 0, W = 0;
  int act = ACT_HSWISH;  // ACT_NONE in training mode (BatchNorm uses batch statistics afterwards)
};
int launch_stem(const StemArgs& a, cudaStream_t st);

// channel sums over pixels: in [B][HW][C] bf16 -> out [B][C] fp32 (sums, not means)
int launch_gap(const bf16* in, float* out, int B, int HW, int C, cudaStream_t st);

// Two-layer (or one-layer when w2 == nullptr) per-image MLP on pooled features:
//   mean = sum(partials)/HW ; h = act1(W1 mean + b1) ; out = act2(W2 h + b2)   (out = h if no W2)
struct SeMlpArgs {
  const float* sums = nullptr;  // [B][chunks][C]
  int chunks = 1;
  int B = 0, C = 0, SQ = 0, HW = 0;
  const bf16* w1 = nullptr;   // [SQ][C]
  const float* b1 = nullptr;  // [SQ] or nullptr
  int act1 = ACT_RELU;
  const bf16* w2 = nullptr;   // [C][SQ] or nullptr
  const float* b2 = nullptr;
  int act2 = ACT_HSIGMOID;
  float* out = nullptr;     // [B][C] (or [B][SQ] when single layer)
  float* hidden = nullptr;  // [B][SQ] scratch, required for the two-layer form
};
int launch_se_mlp(const SeMlpArgs& a, cudaStream_t st);

struct HeadMixArgs {
  const bf16* cbr = nullptr;    // [B][Hh][Wh][IC]  relu(bn(conv3x3(high)))
  const float* s = nullptr;     // [B][IC]          sigmoid branch
  const bf16* low = nullptr;    // [B][Hl][Wl][LC]
  const float* w_high = nullptr;  // [NC][IC]
  const float* b_high = nullptr;  // [NC]
  const float* w_low = nullptr;   // [NC][LC]
  const float* b_low = nullptr;   // [NC]
  float* out = nullptr;           // [B][Hl][Wl][NC] fp32 low-resolution logits
  int B = 0, Hh = 0, Wh = 0, Hl = 0, Wl = 0, IC = 0, LC = 0, NC = 0;
};
int launch_head_mix(const HeadMixArgs& a, cudaStream_t st);

enum LogitsDtype : int { LOGITS_NONE = 0, LOGITS_F32 = 1, LOGITS_BF16 = 2, LOGITS_F16 = 3 };
struct UpsampleOutArgs {
  const float* lowres = nullptr;  // [B][Hl][Wl][NC]
  void* logits = nullptr;         // [B][NC][H][W] (dtype below) or nullptr
  int logits_dtype = LOGITS_F32;
  uint8_t* mask = nullptr;            // [B][H][W] argmax (ties -> lowest class) or nullptr
  const int64_t* targets = nullptr;   // [B][H][W] for counts
  unsigned long long* counts = nullptr;  // [4] n00,n01,n10,n11 (accumulated with atomics), NC == 2 only
  int B = 0, Hl = 0, Wl = 0, H = 0, W = 0, NC = 0;
};
int launch_upsample_out(const UpsampleOutArgs& a, cudaStream_t st);

// 2x2 confusion counts from full-resolution logits [B][2][H][W] and int64 targets.
int launch_metric_counts(const void* logits, int logits_dtype, const int64_t* targets, unsigned long long* counts4,
                         long long batch, long long hw, cudaStream_t st);

// ---- fused CombinedLoss forward + gradient (loss.cu) ---------------------------------------------
size_t loss_scratch_bytes();
// loss3 = {total, dice_loss, ce_loss}; dlogits (logits' layout; logits' dtype or float32; nullable) = dLoss/dlogits
int launch_loss(const void* logits, int dtype, const int64_t* targets, void* dlogits, int dlogits_dtype, float* scratch, float* loss3,
                long long batch, long long hw, int nc, float dice_w, float ce_w, float smooth, cudaStream_t st);

// the same loss from the low-resolution logits [B][Hl][Wl][NC] of the head (the x8 bilinear upsample recomputed on the fly):
// loss3 and d_lowres = dLoss/d(lowres); scratch >= lowres_loss_scratch_floats() floats.  Deterministic.
size_t lowres_loss_scratch_floats(int B, int Hl, int Wl);
int launch_lowres_loss(const float* lowres, const int64_t* targets, float* d_lowres, float* scratch, float* loss3, int B, int Hl, int Wl,
                       int H, int W, int nc, float dice_w, float ce_w, float smooth, cudaStream_t st);

// ---- training-mode BatchNorm forward / backward (bn_train.cu) --------------------------------------
int bn_chunks(int HW, int C);
size_t bn_partial_floats(int B, int HW, int C);
struct BnTrainFwdArgs {
  const bf16* z = nullptr;      // [B][HW][C] raw conv output
  bf16* y = nullptr;            // act(bn(z)) (+ residual)
  const bf16* residual = nullptr;
  const float* gamma = nullptr; const float* beta = nullptr;
  float eps = 1e-3f, momentum = 1e-2f;
  float* running_mean = nullptr; float* running_var = nullptr; long long* num_batches_tracked = nullptr;  // updated in place
  float* scale = nullptr; float* shift = nullptr; float* save_mean = nullptr; float* save_rstd = nullptr;  // [C] each (saved for bwd)
  double* stat = nullptr;       // [2][C] fp64 accumulators: sum z, sum z^2
  bool stats_done = false;      // true: the producing conv kernel already accumulated `stat` from its epilogue (zeroed before it ran)
  float* gap = nullptr; int gap_chunks = 1;  // optional [B][gap_chunks][C] channel sums of y (squeeze-excite pool)
  int act = ACT_NONE, B = 0, HW = 0, C = 0;
};
int launch_bn_train_fwd(const BnTrainFwdArgs& a, cudaStream_t st);
struct BnTrainBwdArgs {
  const bf16* z = nullptr; const bf16* dy = nullptr; bf16* dz = nullptr;
  const float* scale = nullptr; const float* shift = nullptr; const float* save_mean = nullptr; const float* save_rstd = nullptr;
  const float* se_s = nullptr; const float* se_dmean = nullptr;  // optional [B][C]: dy' = dy*se_s + se_dmean/HW
  double* bstat = nullptr;      // [2][C] fp64 accumulators: sum dyh, sum dyh*xhat
  bool bstat_zeroed = false;    // true: the caller cleared `bstat` (one memset for all layers of a step)
  float* dgamma = nullptr; float* dbeta = nullptr;
  int act = ACT_NONE, B = 0, HW = 0, C = 0;
};
int launch_bn_train_bwd(const BnTrainBwdArgs& a, cudaStream_t st);

// ---- convolution backward kernels (train_conv.cu) ---------------------------------------------------
struct WgradArgs {
  const bf16* dz = nullptr;   // [M][N] gradient w.r.t. the raw conv output
  const bf16* x = nullptr;    // [M][K] conv input (NHWC)
  float* dw = nullptr;        // fp32 [N][K][taps] (OIHW), ACCUMULATED with atomics: zero it first
  const float* a_scale = nullptr; int hw = 0;  // squeeze-excite multiplier [B][K] the forward applied to x
  long long M = 0; int N = 0, K = 0, taps = 1, H = 0, W = 0;
};
int launch_wgrad(const WgradArgs& a, cudaStream_t st);         // CUDA-core fallback (train_conv.cu)
int launch_wgrad_tc(const WgradArgs& a, int B, cudaStream_t st);  // tcgen05, MN-major operands (wgrad_tc.cu); needs a.hw = pixels per image
struct DwBwdArgs {
  const bf16* dz = nullptr; const bf16* x = nullptr; const bf16* w = nullptr;  // w: bf16 [k*k][C]
  bf16* dx = nullptr; float* dw = nullptr;                                      // dw: fp32 [C][k*k], accumulated
  int B = 0, H = 0, W = 0, C = 0, k = 3, stride = 1, dil = 1;
};
int launch_dw_dgrad(const DwBwdArgs& a, cudaStream_t st);
int launch_dw_wgrad(const DwBwdArgs& a, cudaStream_t st);
int launch_stem_wgrad(const float* x, const bf16* dz, float* dw, int B, int H, int W, cudaStream_t st);

// ---- small backward kernels + AdamW (train_misc.cu) --------------------------------------------------
int launch_dot_pool(const bf16* a, const bf16* b, float* out, int B, int HW, int C, int chunks, cudaStream_t st);
struct SeBwdArgs {
  const float* ds_partial = nullptr; int chunks = 1;
  const float* s = nullptr; const float* hid = nullptr;
  const float* w1 = nullptr; const float* w2 = nullptr;  // fp32 master weights; w2 == nullptr: single sigmoid layer
  float* dpre2 = nullptr; float* dpre1 = nullptr; float* dmean = nullptr;
  int B = 0, C = 0, SQ = 0;
};
int launch_se_bwd(const SeBwdArgs& a, cudaStream_t st);
int launch_outer_sum(const float* u, const float* v, int v_chunks, float vscale, float* dw, float* dbias, int B, int I, int J,
                     cudaStream_t st);
int launch_upsample_bwd(const void* g, int dtype, float* out, int B, int NC, int Hc, int Wc, int Hf, int Wf, long long sn, long long sc,
                        long long sp, cudaStream_t st);
struct HeadBwdArgs {
  const float* d_o = nullptr; const float* dh2 = nullptr; const bf16* cbr = nullptr; const float* s = nullptr; const bf16* low = nullptr;
  const float* w_high = nullptr; const float* w_low = nullptr;
  bf16* dcbr = nullptr; float* ds = nullptr; bf16* dlow = nullptr;
  float* dw_high = nullptr; float* dw_low = nullptr; float* db_high = nullptr; float* db_low = nullptr;  // accumulated (atomics)
  int B = 0, Hh = 0, Wh = 0, Hl = 0, Wl = 0, IC = 0, LC = 0, NC = 0;
};
int head_bwd_segments(int B);  // pixel segments per image = ds partial slots per image
int launch_head_bwd(const HeadBwdArgs& a, cudaStream_t st);  // a.ds: [B][head_bwd_segments(B)][IC] partial sums, overwritten
int launch_add_bf16(const bf16* a, const bf16* b, bf16* out, size_t n, cudaStream_t st);
int launch_fill_f32(float* p, float v, size_t n, cudaStream_t st);
// chunk_table: device array of {float* p; const float* g; float* m; float* v; int n;} (one CTA per entry)
int launch_adamw(const void* chunk_table, int n_chunks, float lr, float b1, float b2, float eps, float wd, int step,
                 const float* inv_scale, const float* found_inf, cudaStream_t st);
// captured (CUDA-graph) steps: the scalars live in an 8-float device block written by launch_adamw_hyper before every replay
int launch_adamw_hyper(float* hyper, float lr, float b1, float b2, float eps, float wd, int step, cudaStream_t st);
int launch_adamw_dev(const void* chunk_table, int n_chunks, const float* hyper, cudaStream_t st);
// [N][K] fp32 -> [K][N] bf16 ; [O][I][3][3] fp32 -> [I][9 (flipped)][O] bf16   (dgrad operands)
int launch_pack_transpose_bf16(const float* in, bf16* out, int N, int K, cudaStream_t st);
int launch_pack_dgrad3x3(const float* in, bf16* out, int O, int I, cudaStream_t st);
// in [N][K] fp32 -> out [pp * N][pp * K] bf16, block diagonal (pp copies of the matrix, zeros elsewhere)
int launch_pack_blockdiag(const float* in, bf16* out, int N, int K, int pp, cudaStream_t st);

// ---- weight packing (pack.cu) --------------------------------------------------------------------
int launch_cast_bf16(const float* in, bf16* out, size_t n, cudaStream_t st);
int launch_copy_f32(const float* in, float* out, size_t n, cudaStream_t st);
// [O][I][kh][kw] fp32 -> [O][kh*kw][I] bf16
int launch_pack_oihw_to_otapi(const float* in, bf16* out, int O, int I, int taps, cudaStream_t st);
// depthwise [C][1][k][k] fp32 -> [k*k][C] bf16
int launch_pack_dw(const float* in, bf16* out, bf16* out_flip, int C, int taps, cudaStream_t st);
// stem [16][3][3][3] fp32 -> [27][16] fp32
int launch_pack_stem(const float* in, float* out, cudaStream_t st);
// between begin and flush the pack launchers only record their job; flush runs all of them in one kernel (pack.cu)
void pack_batch_begin();
int pack_batch_flush(cudaStream_t st);
void pack_batch_abort();
// scale = gamma / sqrt(var + eps), shift = beta - mean * scale
int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                   float* shift, int C, cudaStream_t st);

}  // namespace mtgseg
