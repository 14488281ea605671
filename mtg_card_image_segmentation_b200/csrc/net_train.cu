// Training step of the segmentation network: forward with batch-statistics BatchNorm (train/train.py:82,96-98 ->
// model.train(); model(images)) and backward (loss.backward(), train/train.py:101-105) on the same workspace.
//
// One layout function assigns every saved tensor (raw conv outputs z, activations y, BatchNorm statistics,
// squeeze-excite vectors) and the gradient scratch buffers inside the caller's workspace; forward and backward
// both derive their pointers from it, so nothing is allocated and the two calls always agree.
//
// Forward per conv+BN(+act):  conv kernel (identity epilogue) -> z ; stats -> finalize (scale/shift, running
// stats EMA) -> apply (y = act(z*scale+shift) [+ residual] [+ SE pool partials]).
// Backward per conv+BN(+act): BN/act backward (two-phase reduce + apply, with the SE scale/pool terms folded
// in) -> dz ; wgrad (fp32, reference OIHW layout) ; dgrad (tcgen05 GEMM with transposed weights, residual
// gradient fused as the epilogue's "residual") or the depthwise dgrad kernel.
#include <stdlib.h>

#include <mutex>

#include "net.h"

namespace mtgseg {

namespace {

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

struct LayerBufs {
  bf16* z = nullptr; bf16* y = nullptr;
  float* scale = nullptr; float* shift = nullptr; float* mean = nullptr; float* rstd = nullptr;
  double* stat = nullptr;   // [2][C] sum z, sum z^2 (filled by the producing kernel's epilogue)
  double* bstat = nullptr;  // [2][C] sum dyh, sum dyh*xhat (backward)
  int C = 0, H = 0, W = 0;  // output geometry
};
struct BlockBufs {
  LayerBufs expand, dw, project;
  float* gap = nullptr; int gap_chunks = 1; float* hid = nullptr; float* s = nullptr;
  float* dpre2 = nullptr; float* dpre1 = nullptr;  // backward: pre-activation gradients of the SE MLP, read by the side stream
  int Hin = 0, Win = 0;
};
struct TrainBufs {
  LayerBufs stem, last, cbr;
  BlockBufs blk[kNumBlocks];
  float* hsum = nullptr; float* hscale = nullptr; float* lowres = nullptr;
  uint8_t* stat_arena = nullptr; size_t stat_bytes = 0;    // all layers' forward accumulators: one memset per forward
  uint8_t* bstat_arena = nullptr; size_t bstat_bytes = 0;  // all layers' backward accumulators: one memset per backward
  float* ones = nullptr; float* zeros = nullptr;  // [max(B*960, 960)] constants
  // backward scratch
  bf16* g[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // activation-sized gradient buffers
  bf16* dzpool[4] = {nullptr, nullptr, nullptr, nullptr};  // dz buffers: rotated so that the side-stream wgrad of one layer may still read its dz while the main stream produces the next
  float* pooled[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [B][16][960] / [B][960] SE backward vectors
  float* d_o = nullptr; float* dh2 = nullptr;
  int Hl = 0, Wl = 0, Hh = 0, Wh = 0;
  size_t bytes = 0;
};

inline int conv_out(int in, int k, int stride, int dil) {
  const int pad = (k - 1) / 2 * dil;
  return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1;
}

void layout(const NetPlan& P, int B, uint8_t* ws, TrainBufs& T) {
  Bump b;
  auto bf = [&](size_t elems) { return reinterpret_cast<bf16*>(ws + b.take(elems * sizeof(bf16))); };
  auto f32 = [&](size_t elems) { return reinterpret_cast<float*>(ws + b.take(elems * sizeof(float))); };
  // BatchNorm accumulators of all 47 layers, contiguous (<= 2 x 16 bytes x 960 channels per layer)
  constexpr size_t kStatBytes = 47 * 2 * 960 * sizeof(double);
  T.stat_arena = ws + b.take(kStatBytes); T.stat_bytes = kStatBytes;
  T.bstat_arena = ws + b.take(kStatBytes); T.bstat_bytes = kStatBytes;
  size_t stat_off = 0;
  auto layer = [&](LayerBufs& L, int C, int H, int W) {
    L.C = C; L.H = H; L.W = W;
    const size_t n = static_cast<size_t>(B) * H * W * C;
    L.z = bf(n); L.y = bf(n);
    L.scale = f32(C); L.shift = f32(C); L.mean = f32(C); L.rstd = f32(C);
    L.stat = reinterpret_cast<double*>(T.stat_arena + stat_off);
    L.bstat = reinterpret_cast<double*>(T.bstat_arena + stat_off);
    stat_off += 2 * static_cast<size_t>(C) * sizeof(double);
  };
  int H = conv_out(P.desc.in_h, 3, 2, 1), W = conv_out(P.desc.in_w, 3, 2, 1);
  size_t max_act = static_cast<size_t>(B) * H * W * 16;
  layer(T.stem, 16, H, W);
  for (int i = 0; i < kNumBlocks; ++i) {
    const BlockCfg& c = P.blocks[i].cfg;
    BlockBufs& K = T.blk[i];
    K.Hin = H; K.Win = W;
    if (P.blocks[i].has_expand) {
      layer(K.expand, c.cexp, H, W);
    }
    max_act = max_act > static_cast<size_t>(B) * H * W * c.cexp ? max_act : static_cast<size_t>(B) * H * W * c.cexp;
    const int stride = c.dil > 1 ? 1 : c.stride;
    const int Ho = conv_out(H, c.k, stride, c.dil), Wo = conv_out(W, c.k, stride, c.dil);
    layer(K.dw, c.cexp, Ho, Wo);
    if (c.se) {
      K.gap_chunks = bn_chunks(Ho * Wo, c.cexp);
      if (K.gap_chunks > 16) K.gap_chunks = 16;
      K.gap = f32(static_cast<size_t>(B) * K.gap_chunks * c.cexp);
      K.hid = f32(static_cast<size_t>(B) * P.blocks[i].sq);
      K.s = f32(static_cast<size_t>(B) * c.cexp);
      K.dpre2 = f32(static_cast<size_t>(B) * c.cexp);
      K.dpre1 = f32(static_cast<size_t>(B) * P.blocks[i].sq);
    }
    H = Ho; W = Wo;
    layer(K.project, c.cout, H, W);
    if (i == 3) { T.Hl = H; T.Wl = W; }
  }
  T.Hh = H; T.Wh = W;
  const int ic = P.desc.inter_channels, nc = P.desc.num_classes;
  layer(T.last, 960, H, W);
  layer(T.cbr, ic, H, W);
  T.hsum = f32(static_cast<size_t>(B) * 960);
  T.hscale = f32(static_cast<size_t>(B) * ic);
  T.lowres = f32(static_cast<size_t>(B) * T.Hl * T.Wl * nc);
  T.ones = f32(static_cast<size_t>(B) * 960);
  T.zeros = f32(static_cast<size_t>(B) * 960);
  for (int k = 0; k < 5; ++k) T.g[k] = bf(max_act);
  T.dzpool[0] = T.g[1];
  for (int k = 1; k < 4; ++k) T.dzpool[k] = bf(max_act);
  for (int k = 0; k < 5; ++k) T.pooled[k] = f32(static_cast<size_t>(B) * 16 * 960);
  T.d_o = f32(static_cast<size_t>(B) * T.Hl * T.Wl * nc);
  T.dh2 = f32(static_cast<size_t>(B) * T.Hh * T.Wh * nc);
  T.bytes = b.off;
}

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

struct Ctx {
  const NetPlan& P; const TrainIO& io; TrainBufs& T; cudaStream_t st; int B;
  const uint8_t* pk() const { return static_cast<const uint8_t*>(io.packed); }
  const bf16* wb(size_t off) const { return reinterpret_cast<const bf16*>(pk() + off); }
  const float* wf(size_t off) const { return reinterpret_cast<const float*>(pk() + off); }
  float* param(int idx) const { return static_cast<float*>(io.params[idx]); }
  float* grad(int idx) const { return io.grads[idx]; }
};

// BatchNorm forward in training mode on an already computed z
int bn_fwd(const Ctx& c, const ConvBnPlan& cp, const LayerBufs& L, int act, const bf16* residual, float* gap, int gap_chunks,
           float momentum, bool stats_done = true) {
  BnTrainFwdArgs a;
  a.z = L.z; a.y = L.y; a.residual = residual;
  a.gamma = c.param(cp.gamma); a.beta = c.param(cp.beta); a.eps = cp.eps; a.momentum = momentum;
  a.running_mean = c.param(cp.mean); a.running_var = c.param(cp.var);
  a.num_batches_tracked = static_cast<long long*>(c.io.params[cp.var + 1]);
  a.scale = L.scale; a.shift = L.shift; a.save_mean = L.mean; a.save_rstd = L.rstd; a.stat = L.stat; a.stats_done = stats_done;
  a.gap = gap; a.gap_chunks = gap_chunks; a.act = act; a.B = c.B; a.HW = L.H * L.W; a.C = L.C;
  return launch_bn_train_fwd(a, c.st);
}

int bn_bwd(const Ctx& c, const ConvBnPlan& cp, const LayerBufs& L, int act, const bf16* dy, bf16* dz, const float* se_s,
           const float* se_dmean) {
  BnTrainBwdArgs a;
  a.z = L.z; a.dy = dy; a.dz = dz; a.scale = L.scale; a.shift = L.shift; a.save_mean = L.mean; a.save_rstd = L.rstd;
  a.se_s = se_s; a.se_dmean = se_dmean; a.bstat = L.bstat; a.bstat_zeroed = true;
  a.dgamma = c.grad(cp.gamma); a.dbeta = c.grad(cp.beta);
  a.act = act; a.B = c.B; a.HW = L.H * L.W; a.C = L.C;
  MTG_REQUIRE(a.dgamma && a.dbeta, MTG_ERR_ARG, "backward: missing gradient buffer for a BatchNorm parameter");
  return launch_bn_train_bwd(a, c.st);
}

// weight gradient: tensor cores (MN-major tcgen05) unless MTGSEG_WGRAD=simt or the shape is outside its tiling
int wgrad(const Ctx& c, WgradArgs& w, int hw) {
  static const bool simt = [] { const char* e = getenv("MTGSEG_WGRAD"); return e && e[0] == 's'; }();
  if (!w.hw) w.hw = hw;
  if (simt || (w.taps == 9 && w.W > 64)) return launch_wgrad(w, c.st);
  return launch_wgrad_tc(w, c.B, c.st);
}

// `stat`: the BatchNorm accumulators of the layer this conv feeds (forward only): filled by the GEMM epilogue
int conv1x1_raw(const Ctx& c, const bf16* a, const bf16* w, bf16* out, int M, int N, int K, const float* a_scale, int hw,
                const bf16* residual, double* stat = nullptr) {
  ConvGemmArgs g;
  g.a = a; g.w = w; g.out = out; g.M = M; g.N = N; g.K = K; g.act = ACT_NONE; g.a_scale = a_scale; g.hw = hw; g.residual = residual;
  g.stat = stat;
  return launch_conv_gemm(g, c.st);
}

}  // namespace

size_t train_workspace_bytes(const NetPlan& P, int batch) {
  TrainBufs T;
  layout(P, batch, nullptr, T);
  return T.bytes;
}

constexpr int kBigBatch = 128;  // per-GPU batch from which the step is throughput bound, not launch-chain bound

int run_train_forward(const NetPlan& P, const TrainIO& io, uint8_t* ws, size_t ws_bytes, cudaStream_t st) {
  TrainBufs T;
  layout(P, io.batch, ws, T);
  // Large per-GPU batches: no programmatic dependent launch (see PdlScope; the two measured points put the break-even at ~105
  // images).  MTGSEG_BN_EPI=0 (A/B): statistics from a separate pass over z instead of the conv epilogues (measured: the epilogue
  // statistics win at B=32, 4.84 -> 4.69 ms; at B=256 the separate pass wins only WITH dependent launch, 23.09 -> 22.79 ms, and
  // loses without it, 22.53 -> 22.81 ms).
  PdlScope pdl_scope(io.batch < kBigBatch);
  static const bool epi_stats = [] { const char* e = getenv("MTGSEG_BN_EPI"); return !(e && e[0] == '0'); }();
  auto SD = [&](double* p) { return epi_stats ? p : nullptr; };
  MTG_REQUIRE(T.bytes <= ws_bytes, MTG_ERR_WORKSPACE, "forward_train: workspace too small: need %zu bytes, got %zu", T.bytes, ws_bytes);
  const int B = io.batch;
  Ctx c{P, io, T, st, B};
  const float mom_bb = 1e-2f, mom_head = 0.1f;
  RC(launch_fill_f32(T.ones, 1.f, static_cast<size_t>(B) * 960, st));
  RC(launch_fill_f32(T.zeros, 0.f, static_cast<size_t>(B) * 960, st));
  MTG_CUDA(cudaMemsetAsync(T.stat_arena, 0, T.stat_bytes, st));
  {
    StemArgs a;
    a.x = io.x; a.w = c.wf(P.stem.w_off); a.scale = T.ones; a.shift = T.zeros; a.out = T.stem.z;
    a.B = B; a.H = P.desc.in_h; a.W = P.desc.in_w; a.act = ACT_NONE;
    RC(launch_stem(a, st));
    RC(bn_fwd(c, P.stem, T.stem, ACT_HSWISH, nullptr, nullptr, 1, mom_bb, /*stats_done=*/false));
  }
  const bf16* t = T.stem.y;
  for (int i = 0; i < kNumBlocks; ++i) {
    const BlockPlan& b = P.blocks[i];
    const BlockCfg& cf = b.cfg;
    const BlockBufs& K = T.blk[i];
    const bf16* inp = t;
    const bf16* e = t;
    if (b.has_expand) {
      RC(conv1x1_raw(c, t, c.wb(b.expand.w_off), K.expand.z, B * K.Hin * K.Win, cf.cexp, cf.cin, nullptr, 0, nullptr, SD(K.expand.stat)));
      RC(bn_fwd(c, b.expand, K.expand, cf.act, nullptr, nullptr, 1, mom_bb, epi_stats));
      e = K.expand.y;
    }
    {
      const int stride = cf.dil > 1 ? 1 : cf.stride;
      DwConvArgs a;
      a.in = e; a.w = c.wb(b.dw.w_off); a.out = K.dw.z; a.scale = T.ones; a.shift = T.zeros; a.act = ACT_NONE;
      a.B = B; a.H = K.Hin; a.W = K.Win; a.C = cf.cexp; a.k = cf.k; a.stride = stride; a.dil = cf.dil;
      a.gap_partial = nullptr; a.chunks = dwconv_chunks(K.Hin, K.Win, cf.cexp, cf.k, stride, cf.dil, false);
      a.stat = SD(K.dw.stat);
      RC(launch_dwconv(a, st));
      RC(bn_fwd(c, b.dw, K.dw, cf.act, nullptr, cf.se ? K.gap : nullptr, K.gap_chunks, mom_bb, epi_stats));
    }
    if (cf.se) {
      SeMlpArgs s;
      s.sums = K.gap; s.chunks = K.gap_chunks; s.B = B; s.C = cf.cexp; s.SQ = b.sq; s.HW = K.dw.H * K.dw.W;
      s.w1 = c.wb(b.fc1_w_off); s.b1 = c.wf(b.fc1_b_off); s.act1 = ACT_RELU;
      s.w2 = c.wb(b.fc2_w_off); s.b2 = c.wf(b.fc2_b_off); s.act2 = ACT_HSIGMOID; s.out = K.s; s.hidden = K.hid;
      RC(launch_se_mlp(s, st));
    }
    RC(conv1x1_raw(c, K.dw.y, c.wb(b.project.w_off), K.project.z, B * K.dw.H * K.dw.W, cf.cout, cf.cexp, cf.se ? K.s : nullptr,
                   K.dw.H * K.dw.W, nullptr, SD(K.project.stat)));
    const bool res = cf.stride == 1 && cf.cin == cf.cout;
    RC(bn_fwd(c, b.project, K.project, ACT_NONE, res ? inp : nullptr, nullptr, 1, mom_bb, epi_stats));
    t = K.project.y;
  }
  const int ic = P.desc.inter_channels, nc = P.desc.num_classes;
  const int Hh = T.Hh, Wh = T.Wh, Mh = B * Hh * Wh;
  RC(conv1x1_raw(c, t, c.wb(P.last.w_off), T.last.z, Mh, 960, 160, nullptr, 0, nullptr, SD(T.last.stat)));
  RC(bn_fwd(c, P.last, T.last, ACT_HSWISH, nullptr, T.hsum, 1, mom_bb, epi_stats));  // pooled sums of `high` for the scale branch
  {
    ConvGemmArgs h;
    h.a = T.last.y; h.w = c.wb(P.cbr.w_off); h.out = T.cbr.z; h.M = Mh; h.N = ic; h.K = 960; h.act = ACT_NONE;
    h.conv3x3 = 1; h.B = B; h.H = Hh; h.W = Wh;
    h.stat = SD(T.cbr.stat);
    RC(launch_conv_gemm(h, st));
    RC(bn_fwd(c, P.cbr, T.cbr, ACT_RELU, nullptr, nullptr, 1, mom_head, epi_stats));
    SeMlpArgs s;
    s.sums = T.hsum; s.chunks = 1; s.B = B; s.C = 960; s.SQ = ic; s.HW = Hh * Wh;
    s.w1 = c.wb(P.scale_w_off); s.b1 = nullptr; s.act1 = ACT_SIGMOID; s.w2 = nullptr; s.out = T.hscale;
    RC(launch_se_mlp(s, st));
    HeadMixArgs m;
    m.cbr = T.cbr.y; m.s = T.hscale; m.low = T.blk[3].project.y; m.w_high = c.wf(P.high_w_off); m.b_high = c.wf(P.high_b_off);
    m.w_low = c.wf(P.low_w_off); m.b_low = c.wf(P.low_b_off); m.out = T.lowres;
    m.B = B; m.Hh = Hh; m.Wh = Wh; m.Hl = T.Hl; m.Wl = T.Wl; m.IC = ic; m.LC = 40; m.NC = nc;
    RC(launch_head_mix(m, st));
    if (io.logits) {  // (a captured training step takes its loss from the low-resolution logits: run_train_loss)
      UpsampleOutArgs u;
      u.lowres = T.lowres; u.logits = io.logits; u.logits_dtype = io.logits_dtype;
      u.B = B; u.Hl = T.Hl; u.Wl = T.Wl; u.H = P.desc.in_h; u.W = P.desc.in_w; u.NC = nc;
      RC(launch_upsample_out(u, st));
    }
  }
  return MTG_OK;
}

// CombinedLoss of the forward that last used `ws`, from its low-resolution logits; leaves dLoss/d(lowres) where the backward
// pass expects the pulled-back gradient (mtgseg_backward with dlogits == NULL).
int run_train_loss(const NetPlan& P, int batch, const int64_t* targets, float* loss3, float dice_w, float ce_w, float smooth, uint8_t* ws,
                   size_t ws_bytes, cudaStream_t st) {
  TrainBufs T;
  layout(P, batch, ws, T);
  MTG_REQUIRE(T.bytes <= ws_bytes, MTG_ERR_WORKSPACE, "train_loss: workspace too small: need %zu bytes, got %zu", T.bytes, ws_bytes);
  return launch_lowres_loss(T.lowres, targets, T.d_o, T.pooled[4], loss3, batch, T.Hl, T.Wl, P.desc.in_h, P.desc.in_w, P.desc.num_classes,
                            dice_w, ce_w, smooth, st);
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradients on a second stream.  At B = 32 the backward chain (BN backward -> dgrad -> BN backward ...) is a
// sequence of small kernels that leave most SMs idle, and the weight-gradient kernels (tcgen05 wgrad, depthwise wgrad, stem
// wgrad: ~21 % of the step in the ncu launch list) only need a layer's dz and its saved input; nothing downstream waits for
// them.  They run on a private non-blocking stream: "dz ready" events main -> side, "dz buffer free" events side -> main (dz
// rotates through four buffers), and the main stream waits for the side stream before the call returns.
// MTGSEG_TRAIN_SIDE=0 keeps everything on one stream (A/B).
// ---------------------------------------------------------------------------------------------------------
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ready[4] = {}, freed[4] = {}, done = nullptr, se_ready = nullptr;
  bool ok = false;
};
SideStream* side_stream() {
  static const bool enabled = [] { const char* e = getenv("MTGSEG_TRAIN_SIDE"); return !(e && e[0] == '0'); }();
  if (!enabled) return nullptr;
  static SideStream per_dev[16];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  SideStream& S = per_dev[dev];
  if (!S.ok) {
    if (cudaStreamCreateWithFlags(&S.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int k = 0; k < 4; ++k) {
      if (cudaEventCreateWithFlags(&S.ready[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&S.freed[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&S.se_ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    S.ok = true;
  }
  return &S;
}

int run_train_backward(const NetPlan& P, const TrainIO& io, uint8_t* ws, size_t ws_bytes, cudaStream_t st) {
  PdlScope pdl_scope(io.batch < kBigBatch);  // see run_train_forward
  TrainBufs T;
  layout(P, io.batch, ws, T);
  MTG_REQUIRE(T.bytes <= ws_bytes, MTG_ERR_WORKSPACE, "backward: workspace too small: need %zu bytes, got %zu", T.bytes, ws_bytes);
  const int B = io.batch;
  Ctx c{P, io, T, st, B};
  const int ic = P.desc.inter_channels, nc = P.desc.num_classes;
  const int Hh = T.Hh, Wh = T.Wh, Hl = T.Hl, Wl = T.Wl, Mh = B * Hh * Wh;
  const int H = P.desc.in_h, W = P.desc.in_w;
  auto need = [&](int idx) { return io.grads[idx] != nullptr; };
  SideStream* side = side_stream();
  Ctx cs{P, io, T, side ? side->s : st, B};  // context of the weight-gradient launches
  int dz_next = 0, dz_cur = 0;
  bool freed_valid[4] = {false, false, false, false};
  // next dz buffer; the main stream first waits until the side-stream reader of its previous contents is done
  auto new_dz = [&]() -> bf16* {
    dz_cur = side ? (dz_next++ & 3) : 0;
    if (side && freed_valid[dz_cur]) cudaStreamWaitEvent(st, side->freed[dz_cur], 0);
    return T.dzpool[dz_cur];
  };
  // hand the current dz buffer (just produced on the main stream) to the side stream; call before the side launch
  auto dz_to_side = [&]() {
    if (!side) return;
    cudaEventRecord(side->ready[dz_cur], st);
    cudaStreamWaitEvent(side->s, side->ready[dz_cur], 0);
  };
  auto dz_side_done = [&]() {  // call after the side launch that read the current dz buffer
    if (!side) return;
    cudaEventRecord(side->freed[dz_cur], side->s);
    freed_valid[dz_cur] = true;
  };
  MTG_REQUIRE(need(P.high_w) && need(P.high_b) && need(P.low_w) && need(P.low_b) && need(P.scale_w) && need(P.cbr.w_idx) &&
                  need(P.last.w_idx) && need(P.stem.w_idx), MTG_ERR_ARG, "backward: missing gradient buffers");
  MTG_CUDA(cudaMemsetAsync(T.bstat_arena, 0, T.bstat_bytes, st));
  // ---- data parallel: gradient buckets in reverse execution order -------------------------------------------
  // grads[] are views of one flat buffer in state_dict order, so a bucket is the slice between the first parameter of its
  // first layer and the first parameter of the next bucket.  A bucket is handed to the communication stream when both the main
  // and the weight-gradient stream have issued everything that writes into it; the exchange of the head (1.4 M floats) and of
  // blocks 14-15 (1.6 M) runs under the backward pass of the early, large layers; only the last 0.1 M floats are exposed.
  auto first_param = [&](int blk) { return P.blocks[blk].has_expand ? P.blocks[blk].expand.w_idx : P.blocks[blk].dw.w_idx; };
  float* bucket_hi = io.dp ? io.flat_grad + io.flat_floats : nullptr;
  auto fire_bucket_from = [&](int param_idx) -> int {  // all-reduce [grads[param_idx], bucket_hi) and lower bucket_hi
    if (!io.dp) return MTG_OK;
    float* lo = io.grads[param_idx];
    MTG_REQUIRE(lo != nullptr && lo >= io.flat_grad && lo <= bucket_hi, MTG_ERR_ARG,
                "backward: gradient pointers are not views of the flat buffer in state_dict order");
    const int rc = dp_fire_bucket(lo, static_cast<size_t>(bucket_hi - lo), st, side ? side->s : nullptr);
    bucket_hi = lo;
    return rc;
  };
  if (io.dp) MTG_REQUIRE(io.flat_grad != nullptr && io.grads[P.stem.w_idx] == io.flat_grad, MTG_ERR_ARG,
                         "backward: data-parallel exchange needs the flat gradient buffer (state_dict order)");

  // ---- head tail ------------------------------------------------------------------------------------
  // dlogits [B][NC][H][W] -> d_o [B][Hl*Wl][NC] -> dh2 [B][Hh*Wh][NC]
  if (io.dlogits)  // else: run_train_loss already left dLoss/d(lowres) in T.d_o
    RC(launch_upsample_bwd(io.dlogits, io.dlogits_dtype, T.d_o, B, nc, Hl, Wl, H, W, static_cast<long long>(nc) * H * W,
                           static_cast<long long>(H) * W, 1, st));
  RC(launch_upsample_bwd(T.d_o, LOGITS_F32, T.dh2, B, nc, Hh, Wh, Hl, Wl, static_cast<long long>(Hl) * Wl * nc, 1, nc, st));
  bf16* dcbr = T.g[0];
  bf16* dlow = T.g[4];  // kept until block 4's output gradient is formed
  float* ds_head = T.pooled[0];  // [B][head_bwd_segments(B)][ic] partial sums
  {
    HeadBwdArgs a;
    a.d_o = T.d_o; a.dh2 = T.dh2; a.cbr = T.cbr.y; a.s = T.hscale; a.low = T.blk[3].project.y;
    a.w_high = c.wf(P.high_w_off); a.w_low = c.wf(P.low_w_off);
    a.dcbr = dcbr; a.ds = ds_head; a.dlow = dlow;
    a.dw_high = c.grad(P.high_w); a.dw_low = c.grad(P.low_w); a.db_high = c.grad(P.high_b); a.db_low = c.grad(P.low_b);
    a.B = B; a.Hh = Hh; a.Wh = Wh; a.Hl = Hl; a.Wl = Wl; a.IC = ic; a.LC = 40; a.NC = nc;
    RC(launch_head_bwd(a, st));
  }
  float* dpre_s = T.pooled[1];
  float* dgap_high = T.pooled[2];  // [B][960]
  {
    SeBwdArgs a;
    a.ds_partial = ds_head; a.chunks = head_bwd_segments(B); a.s = T.hscale; a.w1 = c.param(P.scale_w); a.w2 = nullptr;
    a.dpre1 = dpre_s; a.dmean = dgap_high; a.B = B; a.C = 960; a.SQ = ic;
    RC(launch_se_bwd(a, st));
    RC(launch_outer_sum(dpre_s, T.hsum, 1, 1.f / static_cast<float>(Hh * Wh), c.grad(P.scale_w), nullptr, B, ic, 960, st));
  }
  // cbr: BN(ReLU) backward, wgrad 3x3, dgrad 3x3
  bf16* dz_cbr = new_dz();
  RC(bn_bwd(c, P.cbr, T.cbr, ACT_RELU, dcbr, dz_cbr, nullptr, nullptr));
  {
    WgradArgs w;
    w.dz = dz_cbr; w.x = T.last.y; w.dw = c.grad(P.cbr.w_idx); w.M = Mh; w.N = ic; w.K = 960; w.taps = 9; w.H = Hh; w.W = Wh;
    dz_to_side();
    RC(wgrad(cs, w, Hh * Wh));
    dz_side_done();
    ConvGemmArgs g;
    g.a = dz_cbr; g.w = c.wb(P.cbr.wt_off); g.out = T.g[0]; g.M = Mh; g.N = 960; g.K = ic; g.act = ACT_NONE;
    g.conv3x3 = 1; g.B = B; g.H = Hh; g.W = Wh;
    RC(launch_conv_gemm(g, st));
  }
  // last 1x1 (160 -> 960): d high = conv dgrad + broadcast of the pooled gradient of the scale branch
  bf16* dz_last = new_dz();
  RC(bn_bwd(c, P.last, T.last, ACT_HSWISH, T.g[0], dz_last, T.ones, dgap_high));
  const bf16* last_in = T.blk[kNumBlocks - 1].project.y;
  {
    WgradArgs w;
    w.dz = dz_last; w.x = last_in; w.dw = c.grad(P.last.w_idx); w.M = Mh; w.N = 960; w.K = 160;
    dz_to_side();
    RC(wgrad(cs, w, Hh * Wh));
    dz_side_done();
  }
  bf16* d_out = T.g[0];  // gradient w.r.t. the current block's output
  RC(conv1x1_raw(c, dz_last, c.wb(P.last.wt_off), d_out, Mh, 160, 960, nullptr, 0, nullptr));
  RC(fire_bucket_from(P.last.w_idx));  // bucket A: features[16] + the whole head

  // ---- inverted residual blocks, last to first -------------------------------------------------------
  // buffer roles: d_out = g[0] ; dz_p / dz_d = g[1] ; da / dy_e = g[2] ; dz_e = g[3] ; next d_out is written to g[3]->swap
  for (int i = kNumBlocks - 1; i >= 0; --i) {
    const BlockPlan& b = P.blocks[i];
    const BlockCfg& cf = b.cfg;
    const BlockBufs& K = T.blk[i];
    const int Mo = B * K.dw.H * K.dw.W, Mi = B * K.Hin * K.Win, HWo = K.dw.H * K.dw.W;
    const bool res = cf.stride == 1 && cf.cin == cf.cout;
    const bf16* inp = i == 0 ? T.stem.y : T.blk[i - 1].project.y;
    MTG_REQUIRE(need(b.project.w_idx) && need(b.dw.w_idx), MTG_ERR_ARG, "backward: missing gradient buffers (block %d)", i + 1);
    if (i == 3) {  // features[4] also feeds the head's low classifier
      RC(launch_add_bf16(d_out, dlow, d_out, static_cast<size_t>(Mo) * cf.cout, st));
    }
    bf16* dz_p = new_dz();
    RC(bn_bwd(c, b.project, K.project, ACT_NONE, d_out, dz_p, nullptr, nullptr));
    {
      WgradArgs w;
      w.dz = dz_p; w.x = K.dw.y; w.dw = c.grad(b.project.w_idx); w.M = Mo; w.N = cf.cout; w.K = cf.cexp;
      w.a_scale = cf.se ? K.s : nullptr; w.hw = HWo;
      dz_to_side();
      RC(wgrad(cs, w, HWo));
      dz_side_done();
    }
    bf16* da = T.g[2];
    RC(conv1x1_raw(c, dz_p, c.wb(b.project.wt_off), da, Mo, cf.cexp, cf.cout, nullptr, 0, nullptr));
    const float* se_s = nullptr;
    const float* se_dmean = nullptr;
    if (cf.se) {
      float* ds_part = T.pooled[0];
      const int chunks = K.gap_chunks;
      RC(launch_dot_pool(da, K.dw.y, ds_part, B, HWo, cf.cexp, chunks, st));
      SeBwdArgs a;
      a.ds_partial = ds_part; a.chunks = chunks; a.s = K.s; a.hid = K.hid;
      a.w1 = c.param(b.fc1_w); a.w2 = c.param(b.fc2_w);
      a.dpre2 = K.dpre2; a.dpre1 = K.dpre1; a.dmean = T.pooled[3];
      a.B = B; a.C = cf.cexp; a.SQ = b.sq;
      RC(launch_se_bwd(a, st));
      MTG_REQUIRE(need(b.fc1_w) && need(b.fc1_b) && need(b.fc2_w) && need(b.fc2_b), MTG_ERR_ARG, "backward: missing SE gradient buffers");
      // the FC weight gradients only need dpre2 / dpre1 (per-block buffers) and saved forward state: side stream
      if (side) {
        cudaEventRecord(side->se_ready, st);
        cudaStreamWaitEvent(side->s, side->se_ready, 0);
      }
      RC(launch_outer_sum(K.dpre2, K.hid, 1, 1.f, c.grad(b.fc2_w), c.grad(b.fc2_b), B, cf.cexp, b.sq, cs.st));
      RC(launch_outer_sum(K.dpre1, K.gap, chunks, 1.f / static_cast<float>(HWo), c.grad(b.fc1_w), c.grad(b.fc1_b), B, b.sq,
                          cf.cexp, cs.st));
      se_s = K.s;
      se_dmean = T.pooled[3];
    }
    bf16* dz_d = new_dz();
    RC(bn_bwd(c, b.dw, K.dw, cf.act, da, dz_d, se_s, se_dmean));
    const bf16* dw_in = b.has_expand ? K.expand.y : inp;
    DwBwdArgs d;
    d.dz = dz_d; d.x = dw_in; d.w = c.wb(b.dw.w_off); d.dw = c.grad(b.dw.w_idx);
    d.B = B; d.H = K.Hin; d.W = K.Win; d.C = cf.cexp; d.k = cf.k; d.stride = cf.dil > 1 ? 1 : cf.stride; d.dil = cf.dil;
    dz_to_side();
    RC(launch_dw_wgrad(d, cs.st));
    dz_side_done();
    bf16* dy_e = T.g[2];
    d.dx = dy_e;
    if (d.stride == 1) {
      // dgrad of a stride-1 depthwise conv is the same conv with the taps reversed: reuse the staged forward kernel
      DwConvArgs f;
      f.in = dz_d; f.w = c.wb(b.dw.wt_off); f.out = dy_e; f.scale = T.ones; f.shift = T.zeros; f.act = ACT_NONE;
      f.B = B; f.H = K.Hin; f.W = K.Win; f.C = cf.cexp; f.k = cf.k; f.stride = 1; f.dil = cf.dil;
      f.chunks = dwconv_chunks(K.Hin, K.Win, cf.cexp, cf.k, 1, cf.dil, false);
      RC(launch_dwconv(f, st));
    } else {
      RC(launch_dw_dgrad(d, st));
    }
    bf16* d_inp = T.g[3];
    if (b.has_expand) {
      MTG_REQUIRE(need(b.expand.w_idx), MTG_ERR_ARG, "backward: missing gradient buffers (block %d expand)", i + 1);
      bf16* dz_e = new_dz();
      RC(bn_bwd(c, b.expand, K.expand, cf.act, dy_e, dz_e, nullptr, nullptr));
      WgradArgs w;
      w.dz = dz_e; w.x = inp; w.dw = c.grad(b.expand.w_idx); w.M = Mi; w.N = cf.cexp; w.K = cf.cin;
      dz_to_side();
      RC(wgrad(cs, w, K.Hin * K.Win));
      dz_side_done();
      RC(conv1x1_raw(c, dz_e, c.wb(b.expand.wt_off), d_inp, Mi, cf.cin, cf.cexp, nullptr, 0, res ? d_out : nullptr));
    } else if (res) {
      RC(launch_add_bf16(dy_e, d_out, d_inp, static_cast<size_t>(Mi) * cf.cin, st));
    } else {
      d_inp = dy_e;
    }
    // rotate: the input gradient becomes the next (earlier) block's output gradient, kept in g[0]
    if (d_inp == T.g[3]) { bf16* tmp = T.g[0]; T.g[0] = T.g[3]; T.g[3] = tmp; }
    else { bf16* tmp = T.g[0]; T.g[0] = T.g[2]; T.g[2] = tmp; }
    d_out = T.g[0];
    if (i == 13 || i == 7) RC(fire_bucket_from(first_param(i)));  // bucket B: blocks 14-15 ; bucket C: blocks 8-13
  }
  // ---- stem -----------------------------------------------------------------------------------------
  bf16* dz_stem = new_dz();
  RC(bn_bwd(c, P.stem, T.stem, ACT_HSWISH, d_out, dz_stem, nullptr, nullptr));
  dz_to_side();
  RC(launch_stem_wgrad(io.x, dz_stem, c.grad(P.stem.w_idx), B, H, W, cs.st));
  dz_side_done();
  RC(fire_bucket_from(P.stem.w_idx));  // bucket D: stem + blocks 1-7 (0.1 M floats: the only exposed exchange)
  if (side) {  // every gradient is complete before the caller's next kernel on `st` (optimizer, all-reduce)
    MTG_CUDA(cudaEventRecord(side->done, side->s));
    MTG_CUDA(cudaStreamWaitEvent(st, side->done, 0));
  }
  if (io.dp) RC(dp_join(st));
  return MTG_OK;
}

#undef RC

}  // namespace mtgseg
