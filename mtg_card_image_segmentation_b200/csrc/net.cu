// The network plan: walks the fixed LR-ASPP / dilated MobileNetV3-Large topology (tv:models/mobilenetv3.py:233-251,
// tv:models/segmentation/lraspp.py:43-93, train/model.py:92-142), assigns offsets in the packed-weight arena and
// in the activation workspace, and enqueues the kernels of one inference forward on a stream.
#include "net.h"

#include <stdlib.h>

#include <vector>

namespace mtgseg {

namespace {

const BlockCfg kBlocks[kNumBlocks] = {
    {16, 3, 16, 16, false, ACT_RELU, 1, 1},     {16, 3, 64, 24, false, ACT_RELU, 2, 1},
    {24, 3, 72, 24, false, ACT_RELU, 1, 1},     {24, 5, 72, 40, true, ACT_RELU, 2, 1},
    {40, 5, 120, 40, true, ACT_RELU, 1, 1},     {40, 5, 120, 40, true, ACT_RELU, 1, 1},
    {40, 3, 240, 80, false, ACT_HSWISH, 2, 1},  {80, 3, 200, 80, false, ACT_HSWISH, 1, 1},
    {80, 3, 184, 80, false, ACT_HSWISH, 1, 1},  {80, 3, 184, 80, false, ACT_HSWISH, 1, 1},
    {80, 3, 480, 112, true, ACT_HSWISH, 1, 1},  {112, 3, 672, 112, true, ACT_HSWISH, 1, 1},
    {112, 5, 672, 160, true, ACT_HSWISH, 2, 2}, {160, 5, 960, 160, true, ACT_HSWISH, 1, 2},
    {160, 5, 960, 160, true, ACT_HSWISH, 1, 2},
};

int make_divisible8(int v) {  // tv:models/_utils.py:76-89 with divisor 8
  int nv = (v + 4) / 8 * 8;
  if (nv < 8) nv = 8;
  if (nv * 10 < 9 * v) nv += 8;
  return nv;
}

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

void plan_convbn(ConvBnPlan& c, int& pi, Bump& arena, size_t w_elems, size_t w_elem_bytes, int cout, float eps,
                 bool dgrad_copy = false, int cin = 0) {
  c.cin = cin;
  if (dgrad_copy) c.wt_off = arena.take(w_elems * w_elem_bytes);
  c.w_idx = pi++;
  c.gamma = pi; c.beta = pi + 1; c.mean = pi + 2; c.var = pi + 3;
  pi += 5;  // + num_batches_tracked
  c.cout = cout;
  c.eps = eps;
  c.w_off = arena.take(w_elems * w_elem_bytes);
  c.scale_off = arena.take(cout * sizeof(float));
  c.shift_off = arena.take(cout * sizeof(float));
}

}  // namespace

const BlockCfg* block_table() { return kBlocks; }

// MTGSEG_PIXEL_PACK=0 (A/B): narrow 1x1 layers as plain [M][16] GEMMs
static bool pixel_packing_enabled() {
  static const bool on = [] { const char* e = getenv("MTGSEG_PIXEL_PACK"); return !(e && e[0] == '0'); }();
  return on;
}

void LayerProfiler::begin(const char* name, const char* kernel, double bytes, double flops) {
  Rec r{};
  snprintf(r.name, sizeof(r.name), "%s", name);
  snprintf(r.kernel, sizeof(r.kernel), "%s", kernel);
  r.bytes = bytes; r.flops = flops;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  recs.push_back(r);
}
void LayerProfiler::end() { cudaEventRecord(recs.back().e1, st); }

int build_plan(const mtgseg_net_desc& d, NetPlan& P) {
  MTG_REQUIRE(d.in_h >= 32 && d.in_w >= 32, MTG_ERR_UNSUPPORTED, "input %dx%d too small", d.in_h, d.in_w);
  MTG_REQUIRE(d.num_classes >= 1 && d.num_classes <= 8, MTG_ERR_UNSUPPORTED, "num_classes %d not in [1,8]", d.num_classes);
  MTG_REQUIRE(d.inter_channels % 16 == 0 && d.inter_channels >= 16 && d.inter_channels <= 256, MTG_ERR_UNSUPPORTED,
              "inter_channels %d must be a multiple of 16 in [16,256]", d.inter_channels);
  P.desc = d;
  int pi = 0;
  Bump arena;
  const float eps_bb = 1e-3f, eps_head = 1e-5f;
  plan_convbn(P.stem, pi, arena, 27 * 16, sizeof(float), 16, eps_bb);
  for (int i = 0; i < kNumBlocks; ++i) {
    const BlockCfg& c = kBlocks[i];
    BlockPlan& b = P.blocks[i];
    b.cfg = c;
    b.has_expand = c.cexp != c.cin;
    if (b.has_expand) plan_convbn(b.expand, pi, arena, static_cast<size_t>(c.cexp) * c.cin, 2, c.cexp, eps_bb, true, c.cin);
    plan_convbn(b.dw, pi, arena, static_cast<size_t>(c.cexp) * c.k * c.k, 2, c.cexp, eps_bb, true);  // + tap-reversed copy (dgrad)
    if (c.se) {
      b.sq = make_divisible8(c.cexp / 4);
      b.fc1_w = pi++; b.fc1_b = pi++; b.fc2_w = pi++; b.fc2_b = pi++;
      b.fc1_w_off = arena.take(static_cast<size_t>(b.sq) * c.cexp * 2);
      b.fc1_b_off = arena.take(b.sq * sizeof(float));
      b.fc2_w_off = arena.take(static_cast<size_t>(b.sq) * c.cexp * 2);
      b.fc2_b_off = arena.take(c.cexp * sizeof(float));
    }
    plan_convbn(b.project, pi, arena, static_cast<size_t>(c.cout) * c.cexp, 2, c.cout, eps_bb, true, c.cexp);
    // pixel packing of the narrow full-resolution projections (see ConvBnPlan::pp): 16 -> 16 as 64 -> 64 over four pixels
    // (and 72 -> 24 with its 48-byte output / residual rows as 144 -> 48 over two pixels)
    int want_pp = 1;
    if (!c.se && c.cexp <= 16 && c.cout <= 16) want_pp = 64 / c.cexp;
    else if (!c.se && c.cexp == 72 && c.cout == 24 && c.stride == 1) want_pp = 2;
    if (want_pp > 1 && pixel_packing_enabled()) {
      b.project.pp = want_pp;
      const int pp = b.project.pp;
      b.project.wpp_off = arena.take(static_cast<size_t>(pp) * c.cout * pp * c.cexp * 2);
      b.project.scale_pp_off = arena.take(static_cast<size_t>(pp) * c.cout * sizeof(float));
      b.project.shift_pp_off = arena.take(static_cast<size_t>(pp) * c.cout * sizeof(float));
    }
  }
  plan_convbn(P.last, pi, arena, 960 * 160, 2, 960, eps_bb, true, 160);
  const int ic = d.inter_channels, nc = d.num_classes;
  plan_convbn(P.cbr, pi, arena, static_cast<size_t>(ic) * 960 * 9, 2, ic, eps_head, true, 960);
  P.scale_w = pi++;
  P.scale_w_off = arena.take(static_cast<size_t>(ic) * 960 * 2);
  P.low_w = pi++; P.low_b = pi++; P.high_w = pi++; P.high_b = pi++;
  P.low_w_off = arena.take(nc * 40 * sizeof(float));
  P.low_b_off = arena.take(nc * sizeof(float));
  P.high_w_off = arena.take(static_cast<size_t>(nc) * ic * sizeof(float));
  P.high_b_off = arena.take(nc * sizeof(float));
  P.n_params = pi;
  P.packed_bytes = arena.off;
  return MTG_OK;
}

static int pack_weights_impl(const NetPlan& P, const void* const* params, void* packed, cudaStream_t st) {
  uint8_t* base = static_cast<uint8_t*>(packed);
  auto f = [&](int idx) { return static_cast<const float*>(params[idx]); };
  auto fold = [&](const ConvBnPlan& c) {
    return launch_fold_bn(f(c.gamma), f(c.beta), f(c.mean), f(c.var), c.eps, reinterpret_cast<float*>(base + c.scale_off),
                          reinterpret_cast<float*>(base + c.shift_off), c.cout, st);
  };
  int rc;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
  RC(launch_pack_stem(f(P.stem.w_idx), reinterpret_cast<float*>(base + P.stem.w_off), st));
  RC(fold(P.stem));
  for (int i = 0; i < kNumBlocks; ++i) {
    const BlockPlan& b = P.blocks[i];
    const BlockCfg& c = b.cfg;
    if (b.has_expand) {
      RC(launch_cast_bf16(f(b.expand.w_idx), reinterpret_cast<bf16*>(base + b.expand.w_off), static_cast<size_t>(c.cexp) * c.cin, st));
      RC(launch_pack_transpose_bf16(f(b.expand.w_idx), reinterpret_cast<bf16*>(base + b.expand.wt_off), c.cexp, c.cin, st));
      RC(fold(b.expand));
    }
    RC(launch_pack_dw(f(b.dw.w_idx), reinterpret_cast<bf16*>(base + b.dw.w_off), reinterpret_cast<bf16*>(base + b.dw.wt_off), c.cexp,
                      c.k * c.k, st));
    RC(fold(b.dw));
    if (c.se) {
      const size_t n = static_cast<size_t>(b.sq) * c.cexp;
      RC(launch_cast_bf16(f(b.fc1_w), reinterpret_cast<bf16*>(base + b.fc1_w_off), n, st));
      RC(launch_copy_f32(f(b.fc1_b), reinterpret_cast<float*>(base + b.fc1_b_off), b.sq, st));
      RC(launch_cast_bf16(f(b.fc2_w), reinterpret_cast<bf16*>(base + b.fc2_w_off), n, st));
      RC(launch_copy_f32(f(b.fc2_b), reinterpret_cast<float*>(base + b.fc2_b_off), c.cexp, st));
    }
    RC(launch_cast_bf16(f(b.project.w_idx), reinterpret_cast<bf16*>(base + b.project.w_off), static_cast<size_t>(c.cout) * c.cexp, st));
    RC(launch_pack_transpose_bf16(f(b.project.w_idx), reinterpret_cast<bf16*>(base + b.project.wt_off), c.cout, c.cexp, st));
    RC(fold(b.project));
    if (b.project.pp > 1) {
      const int pp = b.project.pp;
      RC(launch_pack_blockdiag(f(b.project.w_idx), reinterpret_cast<bf16*>(base + b.project.wpp_off), c.cout, c.cexp, pp, st));
      for (int q = 0; q < pp; ++q)
        RC(launch_fold_bn(f(b.project.gamma), f(b.project.beta), f(b.project.mean), f(b.project.var), b.project.eps,
                          reinterpret_cast<float*>(base + b.project.scale_pp_off) + q * c.cout,
                          reinterpret_cast<float*>(base + b.project.shift_pp_off) + q * c.cout, c.cout, st));
    }
  }
  RC(launch_cast_bf16(f(P.last.w_idx), reinterpret_cast<bf16*>(base + P.last.w_off), 960 * 160, st));
  RC(launch_pack_transpose_bf16(f(P.last.w_idx), reinterpret_cast<bf16*>(base + P.last.wt_off), 960, 160, st));
  RC(fold(P.last));
  const int ic = P.desc.inter_channels, nc = P.desc.num_classes;
  RC(launch_pack_oihw_to_otapi(f(P.cbr.w_idx), reinterpret_cast<bf16*>(base + P.cbr.w_off), ic, 960, 9, st));
  RC(launch_pack_dgrad3x3(f(P.cbr.w_idx), reinterpret_cast<bf16*>(base + P.cbr.wt_off), ic, 960, st));
  RC(fold(P.cbr));
  RC(launch_cast_bf16(f(P.scale_w), reinterpret_cast<bf16*>(base + P.scale_w_off), static_cast<size_t>(ic) * 960, st));
  RC(launch_copy_f32(f(P.low_w), reinterpret_cast<float*>(base + P.low_w_off), nc * 40, st));
  RC(launch_copy_f32(f(P.low_b), reinterpret_cast<float*>(base + P.low_b_off), nc, st));
  RC(launch_copy_f32(f(P.high_w), reinterpret_cast<float*>(base + P.high_w_off), static_cast<size_t>(nc) * ic, st));
  RC(launch_copy_f32(f(P.high_b), reinterpret_cast<float*>(base + P.high_b_off), nc, st));
  return MTG_OK;
}

// every conversion of the ~170 tensors is recorded and executed by ONE kernel (pack.cu)
int pack_weights(const NetPlan& P, const void* const* params, void* packed, cudaStream_t st) {
  pack_batch_begin();
  const int rc = pack_weights_impl(P, params, packed, st);
  if (rc != MTG_OK) {
    pack_batch_abort();
    return rc;
  }
  return pack_batch_flush(st);
}

static inline int conv_out(int in, int k, int stride, int dil) {
  const int pad = (k - 1) / 2 * dil;
  return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1;
}

// One inference forward.  With ws == nullptr this is a dry run that only sizes the workspace.
int run_infer(const NetPlan& P, const InferIO& io, uint8_t* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st,
              LayerProfiler* prof) {
  const bool dry = ws == nullptr;
  char nm[48];
  // PROF(kernel, algorithmic bytes, flops, launch): algorithmic bytes = each input read once + each output written once
#define PROF(kernel, bytes, flops, call)                 \
  do {                                                   \
    if (prof) prof->begin(nm, kernel, bytes, flops);     \
    RC(call);                                            \
    if (prof) prof->end();                               \
  } while (0)
  Bump bump;
  const int B = io.batch;
  const uint8_t* pk = static_cast<const uint8_t*>(io.packed);
  auto act_buf = [&](size_t elems) { return reinterpret_cast<bf16*>(ws + bump.take(elems * sizeof(bf16))); };
  auto f32_buf = [&](size_t elems) { return reinterpret_cast<float*>(ws + bump.take(elems * sizeof(float))); };
  auto wb = [&](size_t off) { return reinterpret_cast<const bf16*>(pk + off); };
  auto wf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  int rc;

  int H = conv_out(P.desc.in_h, 3, 2, 1), W = conv_out(P.desc.in_w, 3, 2, 1);
  bf16* t = act_buf(static_cast<size_t>(B) * H * W * 16);
  if (!dry) {
    StemArgs a;
    a.x = io.x; a.x_u8 = io.x_u8; a.w = wf(P.stem.w_off); a.scale = wf(P.stem.scale_off); a.shift = wf(P.stem.shift_off);
    a.out = t; a.B = B; a.H = P.desc.in_h; a.W = P.desc.in_w;
    snprintf(nm, sizeof(nm), "stem");
    PROF("stem", (double)B * (3.0 * P.desc.in_h * P.desc.in_w * 4 + (double)H * W * 16 * 2), 2.0 * B * H * W * 16 * 27, launch_stem(a, st));
  }
  const bf16* low = nullptr;
  int Hl = 0, Wl = 0;
  for (int i = 0; i < kNumBlocks; ++i) {
    const BlockPlan& b = P.blocks[i];
    const BlockCfg& c = b.cfg;
    const bf16* inp = t;
    const bf16* e = t;
    if (b.has_expand) {
      bf16* eb = act_buf(static_cast<size_t>(B) * H * W * c.cexp);
      if (!dry) {
        ConvGemmArgs g;
        g.a = t; g.w = wb(b.expand.w_off); g.out = eb; g.M = B * H * W; g.N = c.cexp; g.K = c.cin;
        g.scale = wf(b.expand.scale_off); g.shift = wf(b.expand.shift_off); g.act = c.act;
        snprintf(nm, sizeof(nm), "b%d.expand %dx%d %d->%d", i + 1, H, W, c.cin, c.cexp);
        PROF("conv_gemm_1x1", 2.0 * g.M * (g.K + g.N) + 2.0 * g.N * g.K, 2.0 * g.M * g.N * g.K, launch_conv_gemm(g, st));
      }
      e = eb;
    }
    const int stride = c.dil > 1 ? 1 : c.stride;
    const int Ho = conv_out(H, c.k, stride, c.dil), Wo = conv_out(W, c.k, stride, c.dil);
    bf16* dwo = act_buf(static_cast<size_t>(B) * Ho * Wo * c.cexp);
    const int chunks = dwconv_chunks(H, W, c.cexp, c.k, stride, c.dil, c.se);
    float* gap = c.se ? f32_buf(static_cast<size_t>(B) * chunks * c.cexp) : nullptr;
    float* sescale = c.se ? f32_buf(static_cast<size_t>(B) * c.cexp) : nullptr;
    float* sehid = c.se ? f32_buf(static_cast<size_t>(B) * b.sq) : nullptr;
    if (!dry) {
      DwConvArgs a;
      a.in = e; a.w = wb(b.dw.w_off); a.out = dwo; a.scale = wf(b.dw.scale_off); a.shift = wf(b.dw.shift_off);
      a.act = c.act; a.B = B; a.H = H; a.W = W; a.C = c.cexp; a.k = c.k; a.stride = stride; a.dil = c.dil;
      a.gap_partial = gap; a.chunks = chunks;
      snprintf(nm, sizeof(nm), "b%d.dw k%d s%d d%d %dx%d C%d", i + 1, c.k, stride, c.dil, H, W, c.cexp);
      PROF("dwconv", 2.0 * B * c.cexp * ((double)H * W + (double)Ho * Wo), 2.0 * B * Ho * Wo * c.cexp * c.k * c.k, launch_dwconv(a, st));
      if (c.se) {
        SeMlpArgs s;
        s.sums = gap; s.chunks = chunks; s.B = B; s.C = c.cexp; s.SQ = b.sq; s.HW = Ho * Wo;
        s.w1 = wb(b.fc1_w_off); s.b1 = wf(b.fc1_b_off); s.act1 = ACT_RELU;
        s.w2 = wb(b.fc2_w_off); s.b2 = wf(b.fc2_b_off); s.act2 = ACT_HSIGMOID; s.out = sescale; s.hidden = sehid;
        snprintf(nm, sizeof(nm), "b%d.se C%d", i + 1, c.cexp);
        PROF("se_mlp", 4.0 * B * c.cexp * (chunks + 1) + 4.0 * b.sq * c.cexp, 4.0 * B * b.sq * c.cexp, launch_se_mlp(s, st));
      }
    }
    H = Ho; W = Wo;
    bf16* o = act_buf(static_cast<size_t>(B) * H * W * c.cout);
    if (!dry) {
      ConvGemmArgs g;
      g.a = dwo; g.w = wb(b.project.w_off); g.out = o; g.M = B * H * W; g.N = c.cout; g.K = c.cexp;
      g.scale = wf(b.project.scale_off); g.shift = wf(b.project.shift_off); g.act = ACT_NONE;
      g.residual = (c.stride == 1 && c.cin == c.cout) ? inp : nullptr;
      g.a_scale = sescale; g.hw = H * W;
      if (b.project.pp > 1 && g.M % b.project.pp == 0) {  // pp pixels per GEMM row, block-diagonal weights: same bytes, 128-byte rows
        const int pp = b.project.pp;
        g.M /= pp; g.K *= pp; g.N *= pp;
        g.w = wb(b.project.wpp_off); g.scale = wf(b.project.scale_pp_off); g.shift = wf(b.project.shift_pp_off);
      }
      snprintf(nm, sizeof(nm), "b%d.project %dx%d %d->%d", i + 1, H, W, c.cexp, c.cout);
      const double pm = static_cast<double>(B) * H * W;  // algorithmic bytes / flops of the layer itself (not of the packed GEMM)
      PROF(sescale ? "conv_gemm_1x1_se" : "conv_gemm_1x1", 2.0 * pm * (c.cexp + c.cout * (g.residual ? 2 : 1)) + 2.0 * c.cout * c.cexp,
           2.0 * pm * c.cout * c.cexp, launch_conv_gemm(g, st));
    }
    t = o;
    if (i == 3) { low = o; Hl = H; Wl = W; }  // features[4] -> 'low' (tv:models/segmentation/lraspp.py:87-91)
  }
  const int ic = P.desc.inter_channels, nc = P.desc.num_classes;
  bf16* high = act_buf(static_cast<size_t>(B) * H * W * 960);
  bf16* cbr = act_buf(static_cast<size_t>(B) * H * W * ic);
  float* hsum = f32_buf(static_cast<size_t>(B) * 960);
  float* hscale = f32_buf(static_cast<size_t>(B) * ic);
  float* lowres = f32_buf(static_cast<size_t>(B) * Hl * Wl * nc);
  if (ws_needed) *ws_needed = bump.off;
  if (dry) return MTG_OK;
  MTG_REQUIRE(bump.off <= ws_bytes, MTG_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", bump.off, ws_bytes);
  {
    ConvGemmArgs g;
    g.a = t; g.w = wb(P.last.w_off); g.out = high; g.M = B * H * W; g.N = 960; g.K = 160;
    g.scale = wf(P.last.scale_off); g.shift = wf(P.last.shift_off); g.act = ACT_HSWISH;
    snprintf(nm, sizeof(nm), "b16.conv %dx%d 160->960", H, W);
    PROF("conv_gemm_1x1", 2.0 * g.M * (g.K + g.N) + 2.0 * g.N * g.K, 2.0 * g.M * g.N * g.K, launch_conv_gemm(g, st));
    ConvGemmArgs h;
    h.a = high; h.w = wb(P.cbr.w_off); h.out = cbr; h.M = B * H * W; h.N = ic; h.K = 960;
    h.scale = wf(P.cbr.scale_off); h.shift = wf(P.cbr.shift_off); h.act = ACT_RELU;
    h.conv3x3 = 1; h.B = B; h.H = H; h.W = W;
    snprintf(nm, sizeof(nm), "head.cbr 3x3 %dx%d 960->%d", H, W, ic);
    PROF("conv_gemm_3x3", 2.0 * h.M * (h.K + h.N) + 2.0 * h.N * h.K * 9, 2.0 * h.M * h.N * h.K * 9, launch_conv_gemm(h, st));
    snprintf(nm, sizeof(nm), "head.gap");
    PROF("gap", 2.0 * B * H * W * 960, 1.0 * B * H * W * 960, launch_gap(high, hsum, B, H * W, 960, st));
    SeMlpArgs s;
    s.sums = hsum; s.chunks = 1; s.B = B; s.C = 960; s.SQ = ic; s.HW = H * W;
    s.w1 = wb(P.scale_w_off); s.b1 = nullptr; s.act1 = ACT_SIGMOID; s.w2 = nullptr; s.out = hscale;
    snprintf(nm, sizeof(nm), "head.scale");
    PROF("se_mlp", 4.0 * B * (960 + ic) + 2.0 * 960 * ic, 2.0 * B * 960 * ic, launch_se_mlp(s, st));
    HeadMixArgs m;
    m.cbr = cbr; m.s = hscale; m.low = low; m.w_high = wf(P.high_w_off); m.b_high = wf(P.high_b_off);
    m.w_low = wf(P.low_w_off); m.b_low = wf(P.low_b_off); m.out = lowres;
    m.B = B; m.Hh = H; m.Wh = W; m.Hl = Hl; m.Wl = Wl; m.IC = ic; m.LC = 40; m.NC = nc;
    snprintf(nm, sizeof(nm), "head.mix");
    PROF("head_mix", 2.0 * B * ((double)H * W * ic + (double)Hl * Wl * 40) + 4.0 * B * Hl * Wl * nc,
         2.0 * B * ((double)H * W * ic + (double)Hl * Wl * 40) * nc, launch_head_mix(m, st));
    UpsampleOutArgs u;
    u.lowres = lowres; u.logits = io.logits; u.logits_dtype = io.logits_dtype; u.mask = io.mask; u.targets = io.targets;
    u.counts = reinterpret_cast<unsigned long long*>(io.counts4);
    u.B = B; u.Hl = Hl; u.Wl = Wl; u.H = P.desc.in_h; u.W = P.desc.in_w; u.NC = nc;
    snprintf(nm, sizeof(nm), "head.upsample_out");
    {
      const double px = (double)B * P.desc.in_h * P.desc.in_w;
      const double lb = io.logits ? (io.logits_dtype == LOGITS_F32 ? 4.0 : 2.0) * nc : 0.0;
      PROF("upsample_out", px * (lb + (io.mask ? 1 : 0) + (io.counts4 ? 8 : 0)) + 4.0 * B * Hl * Wl * nc, 8.0 * px * nc,
           launch_upsample_out(u, st));
    }
  }
  return MTG_OK;
#undef PROF
#undef RC
}

}  // namespace mtgseg
