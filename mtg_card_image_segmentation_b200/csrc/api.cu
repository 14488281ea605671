// extern "C" surface declared in include/mtgseg_b200.h.
#include "net.h"

using namespace mtgseg;

namespace {
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
int plan_for(const mtgseg_net_desc* d, NetPlan& P) {
  MTG_REQUIRE(d != nullptr, MTG_ERR_ARG, "net desc is NULL");
  return build_plan(*d, P);
}
}  // namespace

extern "C" {

int mtgseg_version(void) { return MTGSEG_ABI_VERSION; }
const char* mtgseg_last_error(void) { return get_error(); }

int mtgseg_param_count(void) {
  NetPlan P;
  mtgseg_net_desc d{320, 240, 2, 128};
  return build_plan(d, P) == MTG_OK ? P.n_params : -1;
}

size_t mtgseg_packed_bytes(const mtgseg_net_desc* desc) {
  NetPlan P;
  return plan_for(desc, P) == MTG_OK ? P.packed_bytes : 0;
}

size_t mtgseg_workspace_bytes(const mtgseg_net_desc* desc, int batch) {
  NetPlan P;
  if (plan_for(desc, P) != MTG_OK || batch <= 0) return 0;
  InferIO io;
  io.batch = batch;
  size_t need = 0;
  if (run_infer(P, io, nullptr, 0, &need, nullptr) != MTG_OK) return 0;
  return need;
}

int mtgseg_pack_weights(const mtgseg_net_desc* desc, const void* const* params, int n_params, void* packed, void* stream) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE(params && packed, MTG_ERR_ARG, "pack_weights: null pointer");
  MTG_REQUIRE(n_params == P.n_params, MTG_ERR_ARG, "pack_weights: expected %d state_dict entries, got %d", P.n_params, n_params);
  for (int i = 0; i < n_params; ++i) MTG_REQUIRE(params[i] != nullptr, MTG_ERR_ARG, "pack_weights: params[%d] is NULL", i);
  return pack_weights(P, params, packed, S(stream));
}

static int forward_infer_impl(const mtgseg_net_desc* desc, const float* x, const uint8_t* x_u8, const void* packed, void* logits,
                              int logits_dtype, uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace,
                              size_t workspace_bytes, int batch, void* stream) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE((x || x_u8) && packed && workspace, MTG_ERR_ARG, "forward_infer: null pointer");
  MTG_REQUIRE(batch > 0, MTG_ERR_ARG, "forward_infer: batch must be positive");
  MTG_REQUIRE(logits || mask || counts4, MTG_ERR_ARG, "forward_infer: no output requested");
  MTG_REQUIRE(!logits || (logits_dtype >= LOGITS_F32 && logits_dtype <= LOGITS_F16), MTG_ERR_ARG, "forward_infer: bad logits dtype %d", logits_dtype);
  MTG_REQUIRE(!counts4 || targets, MTG_ERR_ARG, "forward_infer: counts4 needs targets");
  MTG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, MTG_ERR_ARG, "forward_infer: workspace must be 256-byte aligned");
  InferIO io;
  io.x = x; io.x_u8 = x_u8; io.packed = packed; io.logits = logits; io.logits_dtype = logits_dtype; io.mask = mask;
  io.counts4 = counts4; io.targets = targets; io.batch = batch;
  size_t need = 0;
  rc = run_infer(P, io, nullptr, 0, &need, nullptr);
  if (rc) return rc;
  MTG_REQUIRE(need <= workspace_bytes, MTG_ERR_WORKSPACE, "forward_infer: workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  return run_infer(P, io, static_cast<uint8_t*>(workspace), workspace_bytes, nullptr, S(stream));
}

int mtgseg_forward_infer(const mtgseg_net_desc* desc, const float* x, const void* packed, void* logits, int logits_dtype,
                         uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace, size_t workspace_bytes,
                         int batch, void* stream) {
  return forward_infer_impl(desc, x, nullptr, packed, logits, logits_dtype, mask, counts4, targets, workspace, workspace_bytes, batch, stream);
}

int mtgseg_forward_infer_u8(const mtgseg_net_desc* desc, const uint8_t* x_hwc, const void* packed, void* logits, int logits_dtype,
                            uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace, size_t workspace_bytes,
                            int batch, void* stream) {
  return forward_infer_impl(desc, nullptr, x_hwc, packed, logits, logits_dtype, mask, counts4, targets, workspace, workspace_bytes, batch, stream);
}

size_t mtgseg_workspace_bytes_f32(const mtgseg_net_desc* desc, int batch) {
  NetPlan P;
  if (plan_for(desc, P) != MTG_OK || batch <= 0) return 0;
  InferF32IO io;
  io.batch = batch;
  size_t need = 0;
  if (run_infer_f32(P, io, nullptr, 0, &need, nullptr) != MTG_OK) return 0;
  return need;
}

int mtgseg_forward_infer_f32(const mtgseg_net_desc* desc, const float* x, const void* const* params, int n_params, void* logits,
                             int logits_dtype, uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace,
                             size_t workspace_bytes, int batch, void* stream) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE(x && params && workspace, MTG_ERR_ARG, "forward_infer_f32: null pointer");
  MTG_REQUIRE(n_params == P.n_params, MTG_ERR_ARG, "forward_infer_f32: expected %d state_dict entries, got %d", P.n_params, n_params);
  for (int i = 0; i < n_params; ++i) MTG_REQUIRE(params[i] != nullptr, MTG_ERR_ARG, "forward_infer_f32: params[%d] is NULL", i);
  MTG_REQUIRE(batch > 0, MTG_ERR_ARG, "forward_infer_f32: batch must be positive");
  MTG_REQUIRE(logits || mask || counts4, MTG_ERR_ARG, "forward_infer_f32: no output requested");
  MTG_REQUIRE(!logits || (logits_dtype >= LOGITS_F32 && logits_dtype <= LOGITS_F16), MTG_ERR_ARG, "forward_infer_f32: bad logits dtype %d", logits_dtype);
  MTG_REQUIRE(!counts4 || targets, MTG_ERR_ARG, "forward_infer_f32: counts4 needs targets");
  MTG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, MTG_ERR_ARG, "forward_infer_f32: workspace must be 256-byte aligned");
  InferF32IO io;
  io.x = x; io.params = params; io.logits = logits; io.logits_dtype = logits_dtype; io.mask = mask; io.counts4 = counts4;
  io.targets = targets; io.batch = batch;
  return run_infer_f32(P, io, static_cast<uint8_t*>(workspace), workspace_bytes, nullptr, S(stream));
}

size_t mtgseg_train_workspace_bytes(const mtgseg_net_desc* desc, int batch) {
  NetPlan P;
  if (plan_for(desc, P) != MTG_OK || batch <= 0) return 0;
  return train_workspace_bytes(P, batch);
}

int mtgseg_forward_train(const mtgseg_net_desc* desc, const float* x, const void* packed, void* const* params, int n_params,
                         void* logits, int logits_dtype, void* workspace, size_t workspace_bytes, int batch, void* stream) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE(x && packed && params && workspace, MTG_ERR_ARG, "forward_train: null pointer");
  MTG_REQUIRE(n_params == P.n_params, MTG_ERR_ARG, "forward_train: expected %d state_dict entries, got %d", P.n_params, n_params);
  MTG_REQUIRE(batch > 0, MTG_ERR_ARG, "forward_train: batch must be positive");
  MTG_REQUIRE(static_cast<long long>(batch) * (desc->in_h / 16) * (desc->in_w / 16) > 1, MTG_ERR_UNSUPPORTED,
              "forward_train: BatchNorm needs more than one value per channel");
  MTG_REQUIRE(!logits || (logits_dtype >= LOGITS_F32 && logits_dtype <= LOGITS_F16), MTG_ERR_ARG, "forward_train: bad logits dtype");
  MTG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, MTG_ERR_ARG, "forward_train: workspace must be 256-byte aligned");
  TrainIO io;
  io.x = x; io.packed = packed; io.params = params; io.logits = logits; io.logits_dtype = logits_dtype; io.batch = batch;
  return run_train_forward(P, io, static_cast<uint8_t*>(workspace), workspace_bytes, S(stream));
}

int mtgseg_backward(const mtgseg_net_desc* desc, const float* x, const void* packed, void* const* params, float* const* grads,
                    int n_params, const void* dlogits, int dlogits_dtype, void* workspace, size_t workspace_bytes, int batch,
                    float* flat_grad, size_t flat_floats, int dp_allreduce, void* stream) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE(x && packed && params && grads && workspace, MTG_ERR_ARG, "backward: null pointer");
  MTG_REQUIRE(n_params == P.n_params, MTG_ERR_ARG, "backward: expected %d state_dict entries, got %d", P.n_params, n_params);
  MTG_REQUIRE(!dlogits || (dlogits_dtype >= LOGITS_F32 && dlogits_dtype <= LOGITS_F16), MTG_ERR_ARG, "backward: bad dlogits dtype");
  TrainIO io;
  io.x = x; io.packed = packed; io.params = params; io.grads = grads; io.dlogits = dlogits; io.dlogits_dtype = dlogits_dtype;
  io.batch = batch;
  io.flat_grad = flat_grad; io.flat_floats = flat_floats; io.dp = dp_allreduce;
  MTG_REQUIRE(!dp_allreduce || (flat_grad && flat_floats > 0), MTG_ERR_ARG, "backward: dp_allreduce needs flat_grad / flat_floats");
  return run_train_backward(P, io, static_cast<uint8_t*>(workspace), workspace_bytes, S(stream));
}

size_t mtgseg_loss_lowres_scratch_floats(int batch, int Hl, int Wl) { return lowres_loss_scratch_floats(batch, Hl, Wl); }

int mtgseg_loss_lowres(const float* lowres, const int64_t* targets, float* d_lowres, float* scratch, float* loss3, int batch, int Hl, int Wl,
                       int H, int W, int num_classes, float dice_weight, float ce_weight, float smooth, void* stream) {
  return launch_lowres_loss(lowres, targets, d_lowres, scratch, loss3, batch, Hl, Wl, H, W, num_classes, dice_weight, ce_weight, smooth,
                            S(stream));
}

int mtgseg_train_loss(const mtgseg_net_desc* desc, const int64_t* targets, float* loss3, float dice_weight, float ce_weight, float smooth,
                      void* workspace, size_t workspace_bytes, int batch, void* stream) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE(targets && loss3 && workspace && batch > 0, MTG_ERR_ARG, "train_loss: bad arguments");
  return run_train_loss(P, batch, targets, loss3, dice_weight, ce_weight, smooth, static_cast<uint8_t*>(workspace), workspace_bytes, S(stream));
}

int mtgseg_adamw_step(const void* chunk_table, int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, const float* inv_scale, const float* found_inf, void* stream) {
  return launch_adamw(chunk_table, n_chunks, lr, beta1, beta2, eps, weight_decay, step, inv_scale, found_inf, S(stream));
}

int mtgseg_adamw_hyper(float* hyper, float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  return launch_adamw_hyper(hyper, lr, beta1, beta2, eps, weight_decay, step, S(stream));
}

int mtgseg_adamw_step_dev(const void* chunk_table, int n_chunks, const float* hyper, void* stream) {
  return launch_adamw_dev(chunk_table, n_chunks, hyper, S(stream));
}

// ---- per-operator training entry points (unit tests) ---------------------------------------------------
size_t mtgseg_bn_scratch_floats(int B, int HW, int C) { return bn_partial_floats(B, HW, C); }

int mtgseg_bn_train_fwd(const void* z, void* y, const void* residual, const float* gamma, const float* beta, float eps,
                        float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, float* scale,
                        float* shift, float* save_mean, float* save_rstd, float* scratch, float* gap, int gap_chunks, int act,
                        int B, int HW, int C, void* stream) {
  BnTrainFwdArgs a;
  a.z = static_cast<const bf16*>(z); a.y = static_cast<bf16*>(y); a.residual = static_cast<const bf16*>(residual);
  a.gamma = gamma; a.beta = beta; a.eps = eps; a.momentum = momentum; a.running_mean = running_mean; a.running_var = running_var;
  a.num_batches_tracked = reinterpret_cast<long long*>(num_batches_tracked);
  a.scale = scale; a.shift = shift; a.save_mean = save_mean; a.save_rstd = save_rstd;
  a.stat = reinterpret_cast<double*>(scratch); a.stats_done = false;  // scratch (8-byte aligned) holds the fp64 accumulators
  MTG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, MTG_ERR_ARG, "bn_train_fwd: scratch must be 8-byte aligned");
  a.gap = gap; a.gap_chunks = gap_chunks; a.act = act; a.B = B; a.HW = HW; a.C = C;
  return launch_bn_train_fwd(a, S(stream));
}

int mtgseg_bn_train_bwd(const void* z, const void* dy, void* dz, const float* scale, const float* shift, const float* save_mean,
                        const float* save_rstd, const float* se_s, const float* se_dmean, float* scratch, float* dgamma,
                        float* dbeta, int act, int B, int HW, int C, void* stream) {
  BnTrainBwdArgs a;
  a.z = static_cast<const bf16*>(z); a.dy = static_cast<const bf16*>(dy); a.dz = static_cast<bf16*>(dz);
  a.scale = scale; a.shift = shift; a.save_mean = save_mean; a.save_rstd = save_rstd; a.se_s = se_s; a.se_dmean = se_dmean;
  MTG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, MTG_ERR_ARG, "bn_train_bwd: scratch must be 8-byte aligned");
  a.bstat = reinterpret_cast<double*>(scratch); a.bstat_zeroed = false; a.dgamma = dgamma; a.dbeta = dbeta;
  a.act = act; a.B = B; a.HW = HW; a.C = C;
  return launch_bn_train_bwd(a, S(stream));
}

int mtgseg_wgrad(const void* dz, const void* x, float* dw, const float* a_scale, int hw, int64_t M, int N, int K, int taps, int H,
                 int W, void* stream) {
  WgradArgs a;
  a.dz = static_cast<const bf16*>(dz); a.x = static_cast<const bf16*>(x); a.dw = dw; a.a_scale = a_scale; a.hw = hw;
  a.M = M; a.N = N; a.K = K; a.taps = taps; a.H = H; a.W = W;
  return launch_wgrad(a, S(stream));
}

int mtgseg_wgrad_tc(const void* dz, const void* x, float* dw, const float* a_scale, int B, int hw, int N, int K, int taps, int H,
                    int W, void* stream) {
  WgradArgs a;
  a.dz = static_cast<const bf16*>(dz); a.x = static_cast<const bf16*>(x); a.dw = dw; a.a_scale = a_scale; a.hw = hw;
  a.M = static_cast<long long>(B) * hw; a.N = N; a.K = K; a.taps = taps; a.H = H; a.W = W;
  return launch_wgrad_tc(a, B, S(stream));
}

int mtgseg_dw_bwd(const void* dz, const void* x, const void* w, void* dx, float* dw, int B, int H, int W, int C, int k, int stride,
                  int dil, void* stream) {
  DwBwdArgs a;
  a.dz = static_cast<const bf16*>(dz); a.x = static_cast<const bf16*>(x); a.w = static_cast<const bf16*>(w);
  a.dx = static_cast<bf16*>(dx); a.dw = dw; a.B = B; a.H = H; a.W = W; a.C = C; a.k = k; a.stride = stride; a.dil = dil;
  int rc = MTG_OK;
  if (dx) rc = launch_dw_dgrad(a, S(stream));
  if (rc == MTG_OK && dw) rc = launch_dw_wgrad(a, S(stream));
  return rc;
}

int mtgseg_stem_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, void* stream) {
  return launch_stem_wgrad(x, static_cast<const bf16*>(dz), dw, B, H, W, S(stream));
}

int mtgseg_upsample_bwd(const void* g, int dtype, float* out, int B, int NC, int Hc, int Wc, int Hf, int Wf, void* stream) {
  return launch_upsample_bwd(g, dtype, out, B, NC, Hc, Wc, Hf, Wf, static_cast<long long>(NC) * Hf * Wf,
                             static_cast<long long>(Hf) * Wf, 1, S(stream));
}

size_t mtgseg_se_bwd_scratch_floats(int B, int C, int SQ) { return static_cast<size_t>(B) * (16 * C + C + SQ); }

int mtgseg_se_block_bwd(const void* da, const void* y, const float* s, const float* hid, const float* gap, int gap_chunks,
                        const float* w1, const float* w2, float* dmean, float* dw1, float* db1, float* dw2, float* db2,
                        float* scratch, int B, int HW, int C, int SQ, void* stream) {
  MTG_REQUIRE(da && y && s && hid && gap && w1 && w2 && dmean && dw1 && db1 && dw2 && db2 && scratch, MTG_ERR_ARG, "se_block_bwd: null pointer");
  MTG_REQUIRE(gap_chunks >= 1 && gap_chunks <= 16, MTG_ERR_ARG, "se_block_bwd: gap_chunks must be in [1,16]");
  float* ds_part = scratch;                                     // [B][chunks][C]
  float* dpre2 = scratch + static_cast<size_t>(B) * 16 * C;     // [B][C]
  float* dpre1 = dpre2 + static_cast<size_t>(B) * C;            // [B][SQ]
  int rc = launch_dot_pool(static_cast<const bf16*>(da), static_cast<const bf16*>(y), ds_part, B, HW, C, gap_chunks, S(stream));
  if (rc) return rc;
  SeBwdArgs a;
  a.ds_partial = ds_part; a.chunks = gap_chunks; a.s = s; a.hid = hid; a.w1 = w1; a.w2 = w2;
  a.dpre2 = dpre2; a.dpre1 = dpre1; a.dmean = dmean; a.B = B; a.C = C; a.SQ = SQ;
  rc = launch_se_bwd(a, S(stream));
  if (rc) return rc;
  rc = launch_outer_sum(dpre2, hid, 1, 1.f, dw2, db2, B, C, SQ, S(stream));
  if (rc) return rc;
  return launch_outer_sum(dpre1, gap, gap_chunks, 1.f / static_cast<float>(HW), dw1, db1, B, SQ, C, S(stream));
}

int mtgseg_head_bwd_segments(int B) { return head_bwd_segments(B); }

int mtgseg_head_bwd(const float* d_lowres, const float* d_h2, const void* cbr, const float* s, const void* low, const float* w_high,
                    const float* w_low, void* dcbr, float* ds, void* dlow, float* dw_high, float* dw_low, float* db_high,
                    float* db_low, int B, int Hh, int Wh, int Hl, int Wl, int IC, int LC, int NC, void* stream) {
  MTG_REQUIRE(d_lowres && d_h2 && cbr && s && low && w_high && w_low && dcbr && ds && dlow && dw_high && dw_low && db_high && db_low,
              MTG_ERR_ARG, "head_bwd: null pointer");
  HeadBwdArgs a;
  a.d_o = d_lowres; a.dh2 = d_h2; a.cbr = static_cast<const bf16*>(cbr); a.s = s; a.low = static_cast<const bf16*>(low);
  a.w_high = w_high; a.w_low = w_low; a.dcbr = static_cast<bf16*>(dcbr); a.ds = ds; a.dlow = static_cast<bf16*>(dlow);
  a.dw_high = dw_high; a.dw_low = dw_low; a.db_high = db_high; a.db_low = db_low;
  a.B = B; a.Hh = Hh; a.Wh = Wh; a.Hl = Hl; a.Wl = Wl; a.IC = IC; a.LC = LC; a.NC = NC;
  return launch_head_bwd(a, S(stream));
}

int mtgseg_dp_unique_id(void* out128) { return dp_unique_id(out128); }
int mtgseg_dp_init(const void* id128, int rank, int world) { return dp_init(id128, rank, world); }
int mtgseg_dp_world(void) { return dp_world(); }
int mtgseg_dp_shutdown(void) { return dp_shutdown(); }
int mtgseg_dp_allreduce_avg(float* buf, size_t n, void* stream) {
  MTG_REQUIRE(buf != nullptr, MTG_ERR_ARG, "dp_allreduce_avg: null pointer");
  int rc = dp_fire_bucket(buf, n, S(stream), nullptr);
  if (rc) return rc;
  return dp_join(S(stream));
}

unsigned long long mtgseg_launch_count(void) { return launch_count(); }

int mtgseg_forward_infer_profiled(const mtgseg_net_desc* desc, const float* x, const void* packed, void* logits,
                                  int logits_dtype, uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace,
                                  size_t workspace_bytes, int batch, void* stream, mtgseg_layer_prof* out, int max_layers,
                                  int* n_layers) {
  NetPlan P;
  int rc = plan_for(desc, P);
  if (rc) return rc;
  MTG_REQUIRE(x && packed && workspace && out && n_layers, MTG_ERR_ARG, "forward_infer_profiled: null pointer");
  MTG_REQUIRE(batch > 0 && (logits || mask || counts4), MTG_ERR_ARG, "forward_infer_profiled: bad arguments");
  InferIO io;
  io.x = x; io.packed = packed; io.logits = logits; io.logits_dtype = logits_dtype; io.mask = mask;
  io.counts4 = counts4; io.targets = targets; io.batch = batch;
  size_t need = 0;
  rc = run_infer(P, io, nullptr, 0, &need, nullptr);
  if (rc) return rc;
  MTG_REQUIRE(need <= workspace_bytes, MTG_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  LayerProfiler prof;
  prof.st = S(stream);
  rc = run_infer(P, io, static_cast<uint8_t*>(workspace), workspace_bytes, nullptr, S(stream), &prof);
  cudaError_t e = cudaStreamSynchronize(S(stream));
  int n = 0;
  for (auto& r : prof.recs) {
    if (n < max_layers && e == cudaSuccess && rc == MTG_OK) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.e0, r.e1);
      snprintf(out[n].name, sizeof(out[n].name), "%s", r.name);
      snprintf(out[n].kernel, sizeof(out[n].kernel), "%s", r.kernel);
      out[n].ms = ms; out[n].bytes = r.bytes; out[n].flops = r.flops;
      ++n;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  *n_layers = n;
  if (rc) return rc;
  MTG_REQUIRE(e == cudaSuccess, MTG_ERR_CUDA, "forward_infer_profiled: %s", cudaGetErrorString(e));
  return MTG_OK;
}

int mtgseg_metric_counts(const void* logits, int logits_dtype, const int64_t* targets, uint64_t* counts4, int64_t batch,
                         int64_t hw, void* stream) {
  return launch_metric_counts(logits, logits_dtype, targets, reinterpret_cast<unsigned long long*>(counts4), batch, hw, S(stream));
}

size_t mtgseg_loss_scratch_bytes(void) { return loss_scratch_bytes(); }

int mtgseg_loss_fwd_bwd(const void* logits, int logits_dtype, const int64_t* targets, void* dlogits, int dlogits_dtype,
                        float* scratch, float* loss3, int64_t batch, int64_t hw, int num_classes, float dice_weight,
                        float ce_weight, float smooth, void* stream) {
  return launch_loss(logits, logits_dtype, targets, dlogits, dlogits_dtype, scratch, loss3, batch, hw, num_classes, dice_weight,
                     ce_weight, smooth, S(stream));
}

int mtgseg_conv1x1(const void* a, const void* w, void* out, int M, int N, int K, const float* scale, const float* shift,
                   int act, const void* residual, const float* a_scale, int hw, void* stream) {
  ConvGemmArgs g;
  g.a = static_cast<const bf16*>(a); g.w = static_cast<const bf16*>(w); g.out = static_cast<bf16*>(out);
  g.M = M; g.N = N; g.K = K; g.scale = scale; g.shift = shift; g.act = act;
  g.residual = static_cast<const bf16*>(residual); g.a_scale = a_scale; g.hw = hw;
  return launch_conv_gemm(g, S(stream));
}

int mtgseg_conv3x3(const void* a, const void* w, void* out, int B, int H, int W, int N, int K, const float* scale,
                   const float* shift, int act, void* stream) {
  ConvGemmArgs g;
  g.a = static_cast<const bf16*>(a); g.w = static_cast<const bf16*>(w); g.out = static_cast<bf16*>(out);
  g.M = B * H * W; g.N = N; g.K = K; g.scale = scale; g.shift = shift; g.act = act;
  g.conv3x3 = 1; g.B = B; g.H = H; g.W = W;
  return launch_conv_gemm(g, S(stream));
}

int mtgseg_dwconv_chunks(int H, int W, int C, int k, int stride, int dil, int need_gap) {
  return dwconv_chunks(H, W, C, k, stride, dil, need_gap != 0);
}

int mtgseg_dwconv(const void* in, const void* w, void* out, int B, int H, int W, int C, int k, int stride, int dil,
                  const float* scale, const float* shift, int act, float* gap_partial, int chunks, void* stream) {
  DwConvArgs a;
  a.in = static_cast<const bf16*>(in); a.w = static_cast<const bf16*>(w); a.out = static_cast<bf16*>(out);
  a.scale = scale; a.shift = shift; a.act = act; a.B = B; a.H = H; a.W = W; a.C = C; a.k = k; a.stride = stride; a.dil = dil;
  a.gap_partial = gap_partial; a.chunks = chunks;
  return launch_dwconv(a, S(stream));
}

int mtgseg_stem(const float* x, const float* w, const float* scale, const float* shift, void* out, int B, int H, int W,
                void* stream) {
  StemArgs a;
  a.x = x; a.w = w; a.scale = scale; a.shift = shift; a.out = static_cast<bf16*>(out); a.B = B; a.H = H; a.W = W;
  return launch_stem(a, S(stream));
}

int mtgseg_se_mlp(const float* sums, int chunks, int B, int C, int SQ, int HW, const void* w1, const float* b1, int act1,
                  const void* w2, const float* b2, int act2, float* out, float* hidden, void* stream) {
  SeMlpArgs a;
  a.hidden = hidden;
  a.sums = sums; a.chunks = chunks; a.B = B; a.C = C; a.SQ = SQ; a.HW = HW;
  a.w1 = static_cast<const bf16*>(w1); a.b1 = b1; a.act1 = act1; a.w2 = static_cast<const bf16*>(w2); a.b2 = b2; a.act2 = act2;
  a.out = out;
  return launch_se_mlp(a, S(stream));
}

int mtgseg_gap(const void* in, float* out, int B, int HW, int C, void* stream) {
  return launch_gap(static_cast<const bf16*>(in), out, B, HW, C, S(stream));
}

int mtgseg_head_mix(const void* cbr, const float* s, const void* low, const float* w_high, const float* b_high,
                    const float* w_low, const float* b_low, float* out, int B, int Hh, int Wh, int Hl, int Wl, int IC, int LC,
                    int NC, void* stream) {
  HeadMixArgs a;
  a.cbr = static_cast<const bf16*>(cbr); a.s = s; a.low = static_cast<const bf16*>(low);
  a.w_high = w_high; a.b_high = b_high; a.w_low = w_low; a.b_low = b_low; a.out = out;
  a.B = B; a.Hh = Hh; a.Wh = Wh; a.Hl = Hl; a.Wl = Wl; a.IC = IC; a.LC = LC; a.NC = NC;
  return launch_head_mix(a, S(stream));
}

int mtgseg_upsample_out(const float* lowres, void* logits, int logits_dtype, uint8_t* mask, const int64_t* targets,
                        uint64_t* counts4, int B, int Hl, int Wl, int H, int W, int NC, void* stream) {
  UpsampleOutArgs a;
  a.lowres = lowres; a.logits = logits; a.logits_dtype = logits_dtype; a.mask = mask; a.targets = targets;
  a.counts = reinterpret_cast<unsigned long long*>(counts4);
  a.B = B; a.Hl = Hl; a.Wl = Wl; a.H = H; a.W = W; a.NC = NC;
  return launch_upsample_out(a, S(stream));
}

}  // extern "C"
