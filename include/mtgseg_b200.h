/* mtgseg_b200 — C ABI of the B200 (sm_100a) implementation of the card-segmentation hot path.
 *
 * Drop-in boundary for the network the reference builds in train/model.py:18-48 (torchvision LR-ASPP on a
 * dilated MobileNetV3-Large with the reference's 3x3 head, train/model.py:92-142) as driven by
 * train/train.py:67-153 and train/evaluate.py:41-100.  The reference has no FFI of its own (it is pure
 * Python on top of torch/torchvision); each entry point below names the Python call it stands in for.
 * The Python binding a maintainer adds is shown in INTEGRATION.md (ctypes, ~40 lines).
 *
 * Conventions
 *   - plain C: ints, sizes, raw DEVICE pointers; `stream` is a cudaStream_t passed as void*.
 *   - every call is asynchronous on `stream`, allocates nothing, and never synchronises the device.
 *   - the caller owns all memory (weights, packed arena, workspace, outputs).
 *   - return 0 on success, <0 on error (MTGSEG_ERR_*); mtgseg_last_error() gives the thread-local message.
 *     Unsupported shapes / dtypes are errors: there is no fallback path of any kind.
 */
#ifndef MTGSEG_B200_H
#define MTGSEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTGSEG_ABI_VERSION 2

#define MTGSEG_OK 0
#define MTGSEG_ERR_ARG (-1)
#define MTGSEG_ERR_CUDA (-2)
#define MTGSEG_ERR_UNSUPPORTED (-3)
#define MTGSEG_ERR_WORKSPACE (-4)

/* logits element types (output of the network; input of the metric kernel) */
#define MTGSEG_LOGITS_NONE 0
#define MTGSEG_LOGITS_F32 1
#define MTGSEG_LOGITS_BF16 2
#define MTGSEG_LOGITS_F16 3

/* activation codes of the per-op entry points */
#define MTGSEG_ACT_NONE 0
#define MTGSEG_ACT_RELU 1
#define MTGSEG_ACT_HSWISH 2
#define MTGSEG_ACT_HSIGMOID 3
#define MTGSEG_ACT_SIGMOID 4

/* The architecture is fixed by the reference (create_model, train/model.py:145-156); only the input size
 * (train/config.py:21-22), the class count (config.py:20) and the head width (model.py:47) vary. */
typedef struct mtgseg_net_desc {
  int32_t in_h;           /* Config.INPUT_HEIGHT, 320 */
  int32_t in_w;           /* Config.INPUT_WIDTH, 240  */
  int32_t num_classes;    /* Config.NUM_CLASSES, 2    */
  int32_t inter_channels; /* LRASPPHead inter_channels, 128 */
} mtgseg_net_desc;

int mtgseg_version(void);
const char* mtgseg_last_error(void);

/* Number of state_dict entries expected by mtgseg_pack_weights (319). */
int mtgseg_param_count(void);
/* Bytes of the packed-weight arena / of the activation workspace for `batch` images. */
size_t mtgseg_packed_bytes(const mtgseg_net_desc* desc);
size_t mtgseg_workspace_bytes(const mtgseg_net_desc* desc, int batch);

/* model.load_state_dict / optimizer.step aftermath: convert the reference-layout parameters (device
 * pointers in state_dict order: fp32 OIHW convs, BN weight/bias/running_mean/running_var, int64
 * num_batches_tracked (ignored)) into the kernel layouts (bf16 K-major weights, folded BN scale/shift). */
int mtgseg_pack_weights(const mtgseg_net_desc* desc, const void* const* params, int n_params, void* packed, void* stream);

/* CardSegmentationModel.forward in eval mode (train/model.py:79-89; train/evaluate.py:66, train/train.py:142)
 * plus, optionally, what evaluate.py does with the logits: argmax mask (evaluate.py:74) and the 2x2
 * confusion counts against `targets` (utils.py:94-164, evaluate.py:88).
 *   x        float32 [batch,3,in_h,in_w]   (train/dataset.py:84-88)
 *   logits   [batch,num_classes,in_h,in_w] of `logits_dtype`, or NULL
 *   mask     uint8 [batch,in_h,in_w] argmax (ties -> lowest class), or NULL
 *   counts4  uint64[4] {n00,n01,n10,n11} (index = target*2 + prediction), ACCUMULATED; needs targets; or NULL
 *   targets  int64 [batch,in_h,in_w] or NULL */
int mtgseg_forward_infer(const mtgseg_net_desc* desc, const float* x, const void* packed, void* logits, int logits_dtype,
                         uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace, size_t workspace_bytes,
                         int batch, void* stream);

/* The same forward in IEEE float32 end to end: what train/evaluate.py:66 computes (`self.model(images)`, no autocast) and what
 * train/export.py:159 holds the exported graph to (1e-4).  Reads the fp32 parameters directly (`params` = the 319 state_dict
 * device pointers, as for mtgseg_pack_weights), keeps fp32 NHWC activations in `workspace` (mtgseg_workspace_bytes_f32) and
 * accumulates with round-to-nearest FMAs on the CUDA cores; outputs as for mtgseg_forward_infer. */
size_t mtgseg_workspace_bytes_f32(const mtgseg_net_desc* desc, int batch);
int mtgseg_forward_infer_f32(const mtgseg_net_desc* desc, const float* x, const void* const* params, int n_params, void* logits,
                             int logits_dtype, uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace,
                             size_t workspace_bytes, int batch, void* stream);

/* One training step = mtgseg_forward_train -> (loss, mtgseg_loss_fwd_bwd) -> mtgseg_backward -> mtgseg_adamw_step ->
 * mtgseg_pack_weights, all on one stream with one workspace of mtgseg_train_workspace_bytes().
 *
 * mtgseg_forward_train: model.train(); model(images) (train/train.py:82,96-97): batch-statistics BatchNorm; the
 *   running_mean / running_var / num_batches_tracked entries of `params` are UPDATED in place (momentum 0.01 backbone,
 *   0.1 head; unbiased variance); everything backward needs stays in `workspace`.
 * mtgseg_backward: loss.backward() (train/train.py:101,105) for the forward that last used `workspace`.
 *   dlogits [batch,num_classes,in_h,in_w]; grads[i] = fp32 gradient buffer of state_dict entry i in the reference's
 *   layout (OIHW), NULL for buffers; the caller ZEROES them first (several are accumulated with atomics).
 *   dlogits may be NULL after mtgseg_train_loss (below).
 * mtgseg_adamw_step: torch.optim.AdamW.step (train/train.py:167-171) over all tensors in one launch. chunk_table is a
 *   DEVICE array of n_chunks records {float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int32 n;}
 *   (one CTA each; split big tensors into several records). inv_scale / found_inf: GradScaler hooks, may be NULL.
 * mtgseg_train_loss (optional, replaces mtgseg_loss_fwd_bwd in a step that does not need the logits themselves): CombinedLoss
 *   (train/utils.py:58-92) of the forward that last used `workspace`, computed from the head's LOW-RESOLUTION logits -- the final x8
 *   bilinear upsample (tv:models/segmentation/lraspp.py:46) is linear, so the full-resolution logits are recomputed on the fly and
 *   the gradient is pulled back to 40x30 inside the same kernel.  Call mtgseg_forward_train with logits == NULL before it and
 *   mtgseg_backward with dlogits == NULL after it: no full-resolution logits / dlogits tensor exists in such a step. */
size_t mtgseg_train_workspace_bytes(const mtgseg_net_desc* desc, int batch);
int mtgseg_train_loss(const mtgseg_net_desc* desc, const int64_t* targets, float* loss3, float dice_weight, float ce_weight, float smooth,
                      void* workspace, size_t workspace_bytes, int batch, void* stream);
int mtgseg_forward_train(const mtgseg_net_desc* desc, const float* x, const void* packed, void* const* params, int n_params,
                         void* logits, int logits_dtype, void* workspace, size_t workspace_bytes, int batch, void* stream);
int mtgseg_backward(const mtgseg_net_desc* desc, const float* x, const void* packed, void* const* params, float* const* grads,
                    int n_params, const void* dlogits, int dlogits_dtype, void* workspace, size_t workspace_bytes, int batch,
                    float* flat_grad, size_t flat_floats, int dp_allreduce, void* stream);

/* Data-parallel training (one process per GPU; the reference trains on one device, train/config.py:61): the single exchange of
 * a step is the average of the fp32 gradients over the ranks.  mtgseg_backward(..., flat_grad, flat_floats, dp_allreduce = 1)
 * requires grads[] to be views of ONE flat buffer in state_dict order and averages it with NCCL in four buckets (head +
 * features[16]; blocks 14-15; blocks 8-13; stem + blocks 1-7), each launched on a communication stream as soon as its last
 * gradient has been produced, i.e. overlapped with the rest of the backward pass; `stream` waits for the exchange before the
 * call's work is complete.  BatchNorm statistics stay per replica (the reference has no SyncBN).
 *   mtgseg_dp_unique_id  rank 0: 128-byte NCCL unique id, to be broadcast by the host layer (torch.distributed / MPI / files)
 *   mtgseg_dp_init       every rank, collectively: this library's own communicator on the CURRENT device
 *   mtgseg_dp_world      ranks of the communicator, 0 before init
 *   mtgseg_dp_allreduce_avg  the same exchange on a caller-owned buffer (tests, isolated timing), ordered after `stream`
 * NCCL is resolved from the process at run time (libnccl.so.2); everything is stream-ordered and CUDA-graph capturable. */
int mtgseg_dp_unique_id(void* out128);
int mtgseg_dp_init(const void* id128, int rank, int world);
int mtgseg_dp_world(void);
int mtgseg_dp_shutdown(void);
int mtgseg_dp_allreduce_avg(float* buf, size_t n, void* stream);
int mtgseg_adamw_step(const void* chunk_table, int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, const float* inv_scale, const float* found_inf, void* stream);
/* CUDA-graph form of the same step (train/train.py:105 inside a captured training step): the scalars of a captured launch
 * are frozen, so they are read from `hyper`, an 8-float DEVICE block {lr, beta1, beta2, eps, weight_decay, 1-beta1^step,
 * sqrt(1-beta2^step), 0} that mtgseg_adamw_hyper (one 1-thread launch, outside the graph) rewrites before every replay. */
int mtgseg_adamw_hyper(float* hyper, float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream);
int mtgseg_adamw_step_dev(const void* chunk_table, int n_chunks, const float* hyper, void* stream);

/* Same, fed with the RAW image batch: uint8 [batch,in_h,in_w,3] (HWC, what cv2 / the camera delivers, train/dataset.py:66-70);
 * A.Normalize's (v/255 - mean)/std with the ImageNet constants (train/dataset.py:182-185) is fused into the stem's load, so the
 * host->device copy is 4x smaller and no normalised fp32 tensor is ever materialised. */
int mtgseg_forward_infer_u8(const mtgseg_net_desc* desc, const uint8_t* x_hwc, const void* packed, void* logits, int logits_dtype,
                            uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace, size_t workspace_bytes,
                            int batch, void* stream);

/* Measurement aids (bench.py): kernels launched by this library so far in this process, and one forward with
 * CUDA events around every kernel launch (synchronises `stream`; algorithmic bytes/flops per launch as in
 * each input read once, each output written once). */
typedef struct mtgseg_layer_prof {
  char name[48];   /* e.g. "b2.expand 160x120 16->64" */
  char kernel[24]; /* kernel family, e.g. "conv_gemm_1x1" */
  float ms;
  double bytes;
  double flops;
} mtgseg_layer_prof;
unsigned long long mtgseg_launch_count(void);
int mtgseg_forward_infer_profiled(const mtgseg_net_desc* desc, const float* x, const void* packed, void* logits,
                                  int logits_dtype, uint8_t* mask, uint64_t* counts4, const int64_t* targets, void* workspace,
                                  size_t workspace_bytes, int batch, void* stream, mtgseg_layer_prof* out, int max_layers,
                                  int* n_layers);

/* calculate_iou / calculate_dice_coefficient / calculate_pixel_accuracy (train/utils.py:94-164) and
 * sklearn confusion_matrix (train/evaluate.py:88) reduced to their integer core: accumulates the 2x2
 * counts of argmax(logits[batch,2,hw]) vs targets[batch,hw] into counts4 (uint64[4]). */
int mtgseg_metric_counts(const void* logits, int logits_dtype, const int64_t* targets, uint64_t* counts4, int64_t batch,
                         int64_t hw, void* stream);

/* CombinedLoss.forward + backward (train/utils.py:58-92, train/train.py:96-101) in one pass:
 *   loss3[0] = dice_weight*(1 - dice) + ce_weight*CE, loss3[1] = dice loss, loss3[2] = CE
 *   dlogits (logits' layout, may be NULL) = d loss3[0] / d logits, stored as dlogits_dtype: the logits' dtype or
 *           MTGSEG_LOGITS_F32.  fp16 logits (the reference's autocast dtype, train/train.py:96) want F32: the unscaled
 *           gradient (~1e-7 at batch 32) is an fp16 subnormal until GradScaler's factor (train/train.py:101) is applied.
 * logits [batch,num_classes,hw]; targets int64 [batch,hw]; scratch = mtgseg_loss_scratch_bytes() bytes. */
size_t mtgseg_loss_scratch_bytes(void);
int mtgseg_loss_fwd_bwd(const void* logits, int logits_dtype, const int64_t* targets, void* dlogits, int dlogits_dtype,
                        float* scratch, float* loss3, int64_t batch, int64_t hw, int num_classes, float dice_weight,
                        float ce_weight, float smooth, void* stream);

/* The kernel behind mtgseg_train_loss on caller-owned buffers (unit tests): lowres / d_lowres fp32 [batch,Hl,Wl,num_classes];
 * logits = bilinear(lowres -> H x W, align_corners=False); loss3 as above; d_lowres = d loss3[0] / d lowres.
 * scratch >= mtgseg_loss_lowres_scratch_floats(batch, Hl, Wl) floats.  Deterministic (gather form, fixed order). */
size_t mtgseg_loss_lowres_scratch_floats(int batch, int Hl, int Wl);
int mtgseg_loss_lowres(const float* lowres, const int64_t* targets, float* d_lowres, float* scratch, float* loss3, int batch, int Hl, int Wl,
                       int H, int W, int num_classes, float dice_weight, float ce_weight, float smooth, void* stream);

/* ---- per-operator entry points (unit tests, profiling) ------------------------------------------------ */
/* nn.Conv2d 1x1 (+ folded BN, activation, residual, squeeze-excite input scale) on NHWC bf16:
 * out[M,N] = act((a[M,K] . w[N,K]^T) * scale + shift) + residual ; a rows of image b scaled by a_scale[b,K] */
int mtgseg_conv1x1(const void* a, const void* w, void* out, int M, int N, int K, const float* scale, const float* shift,
                   int act, const void* residual, const float* a_scale, int hw, void* stream);
/* nn.Conv2d 3x3 pad 1 (train/model.py:110) on NHWC bf16 [B,H,W,K], weights [N,9,K] */
int mtgseg_conv3x3(const void* a, const void* w, void* out, int B, int H, int W, int N, int K, const float* scale,
                   const float* shift, int act, void* stream);
/* depthwise k x k conv + folded BN + act on NHWC bf16; weights bf16 [k*k,C]; optional per-chunk channel sums
 * gap_partial float[B,chunks,C] with chunks = mtgseg_dwconv_chunks(...) */
int mtgseg_dwconv_chunks(int H, int W, int C, int k, int stride, int dil, int need_gap);
int mtgseg_dwconv(const void* in, const void* w, void* out, int B, int H, int W, int C, int k, int stride, int dil,
                  const float* scale, const float* shift, int act, float* gap_partial, int chunks, void* stream);
/* stem 3x3/s2 conv 3->16 + BN + Hardswish: x fp32 NCHW -> out bf16 NHWC; w fp32 [27,16] */
int mtgseg_stem(const float* x, const float* w, const float* scale, const float* shift, void* out, int B, int H, int W,
                void* stream);
/* pooled MLP: mean = sum(sums[B,chunks,C])/HW ; h = act1(w1 mean + b1) ; out = act2(w2 h + b2) (out = h if w2 NULL);
 * hidden = float[B,SQ] scratch for the two-layer form */
int mtgseg_se_mlp(const float* sums, int chunks, int B, int C, int SQ, int HW, const void* w1, const float* b1, int act1,
                  const void* w2, const float* b2, int act2, float* out, float* hidden, void* stream);
int mtgseg_gap(const void* in, float* out, int B, int HW, int C, void* stream);
int mtgseg_head_mix(const void* cbr, const float* s, const void* low, const float* w_high, const float* b_high,
                    const float* w_low, const float* b_low, float* out, int B, int Hh, int Wh, int Hl, int Wl, int IC, int LC,
                    int NC, void* stream);
int mtgseg_upsample_out(const float* lowres, void* logits, int logits_dtype, uint8_t* mask, const int64_t* targets,
                        uint64_t* counts4, int B, int Hl, int Wl, int H, int W, int NC, void* stream);

/* ---- per-operator training entry points (unit tests) --------------------------------------------------- */
/* nn.BatchNorm2d in train mode (+ activation, residual, per-image channel sums) on NHWC bf16 z[B,HW,C];
 * scratch >= mtgseg_bn_scratch_floats(B,HW,C) floats (+ 2*C for the backward call) */
size_t mtgseg_bn_scratch_floats(int B, int HW, int C);
int mtgseg_bn_train_fwd(const void* z, void* y, const void* residual, const float* gamma, const float* beta, float eps,
                        float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, float* scale,
                        float* shift, float* save_mean, float* save_rstd, float* scratch, float* gap, int gap_chunks, int act,
                        int B, int HW, int C, void* stream);
/* backward of the above: dz from dy (optionally dy' = dy*se_s[b,c] + se_dmean[b,c]/HW), dgamma, dbeta */
int mtgseg_bn_train_bwd(const void* z, const void* dy, void* dz, const float* scale, const float* shift, const float* save_mean,
                        const float* save_rstd, const float* se_s, const float* se_dmean, float* scratch, float* dgamma,
                        float* dbeta, int act, int B, int HW, int C, void* stream);
/* weight gradient of a 1x1 (taps=1) or 3x3 pad-1 (taps=9) convolution: dw[N,K,taps] (fp32 OIHW) += dz[M,N]^T x[M,K] */
int mtgseg_wgrad(const void* dz, const void* x, float* dw, const float* a_scale, int hw, int64_t M, int N, int K, int taps, int H,
                 int W, void* stream);
/* the same on the tensor cores (tcgen05, MN-major operands; what mtgseg_backward uses): dz[B*hw,N], x[B*hw,K] */
int mtgseg_wgrad_tc(const void* dz, const void* x, float* dw, const float* a_scale, int B, int hw, int N, int K, int taps, int H,
                    int W, void* stream);
/* depthwise conv backward: dx (if non-NULL) and dw[C,k*k] (fp32, accumulated, if non-NULL) */
int mtgseg_dw_bwd(const void* dz, const void* x, const void* w, void* dx, float* dw, int B, int H, int W, int C, int k, int stride,
                  int dil, void* stream);
int mtgseg_stem_wgrad(const float* x, const void* dz, float* dw, int B, int H, int W, void* stream);
/* transpose of the align_corners=False bilinear upsample: g[B,NC,Hf,Wf] -> out fp32 [B,Hc,Wc,NC] */
int mtgseg_upsample_bwd(const void* g, int dtype, float* out, int B, int NC, int Hc, int Wc, int Hf, int Wf, void* stream);

/* backward of SqueezeExcitation (tv:ops/misc.py:252-261: s = hardsigmoid(fc2(relu(fc1(mean(y))))), out = s * y) for the part that
 * goes THROUGH the gate, given da = dL/d(out) [B,HW,C] bf16 and the block's saved forward state (y = the depthwise output the gate
 * multiplied, s [B,C], hid = relu(fc1(...)) [B,SQ], gap = per-chunk channel sums of y [B,gap_chunks,C]):
 *   dmean[B,C] = dL/d(mean(y)) (mtgseg_bn_train_bwd adds dmean/HW to every pixel's gradient through se_dmean),
 *   dw2[C,SQ], db2[C], dw1[SQ,C], db1[SQ] = gradients of fc2 / fc1 (OVERWRITTEN).  w1 / w2: the fp32 master weights.
 * scratch >= mtgseg_se_bwd_scratch_floats(B, C, SQ) floats. */
size_t mtgseg_se_bwd_scratch_floats(int B, int C, int SQ);
int mtgseg_se_block_bwd(const void* da, const void* y, const float* s, const float* hid, const float* gap, int gap_chunks,
                        const float* w1, const float* w2, float* dmean, float* dw1, float* db1, float* dw2, float* db2,
                        float* scratch, int B, int HW, int C, int SQ, void* stream);
/* backward of the head tail (train/model.py:137-142; forward = mtgseg_head_mix): d_lowres [B,Hl,Wl,NC] = gradient of the
 * low-resolution logits, d_h2 [B,Hh,Wh,NC] = its x2-bilinear transpose (mtgseg_upsample_bwd).  Writes dcbr [B,Hh,Wh,IC] bf16,
 * dlow [B,Hl,Wl,LC] bf16 and ds [B, mtgseg_head_bwd_segments(B), IC] fp32 = per-pixel-segment partial sums of dL/ds (summed in
 * fixed order by the consumer: this value feeds the activation-gradient chain and must be reproducible); ACCUMULATES (atomics;
 * zero first) the parameter gradients dw_high [NC,IC], dw_low [NC,LC], db_high [NC], db_low [NC]. */
int mtgseg_head_bwd_segments(int B);
int mtgseg_head_bwd(const float* d_lowres, const float* d_h2, const void* cbr, const float* s, const void* low, const float* w_high,
                    const float* w_low, void* dcbr, float* ds, void* dlow, float* dw_high, float* dw_low, float* db_high,
                    float* db_low, int B, int Hh, int Wh, int Hl, int Wl, int IC, int LC, int NC, void* stream);

/* ---- corner-keypoint head of the pose pipeline (BASELINE.json configs[4]) ----------------------------------
 * HRNetPoseHead.forward in eval mode (train-pose-estimation_custom/model.py:10-77) on a backbone feature map and
 * LiteHRNet.decode_heatmaps (model.py:133-164).  params = the 28 state_dict entries of HRNetPoseHead in order.
 *   features  float32 [batch,in_channels,feat_h,feat_w]            (the timm backbone itself is out of scope)
 *   heatmaps  float32 [batch,num_keypoints,out_h,out_w]            (AdaptiveAvgPool2d target, model.py:52)
 *   coords    float32 [batch,2*num_keypoints] x0,y0,x1,y1,... in [0,1] (argmax, lowest index on ties), or NULL */
typedef struct mtgseg_pose_desc {
  int32_t in_channels; /* backbone feature channels */
  int32_t feat_h, feat_w;
  int32_t num_keypoints; /* 4 */
  int32_t out_h, out_w;  /* 120, 160 */
} mtgseg_pose_desc;
int mtgseg_pose_param_count(void);
size_t mtgseg_pose_packed_bytes(const mtgseg_pose_desc* desc);
size_t mtgseg_pose_workspace_bytes(const mtgseg_pose_desc* desc, int batch);
int mtgseg_pose_pack_weights(const mtgseg_pose_desc* desc, const void* const* params, int n_params, void* packed, void* stream);
int mtgseg_pose_forward(const mtgseg_pose_desc* desc, const float* features, const void* packed, float* heatmaps, float* coords,
                        void* workspace, size_t workspace_bytes, int batch, void* stream);
int mtgseg_decode_heatmaps(const float* heatmaps, float* coords, int batch, int num_keypoints, int H, int W, void* stream);
/* CornerMetrics.update (train-pose-estimation_custom/metrics.py:29-73): per (image, keypoint) the argmax of the predicted and of the
 * target heatmap [B,K,H,W] fp32, scaled to image pixels, Euclidean distance.  ACCUMULATES into acc (32 bytes, zero it to reset):
 * { double sum_of_distances; uint64 n; uint64 n_within_3px; uint64 n_within_6px } (compute(): metrics.py:75-100). */
int mtgseg_corner_metrics(const float* pred, const float* target, void* acc, int batch, int num_keypoints, int H, int W, float image_w,
                          float image_h, void* stream);
/* CornerLoss = nn.MSELoss on heatmaps (metrics.py:105-136): loss[0] = mean((pred-target)^2); dpred (optional) = 2 (pred-target) / n.
 * scratch >= mtgseg_mse_scratch_floats() floats.  Deterministic (fixed reduction order). */
size_t mtgseg_mse_scratch_floats(void);
int mtgseg_mse_loss(const float* pred, const float* target, float* dpred, float* loss, float* scratch, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTGSEG_B200_H */
