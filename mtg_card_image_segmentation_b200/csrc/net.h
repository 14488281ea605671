#pragma once
#include "../../include/mtgseg_b200.h"
#include <vector>

#include "ops.h"

namespace mtgseg {

constexpr int kNumBlocks = 15;

struct BlockCfg { int cin, k, cexp, cout; bool se; int act; int stride, dil; };

struct ConvBnPlan {
  int w_idx = -1, gamma = -1, beta = -1, mean = -1, var = -1;  // indices into the state_dict-ordered parameter list
  int cout = 0;
  float eps = 1e-3f;
  size_t w_off = 0, scale_off = 0, shift_off = 0;  // offsets into the packed arena
  size_t wt_off = 0;  // dgrad operand (transposed / flipped weights), 1x1 and 3x3 convs only
  int cin = 0;        // input channels of the dense convs
  // pixel packing (inference, narrow 1x1 layers): pp consecutive pixels form ONE GEMM row of pp * cin channels and the weights
  // become block diagonal [pp * cout][pp * cin] (scale / shift replicated pp times), so that TMA moves 128-byte rows instead
  // of 32-byte ones.  The activations are the same bytes: a [M][C] row-major matrix IS a [M / pp][pp * C] matrix.
  int pp = 1;
  size_t wpp_off = 0, scale_pp_off = 0, shift_pp_off = 0;
};

struct BlockPlan {
  BlockCfg cfg;
  bool has_expand = false;
  ConvBnPlan expand, dw, project;
  int sq = 0, fc1_w = -1, fc1_b = -1, fc2_w = -1, fc2_b = -1;
  size_t fc1_w_off = 0, fc1_b_off = 0, fc2_w_off = 0, fc2_b_off = 0;
};

struct NetPlan {
  mtgseg_net_desc desc;
  ConvBnPlan stem, last, cbr;
  BlockPlan blocks[kNumBlocks];
  int scale_w = -1, low_w = -1, low_b = -1, high_w = -1, high_b = -1;
  size_t scale_w_off = 0, low_w_off = 0, low_b_off = 0, high_w_off = 0, high_b_off = 0;
  int n_params = 0;
  size_t packed_bytes = 0;
};

struct InferIO {
  const float* x = nullptr;
  const uint8_t* x_u8 = nullptr;  // alternative input: raw uint8 HWC pixels
  const void* packed = nullptr;
  void* logits = nullptr;
  int logits_dtype = LOGITS_F32;
  uint8_t* mask = nullptr;
  uint64_t* counts4 = nullptr;
  const int64_t* targets = nullptr;
  int batch = 0;
};

// optional per-launch profile of one forward (bench.py roofline): CUDA events around every kernel launch
struct LayerProfiler {
  struct Rec { char name[48]; char kernel[24]; double bytes, flops; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  cudaStream_t st = nullptr;
  void begin(const char* name, const char* kernel, double bytes, double flops);
  void end();
};

const BlockCfg* block_table();
int build_plan(const mtgseg_net_desc& d, NetPlan& P);
int pack_weights(const NetPlan& P, const void* const* params, void* packed, cudaStream_t st);
int run_infer(const NetPlan& P, const InferIO& io, uint8_t* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st,
              LayerProfiler* prof = nullptr);

// fp32-exact inference (f32net.cu): what train/evaluate.py:66 computes (no autocast); reads the fp32 master parameters directly
struct InferF32IO {
  const float* x = nullptr;            // [B][3][H][W] fp32
  const void* const* params = nullptr; // the 319 state_dict device pointers
  void* logits = nullptr; int logits_dtype = LOGITS_F32;
  uint8_t* mask = nullptr; uint64_t* counts4 = nullptr; const int64_t* targets = nullptr;
  int batch = 0;
};
int run_infer_f32(const NetPlan& P, const InferF32IO& io, uint8_t* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st);

// training step (net_train.cu): forward with batch-statistics BatchNorm keeping what backward needs in `ws`,
// then backward producing fp32 parameter gradients in the reference (state_dict) layout.
struct TrainIO {
  const float* x = nullptr;             // [B][3][H][W] fp32
  const void* packed = nullptr;         // packed arena (mtgseg_pack_weights)
  void* const* params = nullptr;        // 319 state_dict device pointers (gamma/beta read, running stats updated in forward)
  void* logits = nullptr; int logits_dtype = LOGITS_F32;          // forward output
  const void* dlogits = nullptr; int dlogits_dtype = LOGITS_F32;  // backward input
  float* const* grads = nullptr;        // backward: 319 pointers, fp32 grad per state_dict entry (NULL for buffers); pre-zeroed
  int batch = 0;
  // data parallel: grads[] are views of ONE flat buffer in state_dict order; backward averages it over the ranks in buckets on
  // the communication stream (dp_nccl.cu), each fired when its last gradient has been produced
  float* flat_grad = nullptr; size_t flat_floats = 0; int dp = 0;
};

// data-parallel gradient exchange (dp_nccl.cu)
int dp_unique_id(void* out128);
int dp_init(const void* id128, int rank, int world);
int dp_world();
int dp_shutdown();
int dp_fire_bucket(float* buf, size_t n, cudaStream_t main, cudaStream_t side);
int dp_join(cudaStream_t main);
size_t train_workspace_bytes(const NetPlan& P, int batch);
int run_train_forward(const NetPlan& P, const TrainIO& io, uint8_t* ws, size_t ws_bytes, cudaStream_t st);
int run_train_backward(const NetPlan& P, const TrainIO& io, uint8_t* ws, size_t ws_bytes, cudaStream_t st);
int run_train_loss(const NetPlan& P, int batch, const int64_t* targets, float* loss3, float dice_w, float ce_w, float smooth, uint8_t* ws,
                   size_t ws_bytes, cudaStream_t st);

}  // namespace mtgseg
