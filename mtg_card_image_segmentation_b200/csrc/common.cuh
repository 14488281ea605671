// Shared host/device helpers for the mtgseg sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace mtgseg {

typedef __nv_bfloat16 bf16;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_HSWISH = 2, ACT_HSIGMOID = 3, ACT_SIGMOID = 4 };

// error codes of the C-ABI (include/mtgseg_b200.h)
enum : int { MTG_OK = 0, MTG_ERR_ARG = -1, MTG_ERR_CUDA = -2, MTG_ERR_UNSUPPORTED = -3, MTG_ERR_WORKSPACE = -4 };

void set_error(const char* fmt, ...);
const char* get_error();

#define MTG_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::mtgseg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ::mtgseg::MTG_ERR_CUDA;                                                                \
    }                                                                                               \
  } while (0)

void count_launch();
unsigned long long launch_count();
#define MTG_LAUNCH_CHECK()          \
  do {                              \
    ::mtgseg::count_launch();       \
    MTG_CUDA(cudaGetLastError());   \
  } while (0)

#define MTG_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::mtgseg::set_error(__VA_ARGS__);   \
      return (code);                      \
    }                                     \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// The training step at batch 32 is a chain of ~350 short dependent kernels: between two of them the GPU idles for the launch
// latency plus the next kernel's prologue (barrier init, TMEM allocation, descriptor prefetch, constant loads).  Kernels
// launched through launch_pdl() may start while their predecessor in the stream is still draining; they call pdl_wait()
// before they touch global memory (it returns once the predecessor grid has completed and its writes are visible) and
// pdl_trigger() at their top so that THEIR successor can be scheduled early.  Works the same inside a CUDA-graph capture
// (programmatic edges).  MTGSEG_PDL=0 launches everything with full serialisation (A/B).
bool pdl_enabled();
// Launches made by this thread while a PdlScope(false) is alive are fully serialised.  Programmatic dependent launch pays for
// chains of SHORT kernels; with long kernels the early-launched dependents only take CTA slots from their predecessor while
// they sit in griddepcontrol.wait (measured on the training step: B=32 4.97 -> 4.69 ms with it, B=256 22.53 -> 23.09 ms).
class PdlScope {
 public:
  explicit PdlScope(bool allow);
  ~PdlScope();
  PdlScope(const PdlScope&) = delete;
  PdlScope& operator=(const PdlScope&) = delete;
 private:
  bool prev_;
};
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ONLY for kernels that call pdl_wait() before their first global-memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#ifdef __CUDACC__
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_HSWISH: return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case ACT_HSIGMOID: return fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return x;
  }
}

// 8 bf16 <-> 8 floats through one 128-bit register quad
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack2(f[0], f[1]);
  q.y = pack2(f[2], f[3]);
  q.z = pack2(f[4], f[5]);
  q.w = pack2(f[6], f[7]);
  return q;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
#endif

}  // namespace mtgseg
